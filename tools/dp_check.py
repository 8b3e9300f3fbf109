"""Data-parallel correctness on the path that is TIMED (run under torchrun, >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/dp_check.py

Per rank: the full-size stage-1 step (1.31 B-parameter DiT, ViT-L/14-336, AE) on a rank-specific batch with fixed draws.
  1. single-GPU gradients of the eager step (no exchange) -> their sum over ranks through a plain fp32 all_reduce is the
     expectation;
  2. the GRAPHED data-parallel step (graph.PipelinedTrainStep: NCCL buckets captured as a branch of the step graph,
     dynamic tile schedule while they are in flight, weight gradients on the side stream) must leave exactly that sum in
     every rank's flat gradient buffer (cosine >= 0.99999, bf16 rounding of the NCCL sum aside);
  3. after the (captured) clip + AdamW update every rank must hold bit-identical parameters;
  4. host time of one replay (the 70 ms `host_enqueue_ms_per_step` SCALE_r01 showed at N >= 2) split by call.
Prints one JSON line on rank 0 (also gpurun_out/dp_check.json); exit code 1 when a check fails."""
from __future__ import annotations

import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import bench
from genhancer_b200 import optim
from genhancer_b200.graph import PipelinedTrainStep
from genhancer_b200.parallel import GradReducer, broadcast_parameters


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30))


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B = int(os.environ.get("DP_CHECK_BATCH", "8"))
    w = bench.build_workload("img336_stage1", dev, B)
    groups = optim.flatten(w.trainable)
    broadcast_parameters(groups)
    opt = optim.FusedAdamW(groups, lr=1e-4, engine_managed=[w.dit])
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    x = w.make_inputs(gen)
    h = w.image_size // 8
    noise = torch.randn(B, 16, h, h, device=dev, generator=gen)
    t = torch.sigmoid(torch.randn(B, device=dev, generator=gen))
    x_0 = torch.randn(B, (h // 2) ** 2, 64, device=dev, generator=gen)
    fn = lambda img, before_trainable=None: w.step(img, before_trainable=before_trainable, ae_noise=noise, t=t, x_0=x_0)  # noqa: E731
    # 1. single-GPU gradients, summed over ranks in fp32
    opt.zero_grad()
    loss = fn(x[0])
    loss.backward()
    del loss
    torch.cuda.synchronize()
    expect = [g.flat_g.float() for g in groups]
    for e in expect:
        dist.all_reduce(e)
    local_norm = [float(g.flat_g.float().norm()) for g in groups]
    opt.zero_grad()
    # 2. the graphed data-parallel step
    reducer = GradReducer(groups, engine_modules=[w.dit])
    p_before = [g.flat_p.clone() for g in groups]
    pipe = PipelinedTrainStep(fn, (x[0],), opt, reducer)
    t0 = time.perf_counter()
    pipe(x[0])
    t_replay0 = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
    res = {"world": world, "batch_per_rank": B, "grad_cos": [], "grad_rel": [], "buckets": len(reducer.log)}
    ok = True
    for g, e in zip(groups, expect):
        c = cos(g.flat_g, e)
        r = float((g.flat_g.float() - e).norm() / (e.norm() + 1e-30))
        res["grad_cos"].append(round(c, 7))
        res["grad_rel"].append(round(r, 6))
        ok &= c >= 0.99999
    for g, p in zip(groups, p_before):
        ok &= bool(torch.equal(g.flat_p, p))        # the first replay applies nothing (no update pending)
    # 3. two more replays + flush: three updates; every rank must hold the same bits
    host = []
    for _ in range(2):
        t0 = time.perf_counter()
        pipe(x[0])
        host.append((time.perf_counter() - t0) * 1e3)
    t0 = time.perf_counter()
    torch.cuda.synchronize()
    t_sync = (time.perf_counter() - t0) * 1e3
    pipe.flush()
    torch.cuda.synchronize()
    same = True
    moved = True
    for g, p in zip(groups, p_before):
        ref = g.flat_p.clone()
        dist.broadcast(ref, src=0)
        same &= bool(torch.equal(ref, g.flat_p))
        moved &= not torch.equal(g.flat_p, p)
    flag = torch.tensor([int(same)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res["ranks_bit_identical_after_adamw"] = bool(flag.item())
    res["params_moved"] = moved
    res["optimizer_steps"] = opt.step_count
    ok &= bool(flag.item()) and moved and opt.step_count == 3
    # 4. host time per call of a replay
    res["host_ms"] = {"first_replay": round(t_replay0, 2), "replay": [round(v, 2) for v in host], "sync_after": round(t_sync, 2)}
    t0 = time.perf_counter()
    for _ in range(5):
        pipe.static_in[0].copy_(x[0], non_blocking=True)
    res["host_ms"]["input_copy"] = round((time.perf_counter() - t0) * 1e3 / 5, 3)
    g0 = next(iter(pipe.graphs.values()))
    torch.cuda.synchronize()
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        g0.replay()
        times.append(round((time.perf_counter() - t0) * 1e3, 2))
    torch.cuda.synchronize()
    res["host_ms"]["graph_replay_call"] = times
    res["local_grad_norms"] = [round(v, 4) for v in local_norm]
    res["ok"] = bool(ok)
    if rank == 0:
        print(json.dumps(res), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(res, open(os.path.join(ROOT, "gpurun_out", "dp_check.json"), "w"), indent=1)
    threading.Timer(30.0, lambda: os._exit(0 if ok else 1)).start()
    pipe.reset()
    dist.barrier()
    dist.destroy_process_group()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
