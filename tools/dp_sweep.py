#!/usr/bin/env python
"""Data-parallel overlap sweep (run under torchrun, N >= 2): one model build, then for each
(NCCL max CTAs, SM budget of our persistent grids, deferred exchange yes/no) a few timed steps of bench.py's step.

NCCL's all-reduce kernels need whole SMs (their shared memory does not fit beside a 227 KB GEMM CTA); a persistent
GEMM grid of 148 CTAs that finds n SMs taken runs n CTAs as a second wave.  The sweep finds the (few CTAs for NCCL,
148 - n SMs for us) point at which the exchange hides for free.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_sweep.py
"""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402


def main():
    from genhancer_b200 import _lib, kernels as K, optim
    from genhancer_b200.parallel import GradReducer, broadcast_parameters

    world, rank, local = (int(os.environ[k]) for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, S = 32, 336
    step, clip_vis, dit, vae = bench.build_models(S, dev)
    trainable = list(dit.named_parameters()) + [(f"clip_vis.{n}", p) for n, p in clip_vis.named_parameters()]
    groups = optim.flatten(trainable)
    broadcast_parameters(groups)
    opt = optim.FusedAdamW(groups, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, max_grad_norm=1.0,
                           engine_managed=[dit])
    gen = torch.Generator(device=dev)
    batches = []
    for i in range(4):
        gen.manual_seed(1234 + i + 1000 * rank)
        batches.append(torch.rand(B, 3, S, S, device=dev, generator=gen))
    L = _lib.lib()

    from genhancer_b200.graph import GraphedMicroStep
    graphs = []

    def run(max_ctas, budget, deferred=False, bucket_mb=256, steps=6, warm=3, comm=True, graph=True, dyn=1):
        pg = None
        if max_ctas:
            o = dist.ProcessGroupNCCL.Options()
            o.config.max_ctas = max_ctas
            o.config.min_ctas = min(max_ctas, 1)
            pg = dist.new_group(backend="nccl", pg_options=o)
        reducer = GradReducer(groups, engine_modules=[dit], process_group=pg, bucket_cap_bytes=bucket_mb << 20)
        reducer.enabled = comm
        reducer.dynamic_tiles = dyn == 1      # 1: dynamic tile schedule while buckets are in flight; 2: always; 0: never
        K.DYNAMIC_TILES = dyn == 2
        gscale = 1.0 / world
        os.environ["GH_SM_BUDGET"] = str(budget)   # (read once per process: the sweep over budgets needs fresh processes)
        pend = {"n": 0}

        def flush():
            if pend["n"]:
                reducer.wait()
                opt.step(gscale)
                opt.zero_grad()
                pend["n"] = 0

        if graph:
            opt.zero_grad()
            gm = GraphedMicroStep(step, batches[0], prepare=opt.zero_grad, after_backward=reducer.finish)
            graphs.append(gm)

            def one(img):
                gm(img)
                opt.step(gscale)
                opt.zero_grad()
        else:
            def one(img):
                if deferred:
                    loss = step(img, before_trainable=flush)
                    loss.backward()
                    reducer.issue_rest()
                    pend["n"] = 1
                else:
                    loss = step(img)
                    loss.backward()
                    reducer.finish()
                    opt.step(gscale)
                    opt.zero_grad()

        for i in range(warm):
            one(batches[i % 4])
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            one(batches[i % 4])
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        flush()
        torch.cuda.synchronize()
        mine = e0.elapsed_time(e1) / steps
        ms = torch.tensor([mine], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        lo = torch.tensor([mine], device=dev)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        K.DYNAMIC_TILES = False
        dit._on_grads_ready = None
        if graph:           # a graph keeps its own ~35 GB activation pool: release it before the next configuration
            graphs.remove(gm)
            gm.graph.reset()
            del gm, one
            import gc
            gc.collect()
            torch.cuda.empty_cache()
        row = dict(max_ctas=max_ctas, sm_budget=budget, deferred=deferred, bucket_mb=bucket_mb, comm=comm, graph=graph, dyn=dyn,
                   ms_per_step=round(float(ms), 3), ms_fastest_rank=round(float(lo), 3),
                   img_s=round(world * B / float(ms) * 1e3, 1))
        if rank == 0:
            print(json.dumps(row), flush=True)
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", f"dp_sweep_n{world}.jsonl"), "a") as f:
                f.write(json.dumps(row) + "\n")
        return row

    rows = []
    rows.append(run(0, 0, comm=False, dyn=0))          # no exchange at all: the compute floor (max / min over ranks)
    rows.append(run(0, 0, comm=False, dyn=2))          # ... with the dynamic tile schedule everywhere
    rows.append(run(0, 0, dyn=0))                      # NCCL defaults, static schedule (round-1 state so far)
    rows.append(run(0, 0, dyn=1))                      # dynamic schedule while buckets are in flight
    rows.append(run(0, 0, dyn=2))
    quick = os.environ.get("DP_SWEEP_QUICK", "0") == "1"
    if not quick:
        rows.append(run(16, 0, dyn=1))
        rows.append(run(8, 0, dyn=1))
        rows.append(run(0, 0, dyn=1, bucket_mb=64))
    rows.append(run(0, 0, graph=False, deferred=True, dyn=1))
    sys.stdout.flush()
    import threading
    threading.Timer(30.0, lambda: os._exit(0)).start()
    for gm in graphs:       # NCCL will not tear a communicator down while a captured graph holds its kernels
        gm.graph.reset()
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()
    os._exit(0)


if __name__ == "__main__":
    main()
