#!/usr/bin/env python
"""bench.py -- images/sec of GenHancer's stage-1 training step (BASELINE.json configs[1]):
OpenAI CLIP ViT-L/14-336 tower (frozen) + projectors + FLUX AE encoder (frozen) + lightweight DiT fwd/bwd +
flow-matching velocity-MSE + grad-clip + AdamW, batch 32 per GPU, bf16 tensor-core math, synthetic 336x336 images,
random-init weights.  One process per GPU (torchrun for N > 1), weak scaling, bucketed NCCL gradient all-reduce
overlapped with the DiT backward.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `value` = device-resident inputs; `e2e` = pinned-host inputs copied inside the
timed region + loss read back each step.  Baselines in the same line (N = 1): `cpu_baseline` (the oracle port on the
host cores) and `library_baseline` (the same port through stock PyTorch ops on the GPU).  `--impl reference` times the oracle port of the reference's step on
the host cores (the reference itself has no GPU-free install here: it needs accelerate/diffusers/peft/omegaconf,
none of which are in the image).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "images/sec (stage-1 step, ViT-L/14-336)"
UNIT = "images/s"
# algorithmic FLOPs per image of the cfg-2 step: tower fwd 381.9 + AE enc 472.9 + 3 x DiT 616.2
# + 3 x projectors 0.043 GF
FLOPS_PER_IMAGE = 2703.6e9


# ------------------------------------------------------------------------------------------------------------
# clocks sampler (NVML) -- runs DURING the timed region
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index: int, period_s: float = 0.1):
        self.samples, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self._stop = threading.Event()
        self._t = None
        self.period = period_s
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.nv, self.err = None, repr(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(self.period)

    def start(self):
        if self.nv is not None:
            self._stop.clear()
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        if self._t is not None:
            self._stop.set()
            self._t.join()
            self._t = None

    def summary(self) -> dict:
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "power_w_max": round(max(self.power), 1) if self.power else None}


# ------------------------------------------------------------------------------------------------------------
# per-launch GEMM timing (roofline of the dominant kernel), CUDA events on the launching stream
# ------------------------------------------------------------------------------------------------------------
class GemmTimer:
    def __init__(self):
        self.records = []  # (start, stop, flops)

    def __call__(self, flops: float, tag: str = ""):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.records.append((s, e, flops, tag))
        return s, e

    def summary(self) -> tuple[float, float, int]:
        ms = sum(s.elapsed_time(e) for s, e, _, _ in self.records)
        fl = sum(f for _, _, f, _ in self.records)
        return fl, ms, len(self.records)

    def by_shape(self) -> list[dict]:
        agg: dict[str, list] = {}
        for s, e, f, tag in self.records:
            a = agg.setdefault(tag, [0, 0.0, 0.0])
            a[0] += 1
            a[1] += s.elapsed_time(e)
            a[2] += f
        rows = [dict(shape=k, launches=v[0], ms_total=round(v[1], 3), tflops=round(v[2] / (v[1] * 1e-3) / 1e12, 1) if v[1] > 0 else None)
                for k, v in agg.items()]
        return sorted(rows, key=lambda r: -r["ms_total"])


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def peaks() -> tuple[float, float, str]:
    """(bf16 TFLOP/s for a kernel timed inside a long step, burst TFLOP/s, provenance)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
    return 1400.0, 1590.0, "fallback (B200_PROFILING.md)"


def build_models(image_size: int, device):
    from genhancer_b200.clip_models.build_CLIP import load_clip_model_OpenAICLIP
    from genhancer_b200.flux.util import load_ae, load_flow_model2
    from genhancer_b200.train_step import Stage1ImageStep

    class ClipCfg:
        clip_image_size, clip_dim, t5_dim = image_size, 768, 4096

    import contextlib
    import io
    import warnings
    torch.manual_seed(0)  # identical random init on every rank
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        clip_vis = load_clip_model_OpenAICLIP(ClipCfg, device)
        dit = load_flow_model2("flux-dev", device=device)
        vae = load_ae("flux-dev", device=device)
    # train_SigLIP_stage1.py:130-141: freeze the AE and the tower, train the projectors and the bf16 DiT
    vae.requires_grad_(False)
    clip_vis.requires_grad_(False)
    for n, p in clip_vis.named_parameters():
        if "project_clip" in n or "project_t5" in n:
            p.requires_grad = True
    dit = dit.to(device).to(torch.bfloat16)
    dit.train()
    step = Stage1ImageStep(clip_vis, dit, vae, scale_factor=1.0)
    return step, clip_vis, dit, vae


def run_ours(args):
    import torch.distributed as dist

    from genhancer_b200 import kernels as K, optim
    from genhancer_b200.parallel import GradReducer, broadcast_parameters

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, S = args.batch, args.image_size
    step, clip_vis, dit, vae = build_models(S, dev)
    trainable = list(dit.named_parameters()) + [(f"clip_vis.{n}", p) for n, p in clip_vis.named_parameters()]
    groups = optim.flatten(trainable)
    broadcast_parameters(groups)
    opt = optim.FusedAdamW(groups, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, max_grad_norm=1.0,
                           engine_managed=[dit])
    reducer = GradReducer(groups, engine_modules=[dit]) if world > 1 else None
    gscale = reducer.grad_scale if reducer else 1.0

    # Data parallel: the exchange + update of step n run inside step n+1, right before its first trainable kernel
    # (train_step.Stage1ImageStep.__call__: before_trainable) -- the all-reduce tail hides under the frozen AE / tower
    # forward.  Per timed step: one forward, one backward, one complete exchange, one optimizer update.
    pending = {"n": 0}

    def flush():
        if pending["n"]:
            reducer.wait()
            opt.step(gscale)
            opt.zero_grad()
            pending["n"] = 0

    def train_step(img):
        if reducer is None:
            loss = step(img)
            loss.backward()
            opt.step(gscale)
            opt.zero_grad()
            return loss
        loss = step(img, before_trainable=flush)
        loss.backward()
        reducer.issue_rest()        # the fp32 projector group: issued now, waited for in the next step's flush()
        pending["n"] = 1
        return loss

    # synthetic data: per-step seeded batches (SURVEY.md 8d cfg 2), a pool resident in HBM / pinned host memory
    pool = 4
    gen = torch.Generator(device=dev)
    dev_batches, host_batches = [], []
    for i in range(pool):
        gen.manual_seed(1234 + i + 1000 * rank)
        x = torch.rand(B, 3, S, S, device=dev, generator=gen)
        dev_batches.append(x)
        host_batches.append(x.cpu().pin_memory())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host = {}

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        host["enqueue_ms"] = (time.perf_counter() - t0) * 1e3   # host time to ENQUEUE the steps (no sync inside)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident arm -------------------------------------------------------------------------------
    for i in range(args.warmup):
        train_step(dev_batches[i % pool])
    # forward + backward (+ the overlapped gradient exchange) as ONE CUDA graph (genhancer_b200/graph.py); optimizer
    # + zero_grad stay outside.  Data parallel: the NCCL all-reduces the backward schedule issues per finished block
    # are captured as a forked branch of the graph, joined by reducer.finish() at its end.
    graphed = None
    if not args.no_graph:
        from genhancer_b200.graph import GraphedMicroStep
        flush()
        opt.zero_grad()
        graphed = GraphedMicroStep(step, dev_batches[0], prepare=opt.zero_grad,
                                   after_backward=reducer.finish if reducer is not None else None)

        def train_step(img):  # noqa: F811  (same contract as the eager step above)
            loss = graphed(img)
            opt.step(gscale)
            opt.zero_grad()
            return loss

        for i in range(args.warmup):
            train_step(dev_batches[i % pool])
    clocks = ClockSampler(local)
    n0, r0 = K.LAUNCHES, (graphed.replays if graphed else 0)
    clocks.start()
    ms_value = timed(lambda i: train_step(dev_batches[i % pool]), args.steps)
    host_enqueue_ms = host["enqueue_ms"] / args.steps
    clocks.stop()
    launches = K.LAUNCHES - n0 + ((graphed.replays - r0) * graphed.launches_per_replay if graphed else 0)
    # ---- end-to-end arm: pinned host batch -> H2D -> step -> loss read back, every step ------------------------
    last = {}

    # Input pipeline of the public API as a user would drive it: the pinned batch of step i+1 is copied on a copy
    # stream while step i computes (two device staging buffers), and the loss of step i-1 is read back after step i
    # has been enqueued -- every step still pays its own H2D copy and its own D2H read inside the timed region, they
    # just do not serialise with the compute stream.
    copy_stream = torch.cuda.Stream()
    staged = [torch.empty_like(dev_batches[0]) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]      # H2D of the buffer finished
    consumed = [torch.cuda.Event() for _ in range(2)]   # the step that read the buffer has been enqueued past its read
    loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_done = [torch.cuda.Event() for _ in range(2)]
    e2e_n = {"steps": args.steps}

    def stage(i):
        k = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[k])
            staged[k].copy_(host_batches[i % pool], non_blocking=True)
            ready[k].record(copy_stream)

    def e2e_step(i):
        k = i % 2
        if i == 0:
            stage(0)
        torch.cuda.current_stream().wait_event(ready[k])
        if i + 1 < e2e_n["steps"]:
            stage(i + 1)
        loss = train_step(staged[k])
        consumed[k].record()
        loss_host[k].copy_(loss.detach().float(), non_blocking=True)
        loss_done[k].record()
        if i > 0:
            loss_done[1 - k].synchronize()
            last["loss"] = float(loss_host[1 - k])
        if i + 1 == e2e_n["steps"]:
            loss_done[k].synchronize()
            last["loss"] = float(loss_host[k])

    for ev in consumed:
        ev.record()
    e2e_n["steps"] = min(2, args.warmup)
    for i in range(e2e_n["steps"]):
        e2e_step(i)
    torch.cuda.synchronize()
    e2e_n["steps"] = args.steps
    ms_e2e = timed(e2e_step, args.steps)

    # ---- roofline pass: the same step EAGER with CUDA events around every tcgen05 GEMM / conv launch (events cannot
    # sit inside a graph); its own step time is the denominator of share_of_step ---------------------------------
    flush()   # (data parallel: the last timed step's deferred exchange + update)

    def eager_step(img):
        loss = step(img)
        loss.backward()
        if reducer is not None:
            reducer.finish()
        opt.step(gscale)
        opt.zero_grad()
        return loss

    # (per-launch event timing wants the launches back to back on ONE stream: no side-stream wgrads in this pass)
    from genhancer_b200 import train_step as _ts
    from genhancer_b200.flux import engine as _engine
    _overlap, _engine.GradSink.overlap_wgrad = _engine.GradSink.overlap_wgrad, False
    _overlap_ae, _ts.OVERLAP_AE = _ts.OVERLAP_AE, False
    eager_step(dev_batches[0])
    timer = GemmTimer()
    K.GEMM_TIMER = timer
    n_roof = min(args.steps, 3)
    ms_roof = timed(lambda i: eager_step(dev_batches[i % pool]), n_roof)
    K.GEMM_TIMER = None
    _engine.GradSink.overlap_wgrad = _overlap
    _ts.OVERLAP_AE = _overlap_ae
    gemm_flops, gemm_ms, gemm_n = timer.summary()
    if rank == 0 and args.dump_shapes:
        os.makedirs(os.path.dirname(os.path.abspath(args.dump_shapes)), exist_ok=True)
        json.dump({"steps": n_roof, "ms_step": ms_roof / n_roof, "rows": timer.by_shape()},
                  open(args.dump_shapes, "w"), indent=1)

    images = B * world * args.steps
    value = images / (ms_value / 1e3)
    e2e = images / (ms_e2e / 1e3)
    peak_sust, peak_burst, prov = peaks()
    # dram bytes of ONE launch of the dominant kernel from the committed `ncu --set full` capture (profiles/)
    try:
        import glob
        ncu = json.load(open(sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_gemm_ncu_full.json")))[-1]))
    except (OSError, ValueError):
        ncu = {}
    out = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_value / args.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"OpenAI CLIP ViT-L/14-{S} stage-1 step (AE encode + tower + projectors + DiT fwd/bwd + "
                               f"velocity-MSE + clip + AdamW), batch {B}/GPU, random-init weights",
                   "global_batch": B * world, "image_size": S, "parallelism": f"dp{world}",
                   "execution": ("forward+backward" + ("+overlapped NCCL gradient all-reduce" if world > 1 else "")
                                 + " replayed as one CUDA graph, fused AdamW outside") if graphed else "eager launches",
                   "l2": "per-step working set (weights 3.9 GB + activations) is far larger than the 126 MB L2; "
                         "4 distinct input batches rotate"},
        "mfu": {"flops_per_image": FLOPS_PER_IMAGE,
                "of_measured_sustained": round(value * FLOPS_PER_IMAGE / world / (peak_sust * 1e12), 4),
                "of_nominal_2250": round(value * FLOPS_PER_IMAGE / world / 2250e12, 4), "peak_source": prov},
        "e2e": {"value": round(e2e, 2), "unit": UNIT, "h2d_bytes_per_step": B * 3 * S * S * 4, "d2h_bytes_per_step": 4,
                "ms_per_step": round(ms_e2e / args.steps, 3), "last_loss": last.get("loss")},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
        "roofline": {"bound": "tensor", "kernel": "umma_gemm_kernel (tcgen05 GEMM, all nn.Linear fwd/dgrad/wgrad)",
                     "achieved": round(gemm_flops / (gemm_ms * 1e-3) / 1e12, 1) if gemm_ms > 0 else None,
                     "peak": peak_sust, "unit": "TFLOP/s",
                     "frac": round(gemm_flops / (gemm_ms * 1e-3) / 1e12 / peak_sust, 4) if gemm_ms > 0 else None,
                     "traffic": ncu.get("traffic_bytes_per_launch"),
                     "traffic_of": {k: ncu.get(k) for k in ("capture", "algorithmic_bytes_per_launch", "duration_us",
                                                             "tensor_pipe_active_pct", "sm_clock_ghz_during_capture")},
                     "launches_timed": gemm_n, "peak_source": prov,
                     "share_of_step": round(gemm_ms / ms_roof, 4) if ms_roof > 0 else None,
                     "measured_in": f"eager pass of {n_roof} steps, {round(ms_roof / n_roof, 3)} ms/step (CUDA events cannot be "
                                    "recorded inside the graphed step)" if graphed else "the timed steps"},
        "host_enqueue_ms_per_step": round(host_enqueue_ms, 3),
        "cuda_graph": bool(graphed),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_step(S, budget_s=args.cpu_budget)
    if rank == 0 and world == 1 and not args.no_library_baseline:
        try:
            if graphed is not None:     # give the activation pool of the step graph back first
                graphed.graph.reset()
                graphed = None
            del step, clip_vis, dit, vae, opt, groups, trainable, dev_batches, staged
            import gc
            gc.collect()
            torch.cuda.empty_cache()
            out["library_baseline"] = library_baseline_step(S, B, dev)
        except Exception as e:  # noqa: BLE001  (a baseline that cannot run must not take the measurement down)
            out["library_baseline"] = {"unavailable": repr(e)[:200]}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        # NCCL will not tear a communicator down while a captured graph still holds its kernels: drop the graph
        # first; a watchdog ends the process (exit 0, the line is printed) if the teardown stalls anyway.
        sys.stdout.flush()
        threading.Timer(30.0, lambda: os._exit(0)).start()
        if graphed is not None:
            graphed.graph.reset()
            graphed = None
        barrier()
        dist.destroy_process_group()
        os._exit(0)


# ------------------------------------------------------------------------------------------------------------
# CPU legs: the oracle port of the reference's step (oracle/genhancer_oracle.py), host cores only
# ------------------------------------------------------------------------------------------------------------
def _fast_state_dict(key_shapes: dict, requires_grad: bool) -> dict:
    """Timing-only weights: one random vector tiled into every tensor (RNG-ing 1.6 B values would dominate)."""
    base = torch.randn(1 << 20) * 0.02
    sd = {}
    for k, shp in key_shapes.items():
        n = 1
        for s in shp:
            n *= s
        t = base.repeat((n + base.numel() - 1) // base.numel())[:n].reshape(shp).clone()
        if "norm" in k and k.endswith(("weight", "scale")):
            t = t + 1.0
        sd[k] = t.requires_grad_(requires_grad)
    return sd


class CpuReferenceStep:
    def __init__(self, image_size: int, batch: int):
        from oracle import genhancer_oracle as O
        self.O = O
        self.tc, self.fc, self.ac = O.openai_vit_l14(image_size), O.FluxCfg(), O.AECfg()
        self.sd_t = _fast_state_dict(O.tower_key_shapes(self.tc), False)
        self.sd_w = _fast_state_dict({**O.projector_key_shapes("project_clip", 768, 768),
                                      **O.projector_key_shapes("project_t5", 768, 4096)}, True)
        self.sd_d = _fast_state_dict(O.flux_key_shapes(self.fc), True)
        self.sd_a = _fast_state_dict(O.ae_encoder_key_shapes(self.ac), False)
        self.B, self.S = batch, image_size

    def __call__(self):
        from genhancer_b200.train_step import OPENAI_CLIP_MEAN, OPENAI_CLIP_STD
        B, S = self.B, self.S
        h = S // 8
        img = torch.rand(B, 3, S, S)
        noise = torch.randn(B, 16, h, h)
        t = torch.sigmoid(torch.randn(B))
        x_0 = torch.randn(B, (h // 2) ** 2, 64)
        out = self.O.stage1_image_step(self.sd_t, self.sd_w, self.sd_d, self.sd_a, img, self.tc, self.fc, self.ac,
                                       OPENAI_CLIP_MEAN, OPENAI_CLIP_STD, noise, t, x_0)
        out.loss.backward()
        for sd in (self.sd_w, self.sd_d):
            for v in sd.values():
                v.grad = None
        return float(out.loss.detach())


def library_baseline_step(image_size: int, batch: int, device, steps: int = 5, warmup: int = 2) -> dict:
    """SURVEY.md 8(d), optional "library baseline": the same oracle port of the reference step, but with its tensors on
    the B200 -- stock PyTorch ops (cuBLAS / cuDNN / SDPA), fp32 tower + AE and a bf16 DiT as the reference runs them
    (train_SigLIP_stage1.py:127-133,255-261), forward + backward only (no clip / AdamW).  What a user of the reference
    gets on this GPU without our kernels; reported beside the CPU baseline, never part of `value`."""
    from genhancer_b200.train_step import OPENAI_CLIP_MEAN, OPENAI_CLIP_STD
    from oracle import genhancer_oracle as O
    tc, fc, ac = O.openai_vit_l14(image_size), O.FluxCfg(), O.AECfg()
    to = lambda sd, rg: {k: v.detach().to(device).requires_grad_(rg) for k, v in sd.items()}  # noqa: E731
    sd_t = to(_fast_state_dict(O.tower_key_shapes(tc), False), False)
    sd_w = to(_fast_state_dict({**O.projector_key_shapes("project_clip", 768, 768),
                                **O.projector_key_shapes("project_t5", 768, 4096)}, False), True)
    sd_d = {k: v.to(torch.bfloat16).requires_grad_(True)
            for k, v in to(_fast_state_dict(O.flux_key_shapes(fc), False), False).items()}
    sd_a = to(_fast_state_dict(O.ae_encoder_key_shapes(ac), False), False)
    h = image_size // 8
    gen = torch.Generator(device=device).manual_seed(7)

    def one():
        img = torch.rand(batch, 3, image_size, image_size, device=device, generator=gen)
        noise = torch.randn(batch, 16, h, h, device=device, generator=gen)
        t = torch.sigmoid(torch.randn(batch, device=device, generator=gen))
        x_0 = torch.randn(batch, (h // 2) ** 2, 64, device=device, generator=gen)
        out = O.stage1_image_step(sd_t, sd_w, sd_d, sd_a, img, tc, fc, ac, OPENAI_CLIP_MEAN, OPENAI_CLIP_STD, noise, t,
                                  x_0, dit_dtype=torch.bfloat16)
        out.loss.backward()
        for sd in (sd_w, sd_d):
            for v in sd.values():
                v.grad = None

    for _ in range(warmup):
        one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del sd_t, sd_w, sd_d, sd_a
    torch.cuda.empty_cache()
    return {"value": round(batch / ms * 1e3, 2), "unit": UNIT, "ms_per_step": round(ms, 2), "kind": "port-on-gpu",
            "what": f"oracle port of the reference step through stock PyTorch {torch.__version__} ops on the same GPU "
                    f"(fp32 tower + AE, bf16 DiT, SDPA), batch {batch}, forward + backward only, {steps} steps"}


def cpu_baseline_step(image_size: int, budget_s: float = 45.0, target_s: float = 15.0) -> dict:
    """The oracle port of the reference step on the host cores: one warm-up step, then whole steps until about
    ``target_s`` seconds of CPU work are done (at least 3, never past ``budget_s``)."""
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    stepper = CpuReferenceStep(image_size, 1)
    t0 = time.perf_counter()
    stepper()
    cold = time.perf_counter() - t0
    n, t0 = 0, time.perf_counter()
    while True:
        stepper()
        n += 1
        dt = time.perf_counter() - t0
        if (n >= 3 and dt >= target_s) or dt + cold > budget_s:
            break
    return {"value": round(n / dt, 5), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} steps of 1 image at {image_size}x{image_size} (fwd+bwd, fp32) through oracle/genhancer_oracle.py "
                      f"after one warm-up step ({cold:.1f} s cold), {dt:.1f} s of CPU work"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    S = args.image_size
    stepper = CpuReferenceStep(S, 1)
    t0 = time.perf_counter()
    stepper()
    first = time.perf_counter() - t0
    sample = f"1 image at {S}x{S} per step (fwd+bwd fp32, oracle port of the reference step)"
    if first * (args.warmup + args.steps) > args.ref_budget and S > 224:
        # keep the whole run within a few minutes on slow hosts: fall back to BASELINE.json configs[0]'s image size
        S = 224
        stepper = CpuReferenceStep(S, 1)
        sample = (f"1 image at 224x224 per step (336 projected to take {first * (args.warmup + args.steps):.0f} s "
                  f"> {args.ref_budget:.0f} s budget); fwd+bwd fp32, oracle port of the reference step")
        stepper()
    for _ in range(max(0, args.warmup - 1)):
        stepper()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        stepper()
    dt = time.perf_counter() - t0
    value = args.steps / dt
    out = {"impl": "reference", "metric": METRIC, "value": round(value, 5), "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 1),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"OpenAI CLIP ViT-L/14-{args.image_size} stage-1 step (AE encode + tower + projectors + DiT fwd/bwd + "
                                  f"velocity-MSE + clip + AdamW), batch {args.batch}/GPU, random-init weights",
                      "global_batch": args.batch * args.gpus, "image_size": args.image_size, "parallelism": f"dp{args.gpus}",
                      "execution": f"reference arm: oracle port on {cores} host threads, bounded sample ({sample})"},
           "cpu_baseline": {"value": round(value, 5), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": round(value, 5), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="per-GPU batch (BASELINE configs[1]: 32)")
    ap.add_argument("--image-size", type=int, default=336)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true",
                    help="skip the stock-PyTorch-on-this-GPU run of the oracle port (SURVEY.md 8d, optional baseline)")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the whole-step CUDA graph")
    ap.add_argument("--dump-shapes", default="", help="write the per-shape GEMM/conv timing table (JSON) here")
    ap.add_argument("--cpu-budget", type=float, default=45.0)
    ap.add_argument("--ref-budget", type=float, default=240.0)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        print(f"note: --warmup {args.warmup} < 3 breaks the timing rules; numbers are for debugging only", file=sys.stderr)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
