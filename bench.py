#!/usr/bin/env python
"""bench.py -- throughput of GenHancer's training step on B200s.

Default (`--config img336_stage1`, BASELINE.json configs[1], the configuration the headline metric is quoted on):
OpenAI CLIP ViT-L/14-336 tower (frozen) + projectors + FLUX AE encoder (frozen) + lightweight DiT fwd/bwd +
flow-matching velocity-MSE + grad-clip + AdamW, batch 32 per GPU, bf16 tensor-core math, synthetic 336x336 images,
random-init weights.  One process per GPU (torchrun for N > 1), weak scaling, bucketed NCCL gradient all-reduce
overlapped with the DiT backward.  The other BASELINE configs run through the same code and the same JSON contract:

    --config siglip384_stage2_all    configs[3]: SigLIP-so400m-384, LoRA r16/alpha16 (dropout 0.1, bias lora_only) folded
                                     into the tower GEMMs, tower fwd + dgrad + LoRA wgrads, batch 32 per GPU
    --config use2frames336_stage1    configs[4]: next-frame prediction from 2 conditioning frames (1152 txt tokens), batch 32
    --config sliding336_stage1       configs[4]: sliding windows over 8-frame clips (3 conditioning frames, 1728 txt
                                     tokens, 5 windows per clip), 2 clips = 10 windows per GPU and step

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME]

Prints ONE JSON line (rank 0).  `value` = device-resident inputs; `e2e` = pinned-host inputs copied inside the
timed region + loss read back each step.  The step (update included) is ONE CUDA graph per replay
(genhancer_b200/graph.py: PipelinedTrainStep -- the AdamW update of step n runs on a forked branch under the frozen
forward of step n+1).  `parity_rel` = |loss - oracle loss| / oracle loss of ONE step at the benchmarked shape, the
oracle port running on the same GPU through stock PyTorch ops with the SAME weights, inputs and RNG draws.
Baselines in the same line (N = 1, default config): `cpu_baseline` (the oracle port on the host cores) and
`library_baseline` (the same port through stock PyTorch ops on the GPU, in the reference's dtypes and under bf16
autocast).  `--impl reference` times the oracle port of the reference's step on the host cores (the
reference itself has no GPU-free install here: it needs accelerate/diffusers/peft/omegaconf, none of which are in
the image).
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import sys
import threading
import time
import warnings
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "images/sec (stage-1 step, ViT-L/14-336)"
UNIT = "images/s"
# algorithmic FLOPs per sample: cfg 2 = tower fwd 381.9 + AE enc 472.9 + 3 x DiT 616.2 + 3 x projectors
# 0.043 GF; the others from the same table
CONFIGS = {
    "img336_stage1": dict(
        metric=METRIC, unit=UNIT, flops=2703.6e9, family="OpenAICLIP", size=336, stage2=False, video=None,
        what="OpenAI CLIP ViT-L/14-336 stage-1 step (AE encode + tower + projectors + DiT fwd/bwd + velocity-MSE + clip "
             "+ AdamW)"),
    "siglip384_stage2_all": dict(
        metric="images/sec (stage2_all step, SigLIP-so400m-384, LoRA r16)", unit=UNIT, flops=4428.8e9, family="SigLIP",
        size=384, stage2=True, video=None,
        what="SigLIP-so400m-384 stage2_all step (AE encode + tower fwd/dgrad with LoRA r16/alpha16, dropout 0.1, bias "
             "lora_only folded into the GEMMs + projectors + DiT fwd/bwd + velocity-MSE + clip + AdamW)"),
    "use2frames336_stage1": dict(
        metric="samples/sec (use2frames next-frame-predict stage-1 step, ViT-L/14-336)", unit="samples/s", flops=8455.3e9,
        family="OpenAICLIP", size=336, stage2=False, video=((0, 1), 2),
        what="OpenAI CLIP ViT-L/14-336 use2frames next-frame-prediction stage-1 step (2 tower passes, visual adapter over "
             "1152 patch tokens, AE encode of the target frame, DiT fwd/bwd at L = 441 + 1152, clip + AdamW)"),
    "sliding336_stage1": dict(
        metric="windows/sec (sliding-windows next-frame-predict stage-1 step, ViT-L/14-336)", unit="windows/s",
        flops=11744.5e9, family="OpenAICLIP", size=336, stage2=False, video=((0, 1, 2), 3), sliding=True,
        what="OpenAI CLIP ViT-L/14-336 sliding-windows stage-1 step over 8-frame clips (window gather, 3 tower passes, "
             "visual adapter over 1728 patch tokens, AE encode, DiT fwd/bwd at L = 441 + 1728, clip + AdamW), 5 windows "
             "per clip"),
}


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


def peaks() -> tuple[float, float, str]:
    """(bf16 TFLOP/s for a kernel timed inside a long step, burst TFLOP/s, provenance)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
    return 1400.0, 1590.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------------
# workloads: the step object of one BASELINE config + its synthetic inputs
# ------------------------------------------------------------------------------------------------------------
def build_workload(name: str, dev, B: int):
    """-> namespace(step(*tensors, before_trainable=None) -> loss, trainable [(name, param)], dit, clip_vis, vae, adapter,
    make_inputs(gen) -> tuple of device tensors as a loader yields them, prep(inputs) -> the tensors the step takes,
    samples (per step and GPU), lr, cfg)."""
    from genhancer_b200.clip_models import build_CLIP
    from genhancer_b200.flux.util import load_ae, load_flow_model2
    from genhancer_b200.train_step import (OPENAI_CLIP_MEAN, OPENAI_CLIP_STD, SIGLIP_MEAN, SIGLIP_STD, Stage1ImageStep)

    cfg = CONFIGS[name]
    S = cfg["size"]

    class ClipCfg:
        clip_image_size, clip_dim, t5_dim, clip_type = S, 768, 4096, "large"

    torch.manual_seed(0)  # identical random init on every rank
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        clip_vis = getattr(build_CLIP, f"load_clip_model_{cfg['family']}")(ClipCfg, dev)
        dit = load_flow_model2("flux-dev", device=dev)
        vae = load_ae("flux-dev", device=dev)
    # train_SigLIP_stage1.py:130-141: freeze the AE and the tower, train the projectors (image) / adapter (video) and the DiT
    vae.requires_grad_(False)
    clip_vis.requires_grad_(False)
    mean, std = (SIGLIP_MEAN, SIGLIP_STD) if cfg["family"] == "SigLIP" else (OPENAI_CLIP_MEAN, OPENAI_CLIP_STD)
    if cfg["stage2"]:   # train_SigLIP_stage2_all.py:134-142
        from genhancer_b200.clip_models import lora
        targets = lora.SIGLIP_TARGETS if cfg["family"] == "SigLIP" else "all-linear"
        clip_vis.model = lora.get_peft_model(clip_vis.model, lora.LoraConfig(
            r=16, lora_alpha=16, target_modules=targets, lora_dropout=0.1, bias="lora_only"))
        clip_vis.train()
    dit = dit.to(dev).to(torch.bfloat16)
    dit.train()
    w = SimpleNamespace(name=name, cfg=cfg, dit=dit, clip_vis=clip_vis, vae=vae, adapter=None, mean=mean, std=std,
                        lr=1e-5 if cfg["stage2"] else 1e-4, image_size=S)
    trainable = list(dit.named_parameters())
    if cfg["video"] is None:
        for n, p in clip_vis.named_parameters():
            if "project_clip" in n or "project_t5" in n:
                p.requires_grad = True
        trainable += [(f"clip_vis.{n}", p) for n, p in clip_vis.named_parameters()]
        step = Stage1ImageStep(clip_vis, dit, vae, mean, std, scale_factor=1.0)
        w.step = lambda img, before_trainable=None, **kw: step(img, before_trainable=before_trainable, **kw)
        w.make_inputs = lambda gen: (torch.rand(B, 3, S, S, device=dev, generator=gen),)
        w.prep = lambda inputs: inputs
        w.samples = B
    else:
        from genhancer_b200.video import SuperModel, VideoStep, build_windows_with_mask
        cond_times, target_time = cfg["video"]
        n_cond = len(cond_times)
        sm = SuperModel(clip_vis, dit, adapter_in_dim=clip_vis.model.config.hidden_size,
                        adapter_out_dim=dit.params.context_in_dim).to(dev)
        w.adapter = sm.visual_adapter.float()
        trainable += [(f"visual_adapter.{n}", p) for n, p in w.adapter.named_parameters()]
        step = VideoStep(sm, vae, cond_times=cond_times, target_time=target_time, clip_mean=mean, clip_std=std,
                         scale_factor=1.0)
        w.step = lambda *fr, before_trainable=None, **kw: step(list(fr[:n_cond]), fr[n_cond],
                                                               before_trainable=before_trainable, **kw)
        if cfg.get("sliding"):
            clips, T = max(1, B // 16), 8        # batch 32 -> 2 clips of 8 frames -> 10 windows (SURVEY.md 8d cfg 5)
            w.make_inputs = lambda gen: (torch.rand(clips, T, 3, S, S, device=dev, generator=gen),
                                         torch.ones(clips, T, dtype=torch.bool, device=dev))
            # a7: the mask-aware window gather is part of the step (train_OpenAICLIP_sliding_windows_nextpredic_stage1.py:149-204)
            w.prep = lambda inputs: tuple(build_windows_with_mask(inputs[0], inputs[1], n_cond, 1, 8)[:n_cond + 1])
            w.samples = clips * (T - n_cond)
        else:
            w.make_inputs = lambda gen: tuple(torch.rand(B, 3, S, S, device=dev, generator=gen) for _ in range(n_cond + 1))
            w.prep = lambda inputs: inputs
            w.samples = B
    w.trainable = trainable
    w.raw_step = step
    return w


def build_models(image_size: int, device):
    """(step, clip_vis, dit, vae) of the headline configuration (kept for tools/dp_sweep.py and tools/host_profile.py)."""
    w = build_workload("img336_stage1" if image_size == 336 else "img336_stage1", device, 32)
    return w.raw_step, w.clip_vis, w.dit, w.vae


def oracle_parity(w, inputs, dev) -> dict:
    """ONE step of this workload at the benchmarked shape against the oracle port on the same GPU (stock PyTorch ops;
    fp32 tower + AE, bf16 DiT as the reference runs them, train_SigLIP_stage1.py:127-133,255-261) with the SAME weights,
    inputs and RNG draws.  -> {"parity_rel": ..., "loss": ..., "oracle_loss": ..., cosines of vec / txt / pred}."""
    from oracle import genhancer_oracle as O
    cfg = w.cfg
    S = cfg["size"]
    tc = O.siglip_so400m(S) if cfg["family"] == "SigLIP" else O.openai_vit_l14(S)
    fc, ac = O.FluxCfg(), O.AECfg()
    x = w.prep(inputs)
    Bn = x[-1].shape[0]
    h = S // 8
    gen = torch.Generator(device=dev).manual_seed(4242)
    noise = torch.randn(Bn, 16, h, h, device=dev, generator=gen)
    t = torch.sigmoid(torch.randn(Bn, device=dev, generator=gen))
    x_0 = torch.randn(Bn, (h // 2) ** 2, 64, device=dev, generator=gen)
    was_training = w.clip_vis.training
    w.clip_vis.eval()       # (stage 2: LoRA dropout off for the comparison; B = 0 at init, so the branch adds nothing)
    with torch.no_grad():
        loss, parts = w.step(*x, ae_noise=noise, t=t, x_0=x_0, return_parts=True)
    w.clip_vis.train(was_training)
    ours = {k: parts[k].float().clone() for k in ("vec", "txt", "pred", "x_1")}
    loss = float(loss)
    del parts
    sd_t = {k: v.detach().float() for k, v in w.clip_vis.model.state_dict().items() if "lora_" not in k}
    sd_t = {k.replace(".base_layer.", "."): v for k, v in sd_t.items()}
    sd_d = {k: v.detach() for k, v in w.dit.state_dict().items()}
    sd_a = {k: v.detach().float() for k, v in w.vae.encoder.state_dict().items()}
    with torch.no_grad():
        if cfg["video"] is None:
            sd_w = {k: v.detach().float() for k, v in w.clip_vis.state_dict().items() if k.startswith("project_")}
            out = O.stage1_image_step(sd_t, sd_w, sd_d, sd_a, x[0], tc, fc, ac, w.mean, w.std, noise, t, x_0,
                                      dit_dtype=torch.bfloat16)
        else:
            sd_ad = {k: v.detach().float() for k, v in w.adapter.state_dict().items()}
            out = O.stage1_video_step(sd_t, sd_ad, sd_d, sd_a, list(x[:-1]), x[-1], tc, fc, ac, w.mean, w.std,
                                      cfg["video"][0], cfg["video"][1], noise, t, x_0, dit_dtype=torch.bfloat16)
    cos = lambda a, b: float(torch.nn.functional.cosine_similarity(a.flatten().float(), b.flatten().float(), dim=0))  # noqa: E731
    ref = float(out.loss)
    res = {"parity_rel": round(abs(loss - ref) / abs(ref), 6), "loss": round(loss, 6), "oracle_loss": round(ref, 6),
           "cos_vec": round(cos(ours["vec"], out.vec), 6), "cos_txt": round(cos(ours["txt"], out.txt), 6),
           "cos_pred": round(cos(ours["pred"], out.pred), 6), "cos_x1": round(cos(ours["x_1"], out.x_1), 6),
           "oracle": "oracle/genhancer_oracle.py on the same GPU (stock PyTorch ops, fp32 tower + AE, bf16 DiT), same weights / "
                     f"inputs / draws, batch {Bn} at {S}x{S}"}
    del out, sd_t, sd_d, sd_a, ours
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
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
    B = args.batch
    cfg = CONFIGS[args.config]
    w = build_workload(args.config, dev, B)
    S = w.image_size
    dit = w.dit
    groups = optim.flatten(w.trainable)
    broadcast_parameters(groups)
    opt = optim.FusedAdamW(groups, lr=w.lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, max_grad_norm=1.0,
                           engine_managed=[dit])
    reducer = GradReducer(groups, engine_modules=[dit]) if world > 1 else None
    gscale = reducer.grad_scale if reducer else 1.0

    # synthetic data: per-step seeded batches, a pool resident in HBM / pinned host memory
    pool = 4
    gen = torch.Generator(device=dev)
    dev_batches, host_batches = [], []
    for i in range(pool):
        gen.manual_seed(1234 + i + 1000 * rank)
        x = w.make_inputs(gen)
        dev_batches.append(x)
        host_batches.append(tuple(t.cpu().pin_memory() for t in x))
    h2d_bytes = sum(t.numel() * t.element_size() for t in host_batches[0])

    # ---- parity at the benchmarked shape (one step against the oracle port on this GPU) -------------------------
    parity = None
    if rank == 0 and not args.no_parity:
        try:
            parity = oracle_parity(w, dev_batches[0], dev)
        except Exception as e:  # noqa: BLE001  (report, do not hide: the JSON line says why there is no number)
            parity = {"parity_rel": None, "error": repr(e)[:300]}

    # Data parallel, eager form: the exchange + update of step n run inside step n+1, right before its first trainable
    # kernel (before_trainable) -- the all-reduce tail hides under the frozen AE / tower forward.
    pending = {"n": 0}

    def flush_eager():
        if pending["n"]:
            if reducer is not None:
                reducer.wait()
            opt.step(gscale)
            opt.zero_grad()
            pending["n"] = 0

    def eager_pipelined(*x):
        loss = w.step(*x, before_trainable=flush_eager)
        flush_eager()
        loss.backward()
        if reducer is not None:
            reducer.issue_rest()
        pending["n"] = 1
        return loss

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
    for i in range(min(2, args.warmup)):
        eager_pipelined(*w.prep(dev_batches[i % pool]))
    flush_eager()
    # The whole step -- forward, backward, the overlapped gradient exchange AND the optimizer update -- is ONE CUDA
    # graph per replay (genhancer_b200/graph.py: PipelinedTrainStep): the update of step n sits on a forked branch
    # under the frozen AE / tower forward of step n+1.  Per timed step: one forward, one backward, one complete
    # exchange, one optimizer update.
    pipe = None
    GA = max(1, args.grad_accum)
    if GA > 1 and args.no_graph:
        raise SystemExit("--grad-accum > 1 is measured through the step graph (drop --no-graph)")
    if not args.no_graph:
        from genhancer_b200.graph import PipelinedTrainStep
        pipe = PipelinedTrainStep(w.step, w.prep(dev_batches[0]), opt, reducer, grad_accum=GA)

        def train_step(x):           # one OPTIMIZER step = GA micro-batches (the reference's YAMLs use GA = 2: the gradient
            for _ in range(GA):      # exchange and the update run once per GA micro-steps, train_SigLIP_stage1.py:238-275)
                loss = pipe(*w.prep(x))
            return loss
        flush = pipe.flush
    else:
        train_step = lambda x: eager_pipelined(*w.prep(x))   # noqa: E731
        flush = flush_eager
    for i in range(args.warmup):
        train_step(dev_batches[i % pool])
    clocks = ClockSampler(local)
    n0, r0 = K.LAUNCHES, (pipe.replays if pipe else 0)
    clocks.start()
    ms_value = timed(lambda i: train_step(dev_batches[i % pool]), args.steps)
    host_enqueue_ms = host["enqueue_ms"] / args.steps
    clocks.stop()
    launches = K.LAUNCHES - n0 + ((pipe.replays - r0) * max(pipe.launches_per_replay.values()) if pipe else 0)

    # ---- end-to-end arm: pinned host batch -> H2D -> step -> loss read back, every step ------------------------
    # Input pipeline of the public API as a user would drive it: the pinned batch of step i+1 is copied on a copy
    # stream while step i computes (two device staging buffers), and the loss of step i-1 is read back after step i
    # has been enqueued -- every step still pays its own H2D copy and its own D2H read inside the timed region, they
    # just do not serialise with the compute stream.
    last = {}
    copy_stream = torch.cuda.Stream()
    staged = [tuple(torch.empty_like(t) for t in dev_batches[0]) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]      # H2D of the buffer finished
    consumed = [torch.cuda.Event() for _ in range(2)]   # the step that read the buffer has been enqueued past its read
    loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_done = [torch.cuda.Event() for _ in range(2)]
    e2e_n = {"steps": args.steps}

    def stage(i):
        k = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[k])
            for dst, src in zip(staged[k], host_batches[i % pool]):
                dst.copy_(src, non_blocking=True)
            ready[k].record(copy_stream)

    def e2e_step_ga(i):
        """GA > 1: every micro-batch is its own H2D copy (same stream, no overlap), one loss read per optimizer step."""
        for m in range(GA):
            for dst, src in zip(staged[0], host_batches[(i * GA + m) % pool]):
                dst.copy_(src, non_blocking=True)
            loss = pipe(*w.prep(staged[0]))
        loss_host[0].copy_(loss.detach().float().reshape(()), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        last["loss"] = float(loss_host[0])

    def e2e_step(i):
        if GA > 1:
            return e2e_step_ga(i)
        k = i % 2
        if i == 0:
            stage(0)
        torch.cuda.current_stream().wait_event(ready[k])
        if i + 1 < e2e_n["steps"]:
            stage(i + 1)
        loss = train_step(staged[k])
        consumed[k].record()
        loss_host[k].copy_(loss.detach().float().reshape(()), non_blocking=True)
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
    flush()   # the last step's pending update

    # cross-rank check of the data-parallel arm: after the same number of identical updates every rank must hold the
    # same parameters (sum over |p| of the flat bf16 buffer, compared bit for bit)
    dp_check = None
    if world > 1:
        cs = torch.stack([g.flat_p.float().abs().sum() for g in groups]).double()
        allcs = [torch.zeros_like(cs) for _ in range(world)]
        dist.all_gather(allcs, cs)
        dp_check = {"param_checksum": [float(v) for v in allcs[0]],
                    "ranks_identical": bool(all(torch.equal(allcs[0], c) for c in allcs))}

    # ---- roofline pass: the same step EAGER with CUDA events around every tcgen05 GEMM / conv launch (events cannot
    # sit inside a graph); its own step time is the denominator of share_of_step ---------------------------------
    def eager_step(x):
        loss = w.step(*w.prep(x))
        loss.backward()
        if reducer is not None:
            reducer.finish()
        opt.step(gscale)
        opt.zero_grad()
        return loss

    # (per-launch event timing wants the launches back to back on ONE stream: no side-stream work in this pass)
    from genhancer_b200 import train_step as _ts
    from genhancer_b200.flux import engine as _engine
    _overlap, _engine.GradSink.overlap_wgrad = _engine.GradSink.overlap_wgrad, False
    _overlap_ae, _ts.OVERLAP_AE = _ts.OVERLAP_AE, False
    _overlap_cond, _engine.OVERLAP_COND = getattr(_engine, "OVERLAP_COND", False), False
    timer = GemmTimer()
    n_roof = 0 if args.no_roofline else min(args.steps, 3)
    ms_roof = 0.0
    if n_roof:
        eager_step(dev_batches[0])
        K.GEMM_TIMER = timer
        ms_roof = timed(lambda i: eager_step(dev_batches[i % pool]), n_roof)
        K.GEMM_TIMER = None
    _engine.GradSink.overlap_wgrad = _overlap
    _ts.OVERLAP_AE = _overlap_ae
    _engine.OVERLAP_COND = _overlap_cond
    gemm_flops, gemm_ms, gemm_n = timer.summary()
    if rank == 0 and args.dump_shapes:
        os.makedirs(os.path.dirname(os.path.abspath(args.dump_shapes)), exist_ok=True)
        json.dump({"steps": n_roof, "ms_step": ms_roof / max(n_roof, 1), "rows": timer.by_shape()},
                  open(args.dump_shapes, "w"), indent=1)

    samples = w.samples * world * args.steps * GA
    value = samples / (ms_value / 1e3)
    e2e = samples / (ms_e2e / 1e3)
    peak_sust, peak_burst, prov = peaks()
    flops = cfg["flops"]
    # dram bytes of ONE launch of the dominant kernel from the committed `ncu --set full` capture (profiles/)
    try:
        import glob
        ncu = json.load(open(sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_gemm_ncu_full.json")))[-1]))
    except (OSError, ValueError, IndexError):
        ncu = {}
    out = {
        "metric": cfg["metric"], "value": round(value, 2), "unit": cfg["unit"], "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_value / args.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{cfg['what']}, batch {B}/GPU ({w.samples} samples per step and GPU), random-init weights",
                   "name": args.config, "global_batch": w.samples * world * GA, "gradient_accumulation_steps": GA,
                   "image_size": S, "parallelism": f"dp{world}",
                   "execution": ("forward + backward" + (" + overlapped NCCL gradient all-reduce" if world > 1 else "")
                                 + " + clip/AdamW update replayed as ONE CUDA graph per step (the update of step n on a "
                                   "forked branch under the frozen forward of step n+1)") if pipe else "eager launches",
                   "l2": "per-step working set (weights 3.9 GB + activations) is far larger than the 126 MB L2; "
                         "4 distinct input batches rotate"},
        "mfu": {"flops_per_sample": flops,
                "of_measured_sustained": round(value * flops / world / (peak_sust * 1e12), 4),
                "of_nominal_2250": round(value * flops / world / 2250e12, 4), "peak_source": prov},
        "e2e": {"value": round(e2e, 2), "unit": cfg["unit"], "h2d_bytes_per_step": h2d_bytes * GA, "d2h_bytes_per_step": 4,
                "ms_per_step": round(ms_e2e / args.steps, 3), "last_loss": last.get("loss")},
        "parity_rel": parity.get("parity_rel") if parity else None,
        "parity": parity,
        "gpu_launches": launches,
        "clocks": clocks.summary(),
        "roofline": {"bound": "tensor", "kernel": "umma_gemm_kernel (tcgen05 GEMM / implicit-GEMM conv: all nn.Linear and conv fwd/dgrad/wgrad)",
                     "achieved": round(gemm_flops / (gemm_ms * 1e-3) / 1e12, 1) if gemm_ms > 0 else None,
                     "peak": peak_sust, "unit": "TFLOP/s",
                     "frac": round(gemm_flops / (gemm_ms * 1e-3) / 1e12 / peak_sust, 4) if gemm_ms > 0 else None,
                     "traffic": ncu.get("traffic_bytes_per_launch"),
                     "traffic_of": {k: ncu.get(k) for k in ("capture", "algorithmic_bytes_per_launch", "duration_us",
                                                             "tensor_pipe_active_pct", "sm_clock_ghz_during_capture")},
                     "launches_timed": gemm_n, "peak_source": prov,
                     "share_of_step": round(gemm_ms / ms_roof, 4) if ms_roof > 0 else None,
                     "measured_in": (f"eager single-stream pass of {n_roof} steps, {round(ms_roof / max(n_roof, 1), 3)} ms/step (CUDA "
                                     "events cannot be recorded inside the graphed step)" if n_roof else "skipped (--no-roofline)")},
        "host_enqueue_ms_per_step": round(host_enqueue_ms, 3),
        "cuda_graph": bool(pipe),
    }
    if dp_check is not None:
        out["dp_check"] = dp_check
    default_cfg = args.config == "img336_stage1"
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_step(args.config, budget_s=args.cpu_budget)
    if rank == 0 and world == 1 and default_cfg and not args.no_library_baseline:
        try:
            if pipe is not None:     # give the activation pool of the step graph back first
                pipe.reset()
                pipe = None
            del w, dit, opt, groups, dev_batches, staged, train_step, flush
            import gc
            gc.collect()
            torch.cuda.empty_cache()
            out["library_baseline"] = library_baseline_step(S, B, dev, bf16=False)
            out["library_baseline_bf16"] = library_baseline_step(S, B, dev, bf16=True)
        except Exception as e:  # noqa: BLE001  (a baseline that cannot run must not take the measurement down)
            out.setdefault("library_baseline", {"unavailable": repr(e)[:200]})
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        # NCCL will not tear a communicator down while a captured graph still holds its kernels: drop the graph
        # first; a watchdog ends the process (exit 0, the line is printed) if the teardown stalls anyway.
        sys.stdout.flush()
        threading.Timer(30.0, lambda: os._exit(0)).start()
        if pipe is not None:
            pipe.reset()
            pipe = None
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
        t = base.repeat((n + base.numel() - 1) // base.numel())[:n].reshape(tuple(shp)).clone()
        if "norm" in k and k.endswith(("weight", "scale")):
            t = t + 1.0
        sd[k] = t.requires_grad_(requires_grad)
    return sd


class CpuReferenceStep:
    """The oracle port of ONE sample of a BASELINE config's step (forward + backward, fp32) on the host cores."""

    def __init__(self, config: str, image_size: int | None = None, batch: int = 1):
        from oracle import genhancer_oracle as O
        self.O = O
        self.cfg = cfg = CONFIGS[config]
        S = image_size or cfg["size"]
        self.tc = O.siglip_so400m(S) if cfg["family"] == "SigLIP" else O.openai_vit_l14(S)
        self.fc, self.ac = O.FluxCfg(), O.AECfg()
        self.sd_t = _fast_state_dict(O.tower_key_shapes(self.tc), False)
        self.sd_d = _fast_state_dict(O.flux_key_shapes(self.fc), True)
        self.sd_a = _fast_state_dict(O.ae_encoder_key_shapes(self.ac), False)
        feat = self.tc.feat_dim
        self.lora = None
        if cfg["video"] is None:
            self.sd_w = _fast_state_dict({**O.projector_key_shapes("project_clip", feat, 768),
                                          **O.projector_key_shapes("project_t5", feat, 4096)}, True)
        else:
            self.sd_w = _fast_state_dict(O.adapter_key_shapes(self.tc.hidden, 4096), True)
        if cfg["stage2"]:
            flat = _fast_state_dict(O.lora_key_shapes(self.tc, 16), True)
            names = sorted({k.rsplit(".", 1)[0] for k in flat})
            self.lora = {f"{n}.weight": (flat[f"{n}.lora_A"], flat[f"{n}.lora_B"], 1.0) for n in names}
            self.lora_flat = flat
        self.B, self.S = batch, S

    def __call__(self):
        from genhancer_b200.train_step import OPENAI_CLIP_MEAN, OPENAI_CLIP_STD, SIGLIP_MEAN, SIGLIP_STD
        O, cfg, B, S = self.O, self.cfg, self.B, self.S
        mean, std = (SIGLIP_MEAN, SIGLIP_STD) if cfg["family"] == "SigLIP" else (OPENAI_CLIP_MEAN, OPENAI_CLIP_STD)
        h = S // 8
        noise = torch.randn(B, 16, h, h)
        t = torch.sigmoid(torch.randn(B))
        x_0 = torch.randn(B, (h // 2) ** 2, 64)
        if cfg["video"] is None:
            img = torch.rand(B, 3, S, S)
            out = O.stage1_image_step(self.sd_t, self.sd_w, self.sd_d, self.sd_a, img, self.tc, self.fc, self.ac,
                                      mean, std, noise, t, x_0, lora=self.lora)
        else:
            cond_times, target_time = cfg["video"]
            frames = [torch.rand(B, 3, S, S) for _ in range(len(cond_times) + 1)]
            out = O.stage1_video_step(self.sd_t, self.sd_w, self.sd_d, self.sd_a, frames[:-1], frames[-1], self.tc,
                                      self.fc, self.ac, mean, std, cond_times, target_time, noise, t, x_0)
        out.loss.backward()
        for sd in (self.sd_w, self.sd_d) + ((self.lora_flat,) if self.lora is not None else ()):
            for v in sd.values():
                v.grad = None
        return float(out.loss.detach())


def library_baseline_step(image_size: int, batch: int, device, steps: int = 5, warmup: int = 2, bf16: bool = False) -> dict:
    """SURVEY.md 8(d), optional "library baseline": the same oracle port of the reference step, but with its tensors on
    the B200 -- stock PyTorch ops (cuBLAS / cuDNN / SDPA), forward + backward only (no clip / AdamW).
    ``bf16=False``: fp32 tower + AE and a bf16 DiT, the reference's dtype policy (train_SigLIP_stage1.py:127-133,255-261)
    -- what a user of the reference gets on this GPU.  ``bf16=True``: the same port under ``torch.autocast(bfloat16)``
    (tower + AE on bf16 tensor cores, flash SDPA) -- the fair yard-stick for kernels that run everything in bf16.
    Reported beside the CPU baseline, never part of `value`."""
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
    amp = torch.autocast("cuda", dtype=torch.bfloat16) if bf16 else contextlib.nullcontext()

    def one():
        img = torch.rand(batch, 3, image_size, image_size, device=device, generator=gen)
        noise = torch.randn(batch, 16, h, h, device=device, generator=gen)
        t = torch.sigmoid(torch.randn(batch, device=device, generator=gen))
        x_0 = torch.randn(batch, (h // 2) ** 2, 64, device=device, generator=gen)
        with amp:
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
    policy = "torch.autocast(bfloat16): tower + AE + DiT on bf16 tensor cores, flash SDPA" if bf16 else \
        "fp32 tower + AE, bf16 DiT, SDPA (the reference's dtype policy)"
    return {"value": round(batch / ms * 1e3, 2), "unit": UNIT, "ms_per_step": round(ms, 2), "kind": "port-on-gpu",
            "what": f"oracle port of the reference step through stock PyTorch {torch.__version__} ops on the same GPU "
                    f"({policy}), batch {batch}, forward + backward only, {steps} steps"}


def cpu_baseline_step(config: str, budget_s: float = 45.0, target_s: float = 15.0) -> dict:
    """The oracle port of the reference step on the host cores: one warm-up step, then whole steps until about
    ``target_s`` seconds of CPU work are done (at least 3, never past ``budget_s``)."""
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    stepper = CpuReferenceStep(config)
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
    return {"value": round(n / dt, 5), "unit": CONFIGS[config]["unit"], "cores": cores, "kind": "port",
            "sample": f"{n} steps of 1 sample at {stepper.S}x{stepper.S} (fwd+bwd, fp32) through oracle/genhancer_oracle.py "
                      f"after one warm-up step ({cold:.1f} s cold), {dt:.1f} s of CPU work"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    cfg = CONFIGS[args.config]
    S = cfg["size"]
    stepper = CpuReferenceStep(args.config)
    t0 = time.perf_counter()
    stepper()
    first = time.perf_counter() - t0
    sample = f"1 sample at {S}x{S} per step (fwd+bwd fp32, oracle port of the reference step)"
    if first * (args.warmup + args.steps) > args.ref_budget and S > 224:
        # keep the whole run within a few minutes on slow hosts: fall back to BASELINE.json configs[0]'s image size
        stepper = CpuReferenceStep(args.config, image_size=224)
        sample = (f"1 sample at 224x224 per step ({S} projected to take {first * (args.warmup + args.steps):.0f} s "
                  f"> {args.ref_budget:.0f} s budget); fwd+bwd fp32, oracle port of the reference step")
        stepper()
    for _ in range(max(0, args.warmup - 1)):
        stepper()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        stepper()
    dt = time.perf_counter() - t0
    value = args.steps / dt
    out = {"impl": "reference", "metric": cfg["metric"], "value": round(value, 5), "unit": cfg["unit"], "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 1),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"{cfg['what']}, batch {args.batch}/GPU, random-init weights", "name": args.config,
                      "global_batch": args.batch * args.gpus, "image_size": cfg["size"], "parallelism": f"dp{args.gpus}",
                      "execution": f"reference arm: oracle port on {cores} host threads, bounded sample ({sample})"},
           "cpu_baseline": {"value": round(value, 5), "unit": cfg["unit"], "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": round(value, 5), "unit": cfg["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="img336_stage1", choices=sorted(CONFIGS),
                    help="which BASELINE.json configuration (default: configs[1], the one the headline metric is quoted on)")
    ap.add_argument("--batch", type=int, default=32, help="per-GPU batch (BASELINE configs: 32)")
    ap.add_argument("--grad-accum", type=int, default=1,
                    help="micro-batches per optimizer step (SURVEY.md 8d cfg 3 asks for GA = 1 and the reference's GA = 2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true",
                    help="skip the stock-PyTorch-on-this-GPU runs of the oracle port (SURVEY.md 8d, optional baseline)")
    ap.add_argument("--no-parity", action="store_true", help="skip the one-step comparison with the oracle port on this GPU")
    ap.add_argument("--no-roofline", action="store_true", help="skip the eager per-launch GEMM timing pass (multi-GPU sweeps)")
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
