"""Step time of BASELINE.json configs[3] and configs[4] on ONE B200 (bench.py measures configs[1], the headline):

    cfg 4   SigLIP-so400m-384 stage2_all, LoRA r16/alpha16 (dropout 0.1, bias lora_only) folded into the GEMMs, batch 32
    cfg 5a  OpenAI CLIP-336 use2frames next-frame prediction, stage 1, batch 32 (two conditioning frames, 1152 txt tokens)
    cfg 5b  OpenAI CLIP-336 sliding windows (3 conditioning frames, 1728 txt tokens), stage 1, 8 windows per step

Synthetic inputs, random-init weights, CUDA-event timing over `--steps` steps after `--warmup`; FLOPs per sample from
SURVEY.md 8(d).  One JSON line per config (also appended to gpurun_out/configs.jsonl)."""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import sys
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from genhancer_b200 import optim
from genhancer_b200.clip_models import build_CLIP, lora
from genhancer_b200.flux.util import load_ae, load_flow_model2
from genhancer_b200.train_step import OPENAI_CLIP_MEAN, OPENAI_CLIP_STD, SIGLIP_MEAN, SIGLIP_STD, Stage1ImageStep
from genhancer_b200.video import SuperModel, VideoStep

GF = {"siglip384_stage2_all": 4428.8e9, "use2frames336_stage1": 8455.3e9, "sliding336_stage1": 11744.5e9}


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--only", default="")
    ap.add_argument("--profile", action="store_true", help="print the kernel-time table of one step (torch.profiler / CUPTI)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    peak = 1391.6
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
    except (OSError, KeyError, ValueError):
        pass
    os.makedirs("gpurun_out", exist_ok=True)

    def kernel_table(fn, name):
        from torch.profiler import ProfilerActivity, profile
        fn()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
        tot = sum(e.device_time_total for e in rows)
        lines = [f"# {name}: kernel time of one step {tot / 1e3:.2f} ms (CUPTI, serial sum)"]
        for e in rows[:28]:
            lines.append(f"{100 * e.device_time_total / tot:6.2f}% {e.device_time_total / 1e3:9.3f} ms x{e.count:4d}  {e.key[:110]}")
        print("\n".join(lines), flush=True)
        open(f"gpurun_out/kernels_{name}.txt", "w").write("\n".join(lines) + "\n")

    def report(name, ms, samples, unit):
        out = {"config": name, "ms_per_step": round(ms, 2), "value": round(samples / ms * 1e3, 2), "unit": unit,
               "samples_per_step": samples, "flops_per_sample": GF[name],
               "mfu_of_measured_sustained": round(samples / ms * 1e3 * GF[name] / (peak * 1e12), 4),
               "mfu_of_nominal_2250": round(samples / ms * 1e3 * GF[name] / 2250e12, 4), "n_gpus": 1, "steps": args.steps,
               "warmup": args.warmup, "data": "synthetic", "dtype": "bf16"}
        print(json.dumps(out), flush=True)
        open("gpurun_out/configs.jsonl", "a").write(json.dumps(out) + "\n")

    def models(family, size, clip_dim=768):
        class C:
            clip_image_size, t5_dim, clip_type = size, 4096, "large"
        C.clip_dim = clip_dim
        torch.manual_seed(0)
        with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
            warnings.simplefilter("ignore")
            clip_vis = getattr(build_CLIP, f"load_clip_model_{family}")(C, dev)
            dit = load_flow_model2("flux-dev", device=dev).to(dev).to(torch.bfloat16)
            vae = load_ae("flux-dev", device=dev)
        vae.requires_grad_(False)
        clip_vis.requires_grad_(False)
        dit.train()
        return clip_vis, dit, vae

    want = lambda k: not args.only or k in args.only.split(",")
    if want("siglip384_stage2_all"):
        clip_vis, dit, vae = models("SigLIP", 384)
        clip_vis.model = lora.get_peft_model(clip_vis.model, lora.LoraConfig(r=16, lora_alpha=16, target_modules=lora.SIGLIP_TARGETS,
                                                                             lora_dropout=0.1, bias="lora_only"))
        clip_vis.train()
        for n, p in clip_vis.named_parameters():
            if "project_clip" in n or "project_t5" in n:
                p.requires_grad = True
        step = Stage1ImageStep(clip_vis, dit, vae, SIGLIP_MEAN, SIGLIP_STD)
        groups = optim.flatten(list(dit.named_parameters()) + [(f"clip_vis.{n}", p) for n, p in clip_vis.named_parameters()])
        opt = optim.FusedAdamW(groups, lr=1e-5, engine_managed=[dit])
        x = torch.rand(32, 3, 384, 384, device=dev)

        def fn():
            step(x).backward()
            opt.step()
            opt.zero_grad()
        report("siglip384_stage2_all", timed(fn, args.steps, args.warmup), 32, "images/s")
        if args.profile:
            kernel_table(fn, "siglip384_stage2_all")
        del clip_vis, dit, vae, step, groups, opt, x
        torch.cuda.empty_cache()
    for name, times, n_cond, bs in (("use2frames336_stage1", ((0, 1), 2), 2, 32), ("sliding336_stage1", ((0, 1, 2), 3), 3, 8)):
        if not want(name):
            continue
        clip_vis, dit, vae = models("OpenAICLIP", 336)
        sm = SuperModel(clip_vis, dit, adapter_in_dim=1024, adapter_out_dim=4096).to(dev)
        sm.visual_adapter.float()
        step = VideoStep(sm, vae, cond_times=times[0], target_time=times[1], clip_mean=OPENAI_CLIP_MEAN, clip_std=OPENAI_CLIP_STD)
        groups = optim.flatten(list(dit.named_parameters()) + [(f"visual_adapter.{n}", p) for n, p in sm.visual_adapter.named_parameters()])
        opt = optim.FusedAdamW(groups, lr=1e-4, engine_managed=[dit])
        frames = [torch.rand(bs, 3, 336, 336, device=dev) for _ in range(n_cond + 1)]

        def fn():
            step(frames[:n_cond], frames[n_cond]).backward()
            opt.step()
            opt.zero_grad()
        report(name, timed(fn, args.steps, args.warmup), bs, "samples/s" if n_cond == 2 else "windows/s")
        if args.profile:
            kernel_table(fn, name)
        del clip_vis, dit, vae, sm, step, groups, opt, frames
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
