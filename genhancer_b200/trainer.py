"""Host of the ``train_*.py`` entry points: ``python train_X.py --config <yaml>``.

Keeps the reference's contract -- the single ``--config`` flag, the YAML schema, the loop
semantics (gradient accumulation, clip to ``max_grad_norm`` on sync steps, AdamW, constant LR) and the flat
checkpoint layout in ``output_dir`` (``checkpoint-dit-N.bin``, ``checkpoint-project-clip-N.bin``,
``checkpoint-project-t5-N.bin`` | ``checkpoint-visual-adapter-N.bin``, ``optimizer-state-N.bin``;
/root/reference/Continuous/train_SigLIP_stage1.py:284-300, train_OpenAICLIP_video_stage1.py:504-512) -- while
accelerate / DeepSpeed / diffusers / omegaconf (not in the image) are replaced by this file: one process per GPU
(torchrun), ``parallel.GradReducer`` for the data-parallel exchange, ``optim.FusedAdamW`` for the update.

Data: the reference's loaders (``image_datasets/``: webdataset tars, CPU JPEG decode) are out of scope (SURVEY.md
2.1 #13).  ``data_config.img_dir`` / ``video_dir`` == "synthetic" (or a missing directory) selects seeded synthetic
batches with the reference's batch-dict schema; otherwise ``$GENHANCER_DATA_MODULE`` names a module exposing
``loader(**data_config)`` that yields such dicts.

Resume: the reference's resume logic is dead code; here ``resume_from_checkpoint: latest`` really
resumes from the newest ``checkpoint-dit-N.bin`` + ``optimizer-state-N.bin`` in ``output_dir``.
"""
from __future__ import annotations

import argparse
import importlib
import os
import re
import time
from types import SimpleNamespace

import torch

IMAGE_CKPT_STEPS = (50000,)                                   # train_SigLIP_stage1.py:284
STAGE2_CKPT_STEPS = (50, 100, 200, 300, 500, 1000, 2000, 3000)  # train_SigLIP_stage2_all.py:305


class Config(dict):
    """dict with attribute access, nested (what the scripts use OmegaConf for)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def get(self, k, default=None):  # noqa: D401
        return self[k] if k in self else default


def _wrap(o):
    if isinstance(o, dict):
        return Config({k: _wrap(v) for k, v in o.items()})
    if isinstance(o, list):
        return [_wrap(v) for v in o]
    if isinstance(o, str):  # YAML 1.1 reads "1e-4" as a string; OmegaConf users rely on float()
        if re.fullmatch(r"[+-]?\d+(\.\d*)?[eE][+-]?\d+", o):
            return float(o)
    return o


def load_config(path: str) -> Config:
    import yaml
    with open(path) as f:
        return _wrap(yaml.safe_load(f))


def parse_args(argv=None) -> str:
    p = argparse.ArgumentParser(description="GenHancer training entry point (B200-native).")
    p.add_argument("--config", type=str, default=None, required=True, help="path to config")
    return p.parse_args(argv).config


# ---------------------------------------------------------------------------------------------------------------
# data
# ---------------------------------------------------------------------------------------------------------------
def synthetic_loader(mode: str, batch: int, size: int, seed: int, device, frames: int = 8):
    """Endless seeded synthetic batches with the reference's schema."""
    gen = torch.Generator(device=device)
    step = 0
    while True:
        gen.manual_seed(seed + step)
        if mode == "image":
            yield {"image": torch.rand(batch, 3, size, size, device=device, generator=gen)}
        elif mode == "sliding_windows_nextpredic":
            yield {"full_frames": torch.rand(batch, frames, 3, size, size, device=device, generator=gen),
                   "frame_mask": torch.ones(batch, frames, dtype=torch.bool, device=device)}
        else:
            yield {k: torch.rand(batch, 3, size, size, device=device, generator=gen)
                   for k in ("start_frame", "middle_frame", "end_frame")}
        step += 1


def make_loader(args: Config, mode: str, device, rank: int):
    dc = args.data_config
    src = dc.get("img_dir") if mode == "image" else dc.get("video_dir")
    if src in (None, "synthetic") or not os.path.isdir(str(src)):
        if src not in (None, "synthetic"):
            print(f"[genhancer_b200] data directory {src!r} not found: using synthetic batches")
        return synthetic_loader(mode, dc.train_batch_size, dc.img_size, int(dc.get("seed", 0)) + 100003 * rank, device,
                                int(dc.get("max_frames_per_video", 8)))
    modname = os.environ.get("GENHANCER_DATA_MODULE")
    if not modname:
        raise RuntimeError("real datasets are outside this package: set GENHANCER_DATA_MODULE to a module exposing "
                           "loader(**data_config) (e.g. the reference's image_datasets.dataset_cc3m), or use "
                           "data_config.{img_dir,video_dir}: synthetic")
    return importlib.import_module(modname).loader(**dc)


# ---------------------------------------------------------------------------------------------------------------
# checkpoints (flat files in output_dir, reference names)
# ---------------------------------------------------------------------------------------------------------------
def save_checkpoint(out_dir: str, step: int, dit, clip_vis, adapter, opt, video: bool, save_project_clip: bool = True):
    os.makedirs(out_dir, exist_ok=True)
    sd = lambda m: {k: v.detach().clone().cpu() for k, v in m.state_dict().items()}  # no deepcopy of the module on the GPU
    torch.save(sd(dit), os.path.join(out_dir, f"checkpoint-dit-{step}.bin"))
    if save_project_clip:
        torch.save(sd(clip_vis.project_clip), os.path.join(out_dir, f"checkpoint-project-clip-{step}.bin"))
    if video:
        torch.save(sd(adapter), os.path.join(out_dir, f"checkpoint-visual-adapter-{step}.bin"))
    else:
        torch.save(sd(clip_vis.project_t5), os.path.join(out_dir, f"checkpoint-project-t5-{step}.bin"))
    torch.save(opt.state_dict(), os.path.join(out_dir, f"optimizer-state-{step}.bin"))


def merged_dir_name(family: str, args, step: int) -> str:
    """Directory of the LoRA-merged HF model, named as each reference script names it."""
    size = int(args.clip_config.clip_image_size) if args.clip_config.get("clip_image_size") else 224
    if family == "SigLIP":                                   # train_SigLIP_stage2_all.py:307
        return f"siglip-so400m-patch14-{size}-{step}"
    if family == "MetaCLIP":                                 # train_MetaCLIP_stage2_all.py:306-309
        return f"metaclip-{'l' if args.clip_config.clip_type == 'large' else 'h'}14-fullcc2.5b-{step}"
    return f"clip-vit-large-patch14-336-{step}" if size == 336 else f"clip-vit-large-patch14-{step}"  # OpenAI video scripts


def save_stage2(out_dir: str, step: int, family: str, args, dit, clip_vis, adapter, opt, video: bool):
    """Stage-2 outputs: the tower with LoRA merged as an HF directory (pytorch_model.bin, safe_serialization=False;
    train_SigLIP_stage2_all.py:305-311); the video scripts also write the DiT / project_clip / visual_adapter /
    optimizer files (train_OpenAICLIP_use2frames_nextpredic_stage2_all.py:469-494).  No deepcopy of the model on the
    GPU: the merge happens tensor by tensor on the way to the file."""
    from .clip_models import lora
    os.makedirs(out_dir, exist_ok=True)
    lora.save_pretrained(clip_vis.model, os.path.join(out_dir, merged_dir_name(family, args, step)))
    if video:
        sd = lambda m: {k: v.detach().clone().cpu() for k, v in m.state_dict().items()}
        torch.save(sd(dit), os.path.join(out_dir, f"checkpoint-dit-{step}.bin"))
        torch.save(sd(clip_vis.project_clip), os.path.join(out_dir, f"checkpoint-project-clip-{step}.bin"))
        torch.save(sd(adapter), os.path.join(out_dir, f"checkpoint-visual-adapter-{step}.bin"))
        torch.save(opt.state_dict(), os.path.join(out_dir, f"optimizer-state-{step}.bin"))


def latest_step(out_dir: str) -> int | None:
    if not os.path.isdir(out_dir):
        return None
    steps = [int(m.group(1)) for f in os.listdir(out_dir) if (m := re.fullmatch(r"checkpoint-dit-(\d+)\.bin", f))]
    return max(steps) if steps else None


def load_checkpoint(out_dir: str, step: int, dit, clip_vis, adapter, opt, video: bool, strict: bool = True):
    ld = lambda name: torch.load(os.path.join(out_dir, name), map_location="cpu", weights_only=True)
    with torch.no_grad():
        dit.load_state_dict(ld(f"checkpoint-dit-{step}.bin"), strict=strict)
        pc = os.path.join(out_dir, f"checkpoint-project-clip-{step}.bin")
        if os.path.exists(pc):
            clip_vis.project_clip.load_state_dict(ld(os.path.basename(pc)), strict=strict)
        if video:
            adapter.load_state_dict(ld(f"checkpoint-visual-adapter-{step}.bin"), strict=strict)
        else:
            clip_vis.project_t5.load_state_dict(ld(f"checkpoint-project-t5-{step}.bin"), strict=strict)
    op = os.path.join(out_dir, f"optimizer-state-{step}.bin")
    if opt is not None and os.path.exists(op):
        opt.load_state_dict(torch.load(op, map_location="cpu", weights_only=True))


# ---------------------------------------------------------------------------------------------------------------
# main
# ---------------------------------------------------------------------------------------------------------------
NORMS = {"OpenAICLIP": "openai", "MetaCLIP": "openai", "SigLIP": "siglip"}


def main(family: str, mode: str = "image", stage: str = "stage1", argv=None, max_steps_override: int | None = None):
    """family: OpenAICLIP | SigLIP | MetaCLIP; mode: image | video | nextpredic | use2frames_nextpredic |
    sliding_windows_nextpredic; stage: stage1 | stage2_all | stage2_only."""
    import torch.distributed as dist

    from . import optim
    from .clip_models import build_CLIP
    from .flux.util import load_ae, load_flow_model2
    from .parallel import GradReducer, broadcast_parameters
    from .train_step import (OPENAI_CLIP_MEAN, OPENAI_CLIP_STD, SIGLIP_MEAN, SIGLIP_STD, Stage1ImageStep)
    from .video import SuperModel, VideoStep, build_windows_with_mask

    args = load_config(parse_args(argv))
    if stage not in ("stage1", "stage2_all", "stage2_only"):
        raise ValueError(f"unknown stage {stage!r}")
    stage2 = stage != "stage1"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("genhancer_b200 needs a CUDA (sm_100a) device: there is no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=device)
    is_main = rank == 0
    if is_main and args.get("output_dir"):
        os.makedirs(args.output_dir, exist_ok=True)

    torch.manual_seed(int(args.get("seed", 0)))  # identical initial weights on every rank
    dit = load_flow_model2(args.model_name, device=device)
    vae = load_ae(args.model_name, device=device)
    clip_vis = getattr(build_CLIP, f"load_clip_model_{family}")(args.clip_config, device)
    video = mode != "image"
    # freeze: AE and tower; train DiT (+ projectors | adapter)   (train_SigLIP_stage1.py:130-141)
    vae.requires_grad_(False)
    clip_vis.requires_grad_(False)
    train_project_clip = mode != "sliding_windows_nextpredic"   # that script neither trains nor saves it (Q12)
    if stage2:
        # LoRA on the tower (train_SigLIP_stage2_all.py:134-142; OpenAI / MetaCLIP scripts use target_modules="all-linear")
        from .clip_models import lora
        lc = args.lora_config
        targets = lora.SIGLIP_TARGETS if family == "SigLIP" else "all-linear"
        clip_vis.model = lora.get_peft_model(clip_vis.model, lora.LoraConfig(
            r=int(lc.r), lora_alpha=int(lc.lora_alpha), target_modules=targets, lora_dropout=float(lc.get("lora_dropout", 0.0)),
            bias=str(lc.get("bias", "none"))))
        if is_main:
            lora.print_trainable_parameters(clip_vis.model)
        clip_vis.train()
    if stage != "stage2_only":  # stage2_only: DiT and projectors frozen, only the LoRA'd tower trains (train_SigLIP_stage2_only.py:146-155)
        for n, p in clip_vis.named_parameters():
            if ("project_clip" in n and train_project_clip and not video) or ("project_t5" in n and not video):
                p.requires_grad = True
    dit = dit.to(device).to(torch.bfloat16)
    dit.requires_grad_(stage != "stage2_only")
    dit.train()
    mean, std = (SIGLIP_MEAN, SIGLIP_STD) if NORMS[family] == "siglip" else (OPENAI_CLIP_MEAN, OPENAI_CLIP_STD)
    trainable = list(dit.named_parameters())
    adapter = None
    if video:
        feat = clip_vis.model.config.hidden_size
        sm = SuperModel(clip_vis, dit, adapter_in_dim=feat, adapter_out_dim=dit.params.context_in_dim).to(device)
        adapter = sm.visual_adapter.float()
        trainable += [(f"visual_adapter.{n}", p) for n, p in adapter.named_parameters()]
        if stage2:
            trainable += [(f"clip_vis.{n}", p) for n, p in clip_vis.named_parameters()]
        times = {"video": ((0, 2), 1), "nextpredic": ((0,), 1), "use2frames_nextpredic": ((0, 1), 2),
                 "sliding_windows_nextpredic": ((0, 1, 2), 3)}[mode]
        step_fn = VideoStep(sm, vae, cond_times=times[0], target_time=times[1], clip_mean=mean, clip_std=std,
                            scale_factor=float(args.scale_factor), tower_grad=stage2)
    else:
        trainable += [(f"clip_vis.{n}", p) for n, p in clip_vis.named_parameters()]
        step_fn = Stage1ImageStep(clip_vis, dit, vae, mean, std, scale_factor=float(args.scale_factor))
    if stage2 and args.get("load_dir") is not None and args.get("load_step") is not None:
        # stage 2 starts from the stage-1 DiT / projectors / adapter (train_SigLIP_stage2_all.py:146-156)
        if os.path.exists(os.path.join(str(args.load_dir), f"checkpoint-dit-{args.load_step}.bin")):
            load_checkpoint(str(args.load_dir), int(args.load_step), dit, clip_vis, adapter, None, video, strict=True)
            if is_main:
                print(f"[genhancer_b200] loaded stage-1 weights from {args.load_dir} step {args.load_step}")
        elif is_main:
            print(f"[genhancer_b200] stage-1 checkpoint {args.load_dir}/checkpoint-dit-{args.load_step}.bin not found: "
                  "starting stage 2 from the current (random) initialisation")
    groups = optim.flatten(trainable)
    broadcast_parameters(groups)
    opt = optim.FusedAdamW(groups, lr=float(args.learning_rate), betas=(float(args.adam_beta1), float(args.adam_beta2)),
                           eps=float(args.adam_epsilon), weight_decay=float(args.adam_weight_decay),
                           max_grad_norm=float(args.max_grad_norm), engine_managed=[dit])
    reducer = GradReducer(groups, engine_modules=[dit]) if world > 1 else None
    ga = int(args.get("gradient_accumulation_steps", 1))

    global_step = 0
    if args.get("resume_from_checkpoint") and args.get("output_dir"):
        last = latest_step(args.output_dir) if args.resume_from_checkpoint == "latest" else \
            int(re.findall(r"\d+", str(args.resume_from_checkpoint))[-1])
        if last is not None:
            load_checkpoint(args.output_dir, last, dit, clip_vis, adapter, opt, video, strict=True)
            global_step = last
            if is_main:
                print(f"[genhancer_b200] resumed from step {last}")
    max_steps = max_steps_override or int(args.max_train_steps)
    loader = make_loader(args, mode, device, rank)
    ckpt_extra = STAGE2_CKPT_STEPS if stage2 else IMAGE_CKPT_STEPS
    losses, micro = [], 0
    st = SimpleNamespace(step=global_step, pending=False, t_last=time.time())
    train_loss = torch.zeros((), device=device)
    pending_loss = torch.zeros((), device=device)

    def finish_step():
        """Gradient exchange (wait) + clip + AdamW + bookkeeping of the optimizer step whose backward has been issued.
        With data parallelism it runs INSIDE the next micro-step, right before the first kernel that reads a trainable
        parameter (``before_trainable``), so the all-reduce tail hides under the frozen AE / tower forward."""
        if not st.pending:
            return
        st.pending = False
        if reducer is not None:
            reducer.wait()
        opt.step(reducer.grad_scale if reducer else 1.0)
        opt.zero_grad()
        st.step += 1
        if st.step % 10 == 0 or st.step == max_steps:
            lv = float(pending_loss)        # one host sync per logged step (the reference syncs every micro-step)
            losses.append(lv)
            if is_main:
                dt = time.time() - st.t_last
                print(f"[genhancer_b200] step {st.step} loss {lv:.4f} grad-norm {float(opt.grad_norm()):.3f} "
                      f"({dt:.2f} s since last log)", flush=True)
            st.t_last = time.time()
        if is_main and args.get("output_dir") and (st.step % int(args.checkpointing_steps) == 0
                                                   or st.step in ckpt_extra or st.step >= max_steps):
            if stage2:
                save_stage2(args.output_dir, st.step, family, args, dit, clip_vis, adapter, opt, video)
            else:
                save_checkpoint(args.output_dir, st.step, dit, clip_vis, adapter, opt, video,
                                save_project_clip=train_project_clip)

    def to_input(t):
        """A loader may hand over decoded frames as uint8 [B,H,W,3] (1 byte per value over PCIe): ToTensor then runs
        on the device (gh_u8hwc_to_f32chw); fp32 [B,3,H,W] in [0,1] batches pass through as before."""
        t = t.to(device, non_blocking=True)
        if t.dtype == torch.uint8 and t.dim() == 4 and t.shape[-1] == 3:
            from . import kernels as K
            return K.u8hwc_to_f32chw(t)
        return t.float()

    for batch in loader:
        if st.step + int(st.pending) >= max_steps:
            break
        sync = (micro + 1) % ga == 0
        if reducer is not None:
            reducer.enabled = sync
        if mode == "image":
            loss = step_fn(to_input(batch["image"]), before_trainable=finish_step)
        elif mode == "sliding_windows_nextpredic":
            w = build_windows_with_mask(batch["full_frames"].to(device), batch["frame_mask"].to(device),
                                        int(args.get("window_cond", 3)), int(args.get("window_stride", 1)),
                                        args.get("max_windows_per_video", 8))
            if w is None:
                continue
            loss = step_fn(list(w[:3]), w[3], before_trainable=finish_step)
        else:
            f = {k: to_input(batch[k]) for k in ("start_frame", "middle_frame", "end_frame")}
            cond, tgt = {"video": (("start_frame", "end_frame"), "middle_frame"),
                         "nextpredic": (("start_frame",), "middle_frame"),
                         "use2frames_nextpredic": (("start_frame", "middle_frame"), "end_frame")}[mode]
            loss = step_fn([f[k] for k in cond], f[tgt], before_trainable=finish_step)
        finish_step()                   # (no-op if the step object already called it)
        (loss / ga).backward()          # accelerator.backward divides by gradient_accumulation_steps
        train_loss += loss.detach() / ga
        micro += 1
        if not sync:
            continue
        if reducer is not None:
            reducer.issue_rest()        # projector / adapter groups: issued now, waited for in finish_step()
        pending_loss.copy_(train_loss)
        train_loss.zero_()
        st.pending = True
        if reducer is None:
            finish_step()               # single GPU: nothing to hide
    finish_step()
    global_step = st.step
    if world > 1:
        dist.barrier()
    return SimpleNamespace(global_step=global_step, losses=losses, dit=dit, clip_vis=clip_vis, adapter=adapter, opt=opt)
