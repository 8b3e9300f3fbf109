"""Host of the ``train_*.py`` entry points: ``python train_X.py --config <yaml>``.

Keeps the reference's contract -- the single ``--config`` flag, the YAML schema, the loop
semantics (gradient accumulation, clip to ``max_grad_norm`` on sync steps, AdamW, constant LR) and the flat
checkpoint layout in ``output_dir`` (``checkpoint-dit-N.bin``, ``checkpoint-project-clip-N.bin``,
``checkpoint-project-t5-N.bin`` | ``checkpoint-visual-adapter-N.bin``, ``optimizer-state-N.bin``;
/root/reference/Continuous/train_SigLIP_stage1.py:284-300, train_OpenAICLIP_video_stage1.py:504-512) -- while
accelerate / DeepSpeed / diffusers / omegaconf (not in the image) are replaced by this file: one process per GPU
(torchrun), ``parallel.GradReducer`` for the data-parallel exchange, ``optim.FusedAdamW`` for the update.

Data: the reference's loaders (``image_datasets/``: webdataset tars, CPU JPEG decode) are out of scope (SURVEY.md
2.1 #13).  ``data_config.img_dir`` / ``video_dir`` == "synthetic" selects seeded synthetic batches with the
reference's batch-dict schema; a directory selects ``$GENHANCER_DATA_MODULE``, a module exposing
``loader(**data_config)`` that yields such dicts.  A path that does not exist is an error (as in the reference, which
crashes), never a silent switch to synthetic data.

Execution: whenever the micro-step has static shapes (every mode but sliding windows over ragged clips) the loop
replays ``graph.PipelinedTrainStep`` -- forward, backward, the overlapped NCCL exchange AND clip + AdamW as one CUDA
graph per micro-step, the update of step n on a forked branch under the frozen forward of step n + 1 -- i.e. the path
``bench.py`` measures.  ``cuda_graph: false`` in the YAML (or GH_TRAINER_GRAPH=0) selects plain eager launches.

Resume: the reference's resume logic is dead code; here ``resume_from_checkpoint: latest`` really
resumes from the newest ``checkpoint-dit-N.bin`` + ``optimizer-state-N.bin`` in ``output_dir`` (stage 2: plus
``checkpoint-tower-lora-N.bin``, the un-merged LoRA pairs and trainable biases, which the merged HF export cannot
give back).

Seeds: ``seed`` (YAML, default 0) fixes the initial weights on every rank; after the parameter broadcast each rank
re-seeds with ``seed + 1000003 * rank`` so that t, x_0, the AE noise and the LoRA dropout masks differ across ranks
(the reference never seeds, so its ranks draw independently).
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
    if src is None:
        raise KeyError(f"data_config.{'img_dir' if mode == 'image' else 'video_dir'} is missing (use 'synthetic' for seeded "
                       "synthetic batches)")
    if src != "synthetic" and not os.path.isdir(str(src)):
        raise FileNotFoundError(f"data directory {src!r} does not exist (use 'synthetic' for seeded synthetic batches)")
    if src == "synthetic":
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
    sd = lambda m: {k: v.detach().cpu() for k, v in m.state_dict().items()}  # no deepcopy of the module on the GPU
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


def save_stage2(out_dir: str, step: int, family: str, args, dit, clip_vis, adapter, opt, video: bool,
                dit_trains: bool = True):
    """Stage-2 outputs: the tower with LoRA merged as an HF directory (pytorch_model.bin, safe_serialization=False;
    train_SigLIP_stage2_all.py:305-311); the video scripts also write the DiT / project_clip / visual_adapter /
    optimizer files (train_OpenAICLIP_use2frames_nextpredic_stage2_all.py:469-494).  No deepcopy of the model on the
    GPU: the merge happens tensor by tensor on the way to the file."""
    from .clip_models import lora
    os.makedirs(out_dir, exist_ok=True)
    lora.save_pretrained(clip_vis.model, os.path.join(out_dir, merged_dir_name(family, args, step)))
    # Resumable state.  The video scripts write these files themselves; the image scripts write only the merged
    # directory, from which a run cannot continue (the LoRA pairs are gone) -- the extra flat files are what makes
    # `resume_from_checkpoint: latest` real in stage 2.
    sd = lambda m: {k: v.detach().cpu() for k, v in m.state_dict().items()}
    if dit_trains or video:          # (stage2_only: the DiT / projectors are the frozen stage-1 files, nothing to re-save)
        torch.save(sd(dit), os.path.join(out_dir, f"checkpoint-dit-{step}.bin"))
        torch.save(sd(clip_vis.project_clip), os.path.join(out_dir, f"checkpoint-project-clip-{step}.bin"))
        if video:
            torch.save(sd(adapter), os.path.join(out_dir, f"checkpoint-visual-adapter-{step}.bin"))
        else:
            torch.save(sd(clip_vis.project_t5), os.path.join(out_dir, f"checkpoint-project-t5-{step}.bin"))
    torch.save(tower_adapter_state(clip_vis), os.path.join(out_dir, f"checkpoint-tower-lora-{step}.bin"))
    torch.save(opt.state_dict(), os.path.join(out_dir, f"optimizer-state-{step}.bin"))


def tower_adapter_state(clip_vis) -> dict:
    """What the merged HF export cannot give back: the un-merged LoRA pairs and the trainable (``bias='lora_only'``)
    biases of the tower -- the resumable half of a stage-2 checkpoint."""
    model = clip_vis.model
    sd = {f"lora.{k}": v.detach().cpu() for k, v in model.lora.state_dict().items()} if hasattr(model, "lora") else {}
    for n, p in model.named_parameters():
        if p.requires_grad and not n.startswith("lora."):
            sd[n] = p.detach().cpu()
    return sd


def load_tower_adapter_state(clip_vis, sd: dict) -> None:
    model = clip_vis.model
    own = dict(model.named_parameters())
    with torch.no_grad():
        for k, v in sd.items():
            if k not in own:
                raise KeyError(f"checkpoint-tower-lora: unexpected tensor {k}")
            own[k].copy_(v)
    missing = [n for n, p in own.items() if p.requires_grad and n not in sd]
    if missing:
        raise KeyError(f"checkpoint-tower-lora lacks trainable tower tensors: {missing[:4]} ...")


def latest_step(out_dir: str) -> int | None:
    if not os.path.isdir(out_dir):
        return None
    steps = [int(m.group(1)) for f in os.listdir(out_dir)
             if (m := re.fullmatch(r"checkpoint-(?:dit|tower-lora)-(\d+)\.bin", f))]
    return max(steps) if steps else None


def load_checkpoint(out_dir: str, step: int, dit, clip_vis, adapter, opt, video: bool, strict: bool = True,
                    tower_lora: bool = False):
    """``tower_lora``: resuming a stage-2 run -- the un-merged LoRA pairs / trainable biases must be there too."""
    ld = lambda name: torch.load(os.path.join(out_dir, name), map_location="cpu", weights_only=True)
    if tower_lora:
        tl = os.path.join(out_dir, f"checkpoint-tower-lora-{step}.bin")
        if not os.path.exists(tl):
            raise FileNotFoundError(f"cannot resume stage 2 from step {step}: {tl} is missing (the merged HF directory "
                                    "alone does not hold the LoRA pairs or the optimizer's view of them)")
        load_tower_adapter_state(clip_vis, ld(os.path.basename(tl)))
    # (stage2_only writes no DiT / projector files: they stay the stage-1 weights loaded through load_dir / load_step)
    if not tower_lora or os.path.exists(os.path.join(out_dir, f"checkpoint-dit-{step}.bin")):
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
    if stage2 and args.get("load_dir") is not None and args.get("load_step") is not None and str(args.load_dir).lower() != "none":
        # stage 2 starts from the stage-1 DiT / projectors / adapter (train_SigLIP_stage2_all.py:146-156); a missing file
        # is an error, as in the reference (torch.load raises) -- never a silent start from random weights
        f0 = os.path.join(str(args.load_dir), f"checkpoint-dit-{args.load_step}.bin")
        if not os.path.exists(f0):
            raise FileNotFoundError(f"stage-1 checkpoint {f0} not found (load_dir / load_step of the YAML); set load_dir: none "
                                    "to start stage 2 from the current initialisation on purpose")
        load_checkpoint(str(args.load_dir), int(args.load_step), dit, clip_vis, adapter, None, video, strict=True)
        if is_main:
            print(f"[genhancer_b200] loaded stage-1 weights from {args.load_dir} step {args.load_step}")
    groups = optim.flatten(trainable)
    broadcast_parameters(groups)
    # identical weights everywhere; from here on every rank draws its own t / x_0 / AE noise / dropout masks
    seed = int(args.get("seed", 0))
    torch.manual_seed(seed + 1000003 * rank)
    opt = optim.FusedAdamW(groups, lr=float(args.learning_rate), betas=(float(args.adam_beta1), float(args.adam_beta2)),
                           eps=float(args.adam_epsilon), weight_decay=float(args.adam_weight_decay),
                           max_grad_norm=float(args.max_grad_norm), engine_managed=[dit])
    reducer = GradReducer(groups, engine_modules=[dit]) if world > 1 else None
    gscale = reducer.grad_scale if reducer else 1.0
    ga = int(args.get("gradient_accumulation_steps", 1))

    global_step = 0
    if args.get("resume_from_checkpoint") and args.get("output_dir"):
        last = latest_step(args.output_dir) if args.resume_from_checkpoint == "latest" else \
            int(re.findall(r"\d+", str(args.resume_from_checkpoint))[-1])
        if last is not None:
            load_checkpoint(args.output_dir, last, dit, clip_vis, adapter, opt, video, strict=True, tower_lora=stage2)
            global_step = last
            if is_main:
                print(f"[genhancer_b200] resumed from step {last}")
    max_steps = max_steps_override or int(args.max_train_steps)
    loader = make_loader(args, mode, device, rank)
    ckpt_extra = STAGE2_CKPT_STEPS if stage2 else IMAGE_CKPT_STEPS
    losses = []
    st = SimpleNamespace(step=global_step, micro=0, t_last=time.time(), pipe=None, graph_steps=0, eager_steps=0)
    train_loss = torch.zeros((), device=device)
    use_graph = bool(args.get("cuda_graph", True)) and os.environ.get("GH_TRAINER_GRAPH", "1") != "0"
    # ragged clips give a different number of windows per batch: graph replay only where the shapes repeat, and only
    # without accumulation (a cycle cannot switch between replay and eager half way)
    if mode == "sliding_windows_nextpredic" and ga > 1:
        use_graph = False

    def to_input(t):
        """A loader may hand over decoded frames as uint8 [B,H,W,3]: they cross PCIe at 1 byte per value and go to the
        step AS THEY ARE -- ToTensor (u8 / 255) and the two Normalize transforms happen inside the patch-embed / conv_in
        gathers (gh_patch_im2col_u8hwc, gh_im2col3x3_c3_u8hwc), bit-identical to the fp32 path; fp32 [B,3,H,W] batches
        in [0,1] pass through as before."""
        t = t.to(device, non_blocking=True)
        if t.dtype == torch.uint8 and t.dim() == 4 and t.shape[-1] == 3:
            return t.contiguous()
        return t.float()

    if mode == "image":
        call = lambda img, before_trainable=None: step_fn(img, before_trainable=before_trainable)      # noqa: E731
    else:
        call = lambda *fr, before_trainable=None: step_fn(list(fr[:-1]), fr[-1], before_trainable=before_trainable)  # noqa: E731

    def batch_inputs(batch):
        if mode == "image":
            return (to_input(batch["image"]),)
        if mode == "sliding_windows_nextpredic":
            w = build_windows_with_mask(batch["full_frames"].to(device), batch["frame_mask"].to(device),
                                        int(args.get("window_cond", 3)), int(args.get("window_stride", 1)),
                                        args.get("max_windows_per_video", 8))
            return None if w is None else tuple(w[:-2])
        f = {k: to_input(batch[k]) for k in ("start_frame", "middle_frame", "end_frame")}
        cond, tgt = {"video": (("start_frame", "end_frame"), "middle_frame"),
                     "nextpredic": (("start_frame",), "middle_frame"),
                     "use2frames_nextpredic": (("start_frame", "middle_frame"), "end_frame")}[mode]
        return tuple(f[k] for k in cond) + (f[tgt],)

    def flush():
        if st.pipe is not None:
            st.pipe.flush()

    def after_optimizer_step(loss_value):
        """Bookkeeping once the backward of optimizer step ``st.step`` has been issued: logging (one host sync per
        logged step; the reference syncs every micro-step) and checkpoints (which first apply the pending update)."""
        if st.step % 10 == 0 or st.step >= max_steps:
            lv = float(loss_value)
            losses.append(lv)
            flush()
            if is_main:
                dt = time.time() - st.t_last
                print(f"[genhancer_b200] step {st.step} loss {lv:.4f} grad-norm {float(opt.grad_norm()):.3f} "
                      f"({dt:.2f} s since last log)", flush=True)
            st.t_last = time.time()
        if is_main and args.get("output_dir") and (st.step % int(args.checkpointing_steps) == 0
                                                   or st.step in ckpt_extra or st.step >= max_steps):
            flush()
            if stage2:
                save_stage2(args.output_dir, st.step, family, args, dit, clip_vis, adapter, opt, video,
                            dit_trains=stage != "stage2_only")
            else:
                save_checkpoint(args.output_dir, st.step, dit, clip_vis, adapter, opt, video,
                                save_project_clip=train_project_clip)

    first = True
    for batch in loader:
        if st.step >= max_steps:
            break
        inputs = batch_inputs(batch)
        if inputs is None:
            continue
        sync = (st.micro + 1) % ga == 0
        replay = False
        if use_graph:
            if st.pipe is None:
                from .graph import PipelinedTrainStep
                st.pipe = PipelinedTrainStep(call, inputs, opt, reducer, grad_accum=ga)
            replay = st.pipe.matches(*inputs)
        if first:
            # capture + warm-up consumed RNG draws: both execution modes start the real steps from the same generator
            # state, so a graph run and an eager run of one config see the same t / x_0 / noise sequence
            torch.manual_seed(seed + 1000003 * rank + 1)
            first = False
        if replay:
            loss = st.pipe(*inputs)
            st.graph_steps += 1
        else:
            # eager micro-step (variable shapes, or graphs switched off): the sequential loop of the reference,
            # train_SigLIP_stage1.py:238-275, with the exchange still overlapped with the backward
            flush()
            if reducer is not None:
                reducer.enabled = sync
            loss = call(*inputs)
            (loss / ga).backward()          # accelerator.backward divides by gradient_accumulation_steps
            if sync:
                if reducer is not None:
                    reducer.finish()
                opt.step(gscale)
                opt.zero_grad()
                opt.sync_device_state()
            st.eager_steps += 1
        train_loss += loss.detach().reshape(()) / ga
        st.micro = (st.micro + 1) % ga
        if not sync:
            continue
        st.step += 1
        after_optimizer_step(train_loss)
        train_loss.zero_()
    flush()
    global_step = st.step
    if world > 1:
        dist.barrier()
    if st.pipe is not None and world > 1:
        st.pipe.reset()        # a graph that holds NCCL kernels must go before the communicator does
    return SimpleNamespace(global_step=global_step, losses=losses, dit=dit, clip_vis=clip_vis, adapter=adapter, opt=opt,
                           graph_steps=st.graph_steps, eager_steps=st.eager_steps)
