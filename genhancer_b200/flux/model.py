"""``Flux`` -- drop-in for /root/reference/Continuous/src/flux/model.py (same constructor, forward signature,
error behaviour and state_dict), executed by the fused sm_100a engine."""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import Tensor, nn

from . import engine
from .modules.layers import DoubleStreamBlock, EmbedND, LastLayer, MLPEmbedder, SingleStreamBlock


@dataclass
class FluxParams:  # model.py:12-25
    in_channels: int
    vec_in_dim: int
    context_in_dim: int
    hidden_size: int
    mlp_ratio: float
    num_heads: int
    depth: int
    depth_single_blocks: int
    axes_dim: list[int]
    theta: int
    qkv_bias: bool
    guidance_embed: bool


class _FluxFn(torch.autograd.Function):
    """One autograd node for the whole DiT: forward and backward are the engine's explicit schedules.
    Parameter gradients are accumulated directly into ``param.grad`` by the wgrad GEMM epilogues."""

    @staticmethod
    def forward(ctx, model, img, img_ids, txt, txt_ids, timesteps, y, guidance, *params):
        P = model._param_dict()
        pred, ectx = engine.flux_forward(P, model.params, img, img_ids, txt, txt_ids, timesteps, y, guidance)
        ctx.model, ctx.ectx = model, ectx
        ctx.in_dtypes = (img.dtype, txt.dtype, y.dtype)
        ctx.needs = (img.requires_grad, txt.requires_grad, y.requires_grad)
        return pred

    @staticmethod
    def backward(ctx, dpred):
        model = ctx.model
        P = model._param_dict()
        # FusedAdamW.zero_grad() marks the flat gradient buffer "overwrite on next backward" instead of memsetting it
        accumulate = not model._grad_overwrite
        model._grad_overwrite = False
        sink = engine.GradSink(P, accumulate=accumulate, on_ready=model._on_grads_ready)
        d_img, d_txt, d_y = engine.flux_backward(P, model.params, ctx.ectx, dpred, sink, need_dimg=ctx.needs[0],
                                                 need_dtxt=ctx.needs[1], need_dy=ctx.needs[2])
        ctx.ectx = None
        cast = lambda g, dt: None if g is None else g.to(dt)
        return (None, cast(d_img, ctx.in_dtypes[0]), None, cast(d_txt, ctx.in_dtypes[1]), None, None,
                cast(d_y, ctx.in_dtypes[2]), None) + (None,) * len(P)


class Flux(nn.Module):
    """Transformer model for flow matching on sequences (lightweight FLUX DiT)."""

    def __init__(self, params: FluxParams):
        super().__init__()
        self.params = params
        self.in_channels = params.in_channels
        self.out_channels = self.in_channels
        if params.hidden_size % params.num_heads != 0:
            raise ValueError(f"Hidden size {params.hidden_size} must be divisible by num_heads {params.num_heads}")
        pe_dim = params.hidden_size // params.num_heads
        if sum(params.axes_dim) != pe_dim:
            raise ValueError(f"Got {params.axes_dim} but expected positional dim {pe_dim}")
        self.hidden_size = params.hidden_size
        self.num_heads = params.num_heads
        self.pe_embedder = EmbedND(dim=pe_dim, theta=params.theta, axes_dim=params.axes_dim)
        self.img_in = nn.Linear(self.in_channels, self.hidden_size, bias=True)
        self.time_in = MLPEmbedder(in_dim=256, hidden_dim=self.hidden_size)
        self.vector_in = MLPEmbedder(params.vec_in_dim, self.hidden_size)
        self.guidance_in = MLPEmbedder(in_dim=256, hidden_dim=self.hidden_size) if params.guidance_embed else nn.Identity()
        self.txt_in = nn.Linear(params.context_in_dim, self.hidden_size)
        self.double_blocks = nn.ModuleList([
            DoubleStreamBlock(self.hidden_size, self.num_heads, mlp_ratio=params.mlp_ratio, qkv_bias=params.qkv_bias)
            for _ in range(params.depth)])
        self.single_blocks = nn.ModuleList([
            SingleStreamBlock(self.hidden_size, self.num_heads, mlp_ratio=params.mlp_ratio)
            for _ in range(params.depth_single_blocks)])
        self.final_layer = LastLayer(self.hidden_size, 1, self.out_channels)
        self._grad_overwrite = False   # set by optim.FusedAdamW.zero_grad()
        self._on_grads_ready = None    # set by parallel.DataParallel: fires a bucket all-reduce per finished block
        self.gradient_checkpointing = False  # reference's branch is dead code; nothing is recomputed
        if pe_dim != 128:
            raise NotImplementedError("the sm_100a attention / QK-norm kernels are built for head_dim 128 (flux-dev)")

    def _param_dict(self) -> dict[str, Tensor]:
        P = dict(self.named_parameters())
        for n, p in P.items():
            if p.dtype != torch.bfloat16 or not p.is_cuda:
                raise RuntimeError(
                    f"Flux parameter {n} is {p.dtype} on {p.device}: the B200 engine runs the DiT in bf16 on CUDA "
                    "(the training scripts do `dit.to(torch.bfloat16)`, train_SigLIP_stage1.py:131-132)")
            break
        return P

    def forward(self, img: Tensor, img_ids: Tensor, txt: Tensor, txt_ids: Tensor, timesteps: Tensor, y: Tensor,
                block_controlnet_hidden_states=None, guidance: Tensor | None = None, image_proj: Tensor | None = None,
                ip_scale: Tensor | float = 1.0) -> Tensor:
        if img.ndim != 3 or txt.ndim != 3:
            raise ValueError("Input img and txt tensors must have 3 dimensions.")
        if self.params.guidance_embed and guidance is None:
            raise ValueError("Didn't get guidance strength for guidance distilled model.")
        if block_controlnet_hidden_states is not None or image_proj is not None:
            raise NotImplementedError("controlnet / IP-adapter inputs are unused by every GenHancer training script")
        params = tuple(self.parameters())
        pred = _FluxFn.apply(self, img, img_ids, txt, txt_ids, timesteps, y, guidance, *params)
        return pred.to(img.dtype) if img.dtype != torch.bfloat16 else pred
