"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

A plain-PyTorch (fp32, CPU or any device) functional restatement of the arithmetic on
GenHancer's stage-1/stage-2 training-step hot path, written against state_dicts that use the
reference's own key names.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
cpu_baseline / ``--impl reference`` legs may import it, and only as the checker.

Parity status: PINNED against outputs of the reference itself run in the build container
(``oracle/make_golden.py`` imports /root/reference/Continuous with two import shims, feeds both
implementations the same synthetic weights and inputs, asserts agreement and writes the
fixtures under ``tests/golden/``).  The reference ships no tests or golden vectors of its own
(SURVEY.md section 4), so those fixtures are the only anchor there is.

Citations are to /root/reference/Continuous/... (abbreviated R/) and to HF transformers
(modeling_clip.py / modeling_siglip.py, the un-vendored third-party tower the reference calls;
pinned 4.43.3 by R/requirements.txt:98, 5.5.0 installed here -- same arithmetic, verified).
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# configs (literals of R/src/flux/util.py:131-156 and the HF model cards)
# --------------------------------------------------------------------------------------


@dataclass
class FluxCfg:  # R/src/flux/model.py:12-25
    in_channels: int = 64
    vec_in_dim: int = 768
    context_in_dim: int = 4096
    hidden_size: int = 3072
    mlp_ratio: float = 4.0
    num_heads: int = 24
    depth: int = 2
    depth_single_blocks: int = 4
    axes_dim: tuple = (16, 56, 56)
    theta: int = 10_000
    qkv_bias: bool = True
    guidance_embed: bool = True


@dataclass
class AECfg:  # R/src/flux/modules/autoencoder.py:8-18, util.py:146-156
    resolution: int = 256
    in_channels: int = 3
    ch: int = 128
    out_ch: int = 3
    ch_mult: tuple = (1, 2, 4, 4)
    num_res_blocks: int = 2
    z_channels: int = 16
    scale_factor: float = 0.3611
    shift_factor: float = 0.1159


@dataclass
class TowerCfg:
    kind: str = "clip"  # "clip" (OpenAI / MetaCLIP) or "siglip"
    hidden: int = 1024
    layers: int = 24
    heads: int = 16
    mlp: int = 4096
    image_size: int = 224
    patch: int = 14
    proj_dim: int = 768  # visual_projection output (clip only)
    eps: float = 1e-5
    act: str = "quick_gelu"

    @property
    def grid(self) -> int:
        return self.image_size // self.patch

    @property
    def tokens(self) -> int:
        return self.grid * self.grid + (1 if self.kind == "clip" else 0)

    @property
    def feat_dim(self) -> int:  # dim of the class token handed to the projectors
        return self.proj_dim if self.kind == "clip" else self.hidden


def openai_vit_l14(image_size: int = 224) -> TowerCfg:
    return TowerCfg("clip", 1024, 24, 16, 4096, image_size, 14, 768, 1e-5, "quick_gelu")


def siglip_so400m(image_size: int = 384) -> TowerCfg:
    return TowerCfg("siglip", 1152, 27, 16, 4304, image_size, 14, 1152, 1e-6, "gelu_tanh")


# --------------------------------------------------------------------------------------
# deterministic synthetic weights (shared by make_golden.py, tests and bench)
# --------------------------------------------------------------------------------------


def synth_tensor(key: str, shape, seed: int) -> torch.Tensor:
    """Platform-independent pseudo-random init of one tensor, keyed by its state_dict name."""
    g = torch.Generator().manual_seed((seed * 1_000_003 + zlib.crc32(key.encode())) % (2**63 - 1))
    shape = tuple(shape)
    leaf = key.rsplit(".", 1)[-1]
    is_norm = any(s in key for s in ("norm", "layrnorm", "ln_", "layer_norm")) or (".proj.3." in key) or key.endswith(
        ("project_clip.0.weight", "project_clip.0.bias", "project_t5.0.weight", "project_t5.0.bias"))
    if leaf == "scale":  # RMSNorm scale
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if is_norm and leaf == "weight":
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if is_norm and leaf == "bias":
        return 0.1 * torch.randn(shape, generator=g)
    if leaf in ("class_embedding", "probe") or "position_embedding" in key:
        return 0.02 * torch.randn(shape, generator=g)
    if len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        b = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * b
    # biases of linear / conv layers
    return (torch.rand(shape, generator=g) * 2 - 1) * 0.05


def synth_state_dict(key_shapes: dict, seed: int) -> dict:
    return {k: synth_tensor(k, s, seed) for k, s in key_shapes.items()}


def flux_key_shapes(c: FluxCfg) -> dict:
    """Names/shapes of Flux.state_dict() (R/src/flux/model.py:35-79, layers.py)."""
    H, mlp = c.hidden_size, int(c.hidden_size * c.mlp_ratio)
    D = H // c.num_heads
    ks: dict = {}

    def lin(name, o, i):
        ks[f"{name}.weight"] = (o, i)
        ks[f"{name}.bias"] = (o,)

    lin("img_in", H, c.in_channels)
    for n, i in (("time_in", 256), ("vector_in", c.vec_in_dim), ("guidance_in", 256)):
        lin(f"{n}.in_layer", H, i)
        lin(f"{n}.out_layer", H, H)
    lin("txt_in", H, c.context_in_dim)
    for b in range(c.depth):
        p = f"double_blocks.{b}"
        for s in ("img", "txt"):
            lin(f"{p}.{s}_mod.lin", 6 * H, H)
            lin(f"{p}.{s}_attn.qkv", 3 * H, H)
            ks[f"{p}.{s}_attn.norm.query_norm.scale"] = (D,)
            ks[f"{p}.{s}_attn.norm.key_norm.scale"] = (D,)
            lin(f"{p}.{s}_attn.proj", H, H)
            lin(f"{p}.{s}_mlp.0", mlp, H)
            lin(f"{p}.{s}_mlp.2", H, mlp)
    for b in range(c.depth_single_blocks):
        p = f"single_blocks.{b}"
        lin(f"{p}.linear1", 3 * H + mlp, H)
        lin(f"{p}.linear2", H, H + mlp)
        ks[f"{p}.norm.query_norm.scale"] = (D,)
        ks[f"{p}.norm.key_norm.scale"] = (D,)
        lin(f"{p}.modulation.lin", 3 * H, H)
    lin("final_layer.linear", c.in_channels, H)
    lin("final_layer.adaLN_modulation.1", 2 * H, H)
    return ks


def ae_encoder_key_shapes(c: AECfg) -> dict:
    """Names/shapes of AutoEncoder.encoder.state_dict() (R/src/flux/modules/autoencoder.py:109-157)."""
    ks: dict = {}

    def conv(name, o, i, k):
        ks[f"{name}.weight"] = (o, i, k, k)
        ks[f"{name}.bias"] = (o,)

    def gn(name, ch):
        ks[f"{name}.weight"] = (ch,)
        ks[f"{name}.bias"] = (ch,)

    def res(name, i, o):
        gn(f"{name}.norm1", i)
        conv(f"{name}.conv1", o, i, 3)
        gn(f"{name}.norm2", o)
        conv(f"{name}.conv2", o, o, 3)
        if i != o:
            conv(f"{name}.nin_shortcut", o, i, 1)

    conv("conv_in", c.ch, c.in_channels, 3)
    in_mult = (1,) + tuple(c.ch_mult)
    block_in = c.ch
    for lvl in range(len(c.ch_mult)):
        block_in = c.ch * in_mult[lvl]
        block_out = c.ch * c.ch_mult[lvl]
        for j in range(c.num_res_blocks):
            res(f"down.{lvl}.block.{j}", block_in, block_out)
            block_in = block_out
        if lvl != len(c.ch_mult) - 1:
            conv(f"down.{lvl}.downsample.conv", block_in, block_in, 3)
    res("mid.block_1", block_in, block_in)
    gn("mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        conv(f"mid.attn_1.{n}", block_in, block_in, 1)
    res("mid.block_2", block_in, block_in)
    gn("norm_out", block_in)
    conv("conv_out", 2 * c.z_channels, block_in, 3)
    return ks


def tower_key_shapes(c: TowerCfg) -> dict:
    """Names/shapes of HF CLIPModel / SiglipModel vision-side state (prefix as in `.model`)."""
    ks: dict = {}
    D = c.hidden
    vm = "vision_model"

    def lin(name, o, i, bias=True):
        ks[f"{name}.weight"] = (o, i)
        if bias:
            ks[f"{name}.bias"] = (o,)

    def ln(name, d):
        ks[f"{name}.weight"] = (d,)
        ks[f"{name}.bias"] = (d,)

    if c.kind == "clip":
        ks[f"{vm}.embeddings.class_embedding"] = (D,)
        ks[f"{vm}.embeddings.patch_embedding.weight"] = (D, 3, c.patch, c.patch)
        ks[f"{vm}.embeddings.position_embedding.weight"] = (c.tokens, D)
        ln(f"{vm}.pre_layrnorm", D)
    else:
        ks[f"{vm}.embeddings.patch_embedding.weight"] = (D, 3, c.patch, c.patch)
        ks[f"{vm}.embeddings.patch_embedding.bias"] = (D,)
        ks[f"{vm}.embeddings.position_embedding.weight"] = (c.tokens, D)
    for i in range(c.layers):
        p = f"{vm}.encoder.layers.{i}"
        ln(f"{p}.layer_norm1", D)
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            lin(f"{p}.self_attn.{n}", D, D)
        ln(f"{p}.layer_norm2", D)
        lin(f"{p}.mlp.fc1", c.mlp, D)
        lin(f"{p}.mlp.fc2", D, c.mlp)
    ln(f"{vm}.post_layernorm", D)
    if c.kind == "clip":
        lin("visual_projection", c.proj_dim, D, bias=False)
    else:
        ks[f"{vm}.head.probe"] = (1, 1, D)
        ks[f"{vm}.head.attention.in_proj_weight"] = (3 * D, D)
        ks[f"{vm}.head.attention.in_proj_bias"] = (3 * D,)
        lin(f"{vm}.head.attention.out_proj", D, D)
        ln(f"{vm}.head.layernorm", D)
        lin(f"{vm}.head.mlp.fc1", c.mlp, D)
        lin(f"{vm}.head.mlp.fc2", D, c.mlp)
    return ks


def projector_key_shapes(prefix: str, in_dim: int, out_dim: int) -> dict:
    """nn.Sequential(LayerNorm, Linear, GELU, Linear) of R/clip_models/CLIP_bank.py:17-28."""
    return {f"{prefix}.0.weight": (in_dim,), f"{prefix}.0.bias": (in_dim,),
            f"{prefix}.1.weight": (out_dim, in_dim), f"{prefix}.1.bias": (out_dim,),
            f"{prefix}.3.weight": (out_dim, out_dim), f"{prefix}.3.bias": (out_dim,)}


def adapter_key_shapes(in_dim: int = 1024, out_dim: int = 4096) -> dict:
    """VisualPromptAdapter, R/train_OpenAICLIP_video_stage1.py:85-97."""
    mid = in_dim * 2  # nn.Linear(in_dim, in_dim * 2), train_OpenAICLIP_video_stage1.py:90
    return {"proj.0.weight": (mid, in_dim), "proj.0.bias": (mid,), "proj.2.weight": (out_dim, mid),
            "proj.2.bias": (out_dim,), "proj.3.weight": (out_dim,), "proj.3.bias": (out_dim,)}


# --------------------------------------------------------------------------------------
# tower: HF CLIPVisionTransformer / SiglipVisionTransformer
# --------------------------------------------------------------------------------------


def _act(name: str, x: torch.Tensor) -> torch.Tensor:
    if name == "quick_gelu":  # HF activations.QuickGELUActivation
        return x * torch.sigmoid(1.702 * x)
    if name == "gelu_tanh":
        return F.gelu(x, approximate="tanh")
    if name == "gelu":
        return F.gelu(x)
    raise ValueError(name)


def _mha(x, wq, bq, wk, bk, wv, bv, wo, bo, heads, q_in=None):
    """softmax(q k^T / sqrt(d)) v with separate projections (HF modeling_clip.py:300-336)."""
    B, L, D = x.shape
    qx = x if q_in is None else q_in
    Lq = qx.shape[1]
    d = D // heads
    q = F.linear(qx, wq, bq).view(B, Lq, heads, d).transpose(1, 2)
    k = F.linear(x, wk, bk).view(B, L, heads, d).transpose(1, 2)
    v = F.linear(x, wv, bv).view(B, L, heads, d).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * (d ** -0.5)
    o = torch.softmax(s, dim=-1) @ v
    return F.linear(o.transpose(1, 2).reshape(B, Lq, D), wo, bo)


def tower_forward(sd: dict, pixel_values: torch.Tensor, c: TowerCfg, lora: dict | None = None):
    """Returns (last_hidden_state, pooler_output).

    clip  : HF modeling_clip.py:202-218 (embeddings), :363-385 (layer), :667-696 (transformer);
            last_hidden_state is NOT post-layernormed, pooler = post_layernorm(h[:,0]).
    siglip: HF modeling_siglip.py:586-654; post_layernorm on all tokens, MAP pooling head.
    `lora` maps a linear's weight key -> (A [r,in], B [out,r], scaling) (peft restatement, R/
    train_SigLIP_stage2_all.py:134-142: y = Wx + b + scaling * B(A(x)); dropout taken as identity).
    """
    vm = "vision_model"
    D, p = c.hidden, c.patch

    def linear(x, name):
        y = F.linear(x, sd[f"{name}.weight"], sd.get(f"{name}.bias"))
        if lora and f"{name}.weight" in lora:
            A, Bm, sc = lora[f"{name}.weight"]
            y = y + sc * F.linear(F.linear(x, A), Bm)
        return y

    x = pixel_values.to(sd[f"{vm}.embeddings.patch_embedding.weight"].dtype)
    pe = F.conv2d(x, sd[f"{vm}.embeddings.patch_embedding.weight"],
                  sd.get(f"{vm}.embeddings.patch_embedding.bias"), stride=p)
    h = pe.flatten(2).transpose(1, 2)  # [B, P, D], row-major over the patch grid
    if c.kind == "clip":
        cls = sd[f"{vm}.embeddings.class_embedding"].expand(h.shape[0], 1, D)
        h = torch.cat([cls, h], dim=1)
    h = h + sd[f"{vm}.embeddings.position_embedding.weight"][None]
    if c.kind == "clip":
        h = F.layer_norm(h, (D,), sd[f"{vm}.pre_layrnorm.weight"], sd[f"{vm}.pre_layrnorm.bias"], c.eps)
    for i in range(c.layers):
        q = f"{vm}.encoder.layers.{i}"
        r = h
        y = F.layer_norm(h, (D,), sd[f"{q}.layer_norm1.weight"], sd[f"{q}.layer_norm1.bias"], c.eps)
        B_, L, _ = y.shape
        d = D // c.heads
        qq = linear(y, f"{q}.self_attn.q_proj").view(B_, L, c.heads, d).transpose(1, 2)
        kk = linear(y, f"{q}.self_attn.k_proj").view(B_, L, c.heads, d).transpose(1, 2)
        vv = linear(y, f"{q}.self_attn.v_proj").view(B_, L, c.heads, d).transpose(1, 2)
        s = (qq @ kk.transpose(-1, -2)) * (d ** -0.5)
        o = (torch.softmax(s, dim=-1) @ vv).transpose(1, 2).reshape(B_, L, D)
        h = r + linear(o, f"{q}.self_attn.out_proj")
        r = h
        y = F.layer_norm(h, (D,), sd[f"{q}.layer_norm2.weight"], sd[f"{q}.layer_norm2.bias"], c.eps)
        y = linear(_act(c.act, linear(y, f"{q}.mlp.fc1")), f"{q}.mlp.fc2")
        h = r + y
    if c.kind == "clip":
        pooled = F.layer_norm(h[:, 0], (D,), sd[f"{vm}.post_layernorm.weight"], sd[f"{vm}.post_layernorm.bias"], c.eps)
        return h, pooled
    h = F.layer_norm(h, (D,), sd[f"{vm}.post_layernorm.weight"], sd[f"{vm}.post_layernorm.bias"], c.eps)
    # MAP head: probe attends over tokens (nn.MultiheadAttention, packed in_proj), x + MLP(LN(x))
    W, bvec = sd[f"{vm}.head.attention.in_proj_weight"], sd[f"{vm}.head.attention.in_proj_bias"]
    probe = sd[f"{vm}.head.probe"].expand(h.shape[0], 1, D)
    a = _mha(h, W[:D], bvec[:D], W[D:2 * D], bvec[D:2 * D], W[2 * D:], bvec[2 * D:],
             sd[f"{vm}.head.attention.out_proj.weight"], sd[f"{vm}.head.attention.out_proj.bias"], c.heads,
             q_in=probe)
    y = F.layer_norm(a, (D,), sd[f"{vm}.head.layernorm.weight"], sd[f"{vm}.head.layernorm.bias"], c.eps)
    y = linear(_act(c.act, linear(y, f"{vm}.head.mlp.fc1")), f"{vm}.head.mlp.fc2")
    return h, (a + y)[:, 0]


def projector_forward(sd: dict, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """LayerNorm -> Linear -> GELU(erf) -> Linear (R/clip_models/CLIP_bank.py:17-28)."""
    d = x.shape[-1]
    y = F.layer_norm(x, (d,), sd[f"{prefix}.0.weight"], sd[f"{prefix}.0.bias"], 1e-5)
    y = F.gelu(F.linear(y, sd[f"{prefix}.1.weight"], sd[f"{prefix}.1.bias"]))
    return F.linear(y, sd[f"{prefix}.3.weight"], sd[f"{prefix}.3.bias"])


def visual_projection(sd_model: dict, pooled: torch.Tensor, lora=None) -> torch.Tensor:
    """visual_projection (no bias); wrapped by peft under target_modules='all-linear'
    (R/train_OpenAICLIP_use2frames_nextpredic_stage2_all.py:174-182)."""
    y = F.linear(pooled, sd_model["visual_projection.weight"])
    if lora and "visual_projection.weight" in lora:
        A, Bm, sc = lora["visual_projection.weight"]
        y = y + sc * F.linear(F.linear(pooled, A), Bm)
    return y


LORA_SIGLIP_TARGETS = ("k_proj", "v_proj", "q_proj", "out_proj", "fc1", "fc2")  # R/train_SigLIP_stage2_all.py:137


def lora_key_shapes(c: TowerCfg, r: int = 16, all_linear: bool = False) -> dict:
    """{"<linear>.lora_A": [r, in], "<linear>.lora_B": [out, r]} for every nn.Linear peft wraps on the vision side.
    The MAP head's out_proj sits inside nn.MultiheadAttention, whose forward reads .weight/.bias directly, so a
    LoRA pair there never takes part in the arithmetic (left out)."""
    vm, D = "vision_model", c.hidden
    ks = {}
    def add(name, out_f, in_f):
        ks[f"{name}.lora_A"] = [r, in_f]
        ks[f"{name}.lora_B"] = [out_f, r]
    for i in range(c.layers):
        q = f"{vm}.encoder.layers.{i}"
        for leaf in ("q_proj", "k_proj", "v_proj", "out_proj"):
            add(f"{q}.self_attn.{leaf}", D, D)
        add(f"{q}.mlp.fc1", c.mlp, D)
        add(f"{q}.mlp.fc2", D, c.mlp)
    if c.kind == "siglip":
        add(f"{vm}.head.mlp.fc1", c.mlp, D)
        add(f"{vm}.head.mlp.fc2", D, c.mlp)
    elif all_linear:
        add("visual_projection", c.proj_dim, D)
    return ks


def synth_lora(c: TowerCfg, seed: int, r: int = 16, alpha: float = 16.0, all_linear: bool = False, b_scale: float = 0.3):
    """-> ({weight key: (A, B, scaling)}, flat {name: tensor}).  B is random (not zero) so that the branch and both of
    its gradients are exercised (SURVEY.md 8d cfg 4)."""
    flat = synth_state_dict(lora_key_shapes(c, r, all_linear), seed)
    flat = {k: (v * b_scale if k.endswith("lora_B") else v) for k, v in flat.items()}
    names = sorted({k.rsplit(".", 1)[0] for k in flat})
    return {f"{n}.weight": (flat[f"{n}.lora_A"], flat[f"{n}.lora_B"], alpha / r) for n in names}, flat


def clip_wrapper_forward(sd_model: dict, sd_wrap: dict, images: torch.Tensor, c: TowerCfg, lora=None):
    """OpenAICLIP/MetaCLIP.forward (R/clip_models/CLIP_bank.py:32-40,115-122) and SigLIP.forward (:67-73).

    Returns (class_token, projection_clip, projection_t5[:, None, :])."""
    _, pooled = tower_forward(sd_model, images, c, lora)
    cls = visual_projection(sd_model, pooled, lora) if c.kind == "clip" else pooled
    return cls, projector_forward(sd_wrap, "project_clip", cls), projector_forward(sd_wrap, "project_t5", cls[:, None, :])


def adapter_forward(sd: dict, x: torch.Tensor) -> torch.Tensor:
    """VisualPromptAdapter: Linear -> SiLU -> Linear -> LayerNorm (R/train_OpenAICLIP_video_stage1.py:85-97)."""
    y = F.silu(F.linear(x, sd["proj.0.weight"], sd["proj.0.bias"]))
    y = F.linear(y, sd["proj.2.weight"], sd["proj.2.bias"])
    return F.layer_norm(y, (y.shape[-1],), sd["proj.3.weight"], sd["proj.3.bias"], 1e-5)


# --------------------------------------------------------------------------------------
# FLUX autoencoder encoder (R/src/flux/modules/autoencoder.py)
# --------------------------------------------------------------------------------------


def _gn_swish(x, w, b, swish=True):
    y = F.group_norm(x, 32, w, b, 1e-6)
    return y * torch.sigmoid(y) if swish else y


def _resnet(sd, name, x):  # autoencoder.py:69-82
    h = F.conv2d(_gn_swish(x, sd[f"{name}.norm1.weight"], sd[f"{name}.norm1.bias"]),
                 sd[f"{name}.conv1.weight"], sd[f"{name}.conv1.bias"], padding=1)
    h = F.conv2d(_gn_swish(h, sd[f"{name}.norm2.weight"], sd[f"{name}.norm2.bias"]),
                 sd[f"{name}.conv2.weight"], sd[f"{name}.conv2.bias"], padding=1)
    if f"{name}.nin_shortcut.weight" in sd:
        x = F.conv2d(x, sd[f"{name}.nin_shortcut.weight"], sd[f"{name}.nin_shortcut.bias"])
    return x + h


def ae_encoder_forward(sd: dict, x: torch.Tensor, c: AECfg) -> torch.Tensor:
    """Encoder.forward (autoencoder.py:159-180) -> [B, 2*z, H/8, W/8] moments."""
    h = F.conv2d(x, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)
    n_lvl = len(c.ch_mult)
    for lvl in range(n_lvl):
        for j in range(c.num_res_blocks):
            h = _resnet(sd, f"down.{lvl}.block.{j}", h)
        if lvl != n_lvl - 1:  # Downsample: zero-pad right/bottom by one, 3x3 stride 2 (autoencoder.py:85-95)
            h = F.conv2d(F.pad(h, (0, 1, 0, 1)), sd[f"down.{lvl}.downsample.conv.weight"],
                         sd[f"down.{lvl}.downsample.conv.bias"], stride=2)
    h = _resnet(sd, "mid.block_1", h)
    # AttnBlock (autoencoder.py:37-55): single head over H*W tokens, scale C^-0.5, residual
    y = _gn_swish(h, sd["mid.attn_1.norm.weight"], sd["mid.attn_1.norm.bias"], swish=False)
    q = F.conv2d(y, sd["mid.attn_1.q.weight"], sd["mid.attn_1.q.bias"])
    k = F.conv2d(y, sd["mid.attn_1.k.weight"], sd["mid.attn_1.k.bias"])
    v = F.conv2d(y, sd["mid.attn_1.v.weight"], sd["mid.attn_1.v.bias"])
    B, C, Hh, Ww = q.shape
    qf, kf, vf = (t.flatten(2).transpose(1, 2) for t in (q, k, v))  # [B, HW, C]
    a = torch.softmax((qf @ kf.transpose(1, 2)) * (C ** -0.5), dim=-1) @ vf
    a = a.transpose(1, 2).reshape(B, C, Hh, Ww)
    h = h + F.conv2d(a, sd["mid.attn_1.proj_out.weight"], sd["mid.attn_1.proj_out.bias"])
    h = _resnet(sd, "mid.block_2", h)
    h = _gn_swish(h, sd["norm_out.weight"], sd["norm_out.bias"])
    return F.conv2d(h, sd["conv_out.weight"], sd["conv_out.bias"], padding=1)


def ae_encode(sd: dict, x: torch.Tensor, c: AECfg, noise: torch.Tensor) -> torch.Tensor:
    """AutoEncoder.encode (autoencoder.py:302-305) with DiagonalGaussian sampling (:268-274).
    `noise` is the randn_like(mean) draw, supplied by the caller so RNG order stays the reference's."""
    mean, logvar = torch.chunk(ae_encoder_forward(sd, x, c), 2, dim=1)
    z = mean + torch.exp(0.5 * logvar) * noise
    return c.scale_factor * (z - c.shift_factor)


def _mid_attn(sd, h):  # AttnBlock (autoencoder.py:37-55)
    y = _gn_swish(h, sd["mid.attn_1.norm.weight"], sd["mid.attn_1.norm.bias"], swish=False)
    q = F.conv2d(y, sd["mid.attn_1.q.weight"], sd["mid.attn_1.q.bias"])
    k = F.conv2d(y, sd["mid.attn_1.k.weight"], sd["mid.attn_1.k.bias"])
    v = F.conv2d(y, sd["mid.attn_1.v.weight"], sd["mid.attn_1.v.bias"])
    B, C, Hh, Ww = q.shape
    qf, kf, vf = (t.flatten(2).transpose(1, 2) for t in (q, k, v))
    a = torch.softmax((qf @ kf.transpose(1, 2)) * (C ** -0.5), dim=-1) @ vf
    a = a.transpose(1, 2).reshape(B, C, Hh, Ww)
    return h + F.conv2d(a, sd["mid.attn_1.proj_out.weight"], sd["mid.attn_1.proj_out.bias"])


def ae_decoder_key_shapes(c: AECfg, out_ch: int = 3) -> dict:
    """Names/shapes of AutoEncoder.decoder.state_dict() (R/src/flux/modules/autoencoder.py:183-234)."""
    ks: dict = {}

    def conv(name, o, i, k):
        ks[f"{name}.weight"] = (o, i, k, k)
        ks[f"{name}.bias"] = (o,)

    def gn(name, ch):
        ks[f"{name}.weight"] = (ch,)
        ks[f"{name}.bias"] = (ch,)

    def res(name, i, o):
        gn(f"{name}.norm1", i)
        conv(f"{name}.conv1", o, i, 3)
        gn(f"{name}.norm2", o)
        conv(f"{name}.conv2", o, o, 3)
        if i != o:
            conv(f"{name}.nin_shortcut", o, i, 1)

    n_lvl = len(c.ch_mult)
    block_in = c.ch * c.ch_mult[n_lvl - 1]
    conv("conv_in", block_in, c.z_channels, 3)
    res("mid.block_1", block_in, block_in)
    gn("mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        conv(f"mid.attn_1.{n}", block_in, block_in, 1)
    res("mid.block_2", block_in, block_in)
    for lvl in reversed(range(n_lvl)):
        block_out = c.ch * c.ch_mult[lvl]
        for j in range(c.num_res_blocks + 1):
            res(f"up.{lvl}.block.{j}", block_in, block_out)
            block_in = block_out
        if lvl != 0:
            conv(f"up.{lvl}.upsample.conv", block_in, block_in, 3)
    gn("norm_out", block_in)
    conv("conv_out", out_ch, block_in, 3)
    return ks


def ae_decoder_forward(sd: dict, z: torch.Tensor, c: AECfg) -> torch.Tensor:
    """Decoder.forward (autoencoder.py:236-259): conv_in -> mid -> up levels (num_res_blocks + 1 ResnetBlocks, then
    nearest 2x Upsample + 3x3 conv, :98-106) -> GroupNorm + swish -> conv_out."""
    h = F.conv2d(z, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)
    h = _resnet(sd, "mid.block_1", h)
    h = _mid_attn(sd, h)
    h = _resnet(sd, "mid.block_2", h)
    for lvl in reversed(range(len(c.ch_mult))):
        for j in range(c.num_res_blocks + 1):
            h = _resnet(sd, f"up.{lvl}.block.{j}", h)
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, sd[f"up.{lvl}.upsample.conv.weight"], sd[f"up.{lvl}.upsample.conv.bias"], padding=1)
    h = _gn_swish(h, sd["norm_out.weight"], sd["norm_out.bias"])
    return F.conv2d(h, sd["conv_out.weight"], sd["conv_out.bias"], padding=1)


def ae_decode(sd: dict, z: torch.Tensor, c: AECfg) -> torch.Tensor:
    """AutoEncoder.decode (autoencoder.py:307-309): z / scale_factor + shift_factor -> decoder."""
    return ae_decoder_forward(sd, z / c.scale_factor + c.shift_factor, c)


# --------------------------------------------------------------------------------------
# conditioning packer / ids
# --------------------------------------------------------------------------------------


def patchify(z: torch.Tensor) -> torch.Tensor:
    """einops 'b c (h ph) (w pw) -> b (h w) (c ph pw)', ph=pw=2 (R/clip_models/sampling.py:26)."""
    B, C, H, W = z.shape
    return z.view(B, C, H // 2, 2, W // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(B, (H // 2) * (W // 2), C * 4)


def make_img_ids(bs: int, h2: int, w2: int, t: float = 0.0) -> torch.Tensor:
    """(t, row, col) ids, [bs, h2*w2, 3] (R/clip_models/sampling.py:30-33; t=0 in image mode)."""
    ids = torch.zeros(h2, w2, 3)
    ids[..., 0] = t
    ids[..., 1] = torch.arange(h2)[:, None]
    ids[..., 2] = torch.arange(w2)[None, :]
    return ids.reshape(1, h2 * w2, 3).repeat(bs, 1, 1)


def create_spatio_temporal_ids(bs: int, t: int, h: int, w: int) -> torch.Tensor:
    """R/train_OpenAICLIP_video_stage1.py:128-151 -- ids[...,0]=time, [1]=row, [2]=col; [bs, h*w, 3]."""
    return make_img_ids(bs, h, w, float(t))


def build_windows_with_mask(frames: torch.Tensor, frame_mask: torch.Tensor, window_cond: int = 3, stride: int = 1,
                            max_windows: int = 8, rng=None):
    """R/train_OpenAICLIP_sliding_windows_nextpredic_stage1.py:149-204.
    frames [B,T,3,H,W], mask [B,T] -> (list of window_cond cond tensors, target) each [bs_eff,3,H,W],
    plus the number of windows per video.  rng: a `random.Random` used only when a video yields more than
    max_windows windows (random.sample then sorted, as in the reference)."""
    B = frames.shape[0]
    conds = [[] for _ in range(window_cond)]
    tgt, counts = [], []
    for b in range(B):
        Ti = int(frame_mask[b].sum().item())
        starts = list(range(0, Ti - window_cond, stride))
        if max_windows is not None and len(starts) > max_windows:
            starts = sorted((rng or __import__("random")).sample(starts, max_windows))
        counts.append(len(starts))
        for s in starts:
            for j in range(window_cond):
                conds[j].append(frames[b, s + j])
            tgt.append(frames[b, s + window_cond])
    if not tgt:
        return None, None, counts
    return [torch.stack(c, 0) for c in conds], torch.stack(tgt, 0), counts


# --------------------------------------------------------------------------------------
# DiT (R/src/flux/model.py, modules/layers.py, math.py)
# --------------------------------------------------------------------------------------


def timestep_embedding(t: torch.Tensor, dim: int = 256, max_period: int = 10000, time_factor: float = 1000.0):
    """layers.py:28-49 (the multiply happens in t's own dtype BEFORE .float(): bf16 t -> bf16(1000 t))."""
    t = time_factor * t
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    return emb.to(t) if torch.is_floating_point(t) else emb


def rope_table(ids: torch.Tensor, axes_dim, theta: int) -> torch.Tensor:
    """EmbedND + rope (layers.py:18-25, math.py:15-22): [B, 1, L, sum(axes)/2, 2, 2] fp32 built in float64."""
    outs = []
    for i, dim in enumerate(axes_dim):
        pos = ids[..., i]
        scale = torch.arange(0, dim, 2, dtype=torch.float64, device=ids.device) / dim
        omega = 1.0 / (theta ** scale)
        ang = torch.einsum("...n,d->...nd", pos.to(torch.float64) if pos.dtype != torch.float64 else pos, omega) \
            if False else torch.einsum("...n,d->...nd", pos, omega.to(pos.dtype) if False else omega)
        outs.append(torch.stack([torch.cos(ang), -torch.sin(ang), torch.sin(ang), torch.cos(ang)], dim=-1)
                    .reshape(*ang.shape, 2, 2).float())
    return torch.cat(outs, dim=-3).unsqueeze(1)


def apply_rope(xq, xk, pe):
    """math.py:25-30: pairs (x[2j], x[2j+1]) rotated in fp32, cast back."""
    def rot(x):
        x_ = x.float().reshape(*x.shape[:-1], -1, 1, 2)
        return (pe[..., 0] * x_[..., 0] + pe[..., 1] * x_[..., 1]).reshape(*x.shape).type_as(x)
    return rot(xq), rot(xk)


def _rms(x, scale):  # layers.py:63-72
    xf = x.float()
    rr = torch.rsqrt(torch.mean(xf ** 2, dim=-1, keepdim=True) + 1e-6)
    return (xf * rr).to(x.dtype) * scale


def _attention(q, k, v, pe):  # math.py:6-12
    q, k = apply_rope(q, k, pe)
    x = F.scaled_dot_product_attention(q, k, v)
    B, H, L, D = x.shape
    return x.transpose(1, 2).reshape(B, L, H * D)


def _ln(x):
    return F.layer_norm(x, (x.shape[-1],), None, None, 1e-6)


def _lin(sd, name, x):
    return F.linear(x, sd[f"{name}.weight"], sd.get(f"{name}.bias"))


def _mlp_embed(sd, name, x):  # layers.py:52-60
    return _lin(sd, f"{name}.out_layer", F.silu(_lin(sd, f"{name}.in_layer", x)))


def _heads(qkv, H):
    B, L, _ = qkv.shape
    return qkv.view(B, L, 3, H, -1).permute(2, 0, 3, 1, 4)  # K B H L D


def flux_forward(sd: dict, c: FluxCfg, img, img_ids, txt, txt_ids, timesteps, y, guidance):
    """Flux.forward (model.py:137-228) with the live processors (layers.py:303-337, 485-501, 561-572)."""
    if img.ndim != 3 or txt.ndim != 3:
        raise ValueError("Input img and txt tensors must have 3 dimensions.")
    H = c.num_heads
    img = _lin(sd, "img_in", img)
    vec = _mlp_embed(sd, "time_in", timestep_embedding(timesteps, 256))
    if c.guidance_embed:
        if guidance is None:
            raise ValueError("Didn't get guidance strength for guidance distilled model.")
        vec = vec + _mlp_embed(sd, "guidance_in", timestep_embedding(guidance, 256))
    vec = vec + _mlp_embed(sd, "vector_in", y)
    txt = _lin(sd, "txt_in", txt)
    pe = rope_table(torch.cat((txt_ids, img_ids), dim=1), c.axes_dim, c.theta)
    n_txt = txt.shape[1]
    for b in range(c.depth):
        p = f"double_blocks.{b}"
        im = _lin(sd, f"{p}.img_mod.lin", F.silu(vec))[:, None, :].chunk(6, dim=-1)
        tm = _lin(sd, f"{p}.txt_mod.lin", F.silu(vec))[:, None, :].chunk(6, dim=-1)
        iq, ik, iv = _heads(_lin(sd, f"{p}.img_attn.qkv", (1 + im[1]) * _ln(img) + im[0]), H)
        iq = _rms(iq, sd[f"{p}.img_attn.norm.query_norm.scale"]).to(iv)
        ik = _rms(ik, sd[f"{p}.img_attn.norm.key_norm.scale"]).to(iv)
        tq, tk, tv = _heads(_lin(sd, f"{p}.txt_attn.qkv", (1 + tm[1]) * _ln(txt) + tm[0]), H)
        tq = _rms(tq, sd[f"{p}.txt_attn.norm.query_norm.scale"]).to(tv)
        tk = _rms(tk, sd[f"{p}.txt_attn.norm.key_norm.scale"]).to(tv)
        a = _attention(torch.cat((tq, iq), 2), torch.cat((tk, ik), 2), torch.cat((tv, iv), 2), pe)
        ta, ia = a[:, :n_txt], a[:, n_txt:]
        img = img + im[2] * _lin(sd, f"{p}.img_attn.proj", ia)
        img = img + im[5] * _lin(sd, f"{p}.img_mlp.2", F.gelu(_lin(sd, f"{p}.img_mlp.0", (1 + im[4]) * _ln(img) + im[3]),
                                                              approximate="tanh"))
        txt = txt + tm[2] * _lin(sd, f"{p}.txt_attn.proj", ta)
        txt = txt + tm[5] * _lin(sd, f"{p}.txt_mlp.2", F.gelu(_lin(sd, f"{p}.txt_mlp.0", (1 + tm[4]) * _ln(txt) + tm[3]),
                                                              approximate="tanh"))
    x = torch.cat((txt, img), 1)
    hid = c.hidden_size
    for b in range(c.depth_single_blocks):
        p = f"single_blocks.{b}"
        m = _lin(sd, f"{p}.modulation.lin", F.silu(vec))[:, None, :].chunk(3, dim=-1)
        l1 = _lin(sd, f"{p}.linear1", (1 + m[1]) * _ln(x) + m[0])
        qkv, mlp = l1[..., : 3 * hid], l1[..., 3 * hid:]
        q, k, v = _heads(qkv, H)
        q = _rms(q, sd[f"{p}.norm.query_norm.scale"]).to(v)
        k = _rms(k, sd[f"{p}.norm.key_norm.scale"]).to(v)
        a = _attention(q, k, v, pe)
        x = x + m[2] * _lin(sd, f"{p}.linear2", torch.cat((a, F.gelu(mlp, approximate="tanh")), 2))
    x = x[:, n_txt:]
    shift, scale = _lin(sd, "final_layer.adaLN_modulation.1", F.silu(vec)).chunk(2, dim=1)
    x = (1 + scale[:, None, :]) * _ln(x) + shift[:, None, :]
    return _lin(sd, "final_layer.linear", x)


# --------------------------------------------------------------------------------------
# the step (R/train_SigLIP_stage1.py:238-270 and the video variants)
# --------------------------------------------------------------------------------------


@dataclass
class StepOut:
    loss: torch.Tensor
    pred: torch.Tensor
    x_t: torch.Tensor
    x_1: torch.Tensor
    class_token: torch.Tensor
    vec: torch.Tensor
    txt: torch.Tensor
    extras: dict = field(default_factory=dict)


def fm_interp(x_1, x_0, t):
    return (1 - t[:, None, None]) * x_1 + t[:, None, None] * x_0


def stage1_image_step(sd_tower, sd_wrap, sd_dit, sd_ae, img01, tcfg: TowerCfg, fcfg: FluxCfg, acfg: AECfg,
                      clip_mean, clip_std, ae_noise, t, x_0, dit_dtype=torch.float32, lora=None) -> StepOut:
    """One image-mode micro-step, forward only (autograd on the inputs gives the oracle gradients).

    img01: [B,3,S,S] in [0,1].  ae_noise/t/x_0 are the three RNG draws in the reference's order
    (R/train_SigLIP_stage1.py:243 -> :248 -> :249), supplied by the caller.
    dit_dtype=torch.bfloat16 reproduces the reference's weight_dtype casts (:255-261)."""
    dev = img01.device
    mean = torch.as_tensor(clip_mean, dtype=torch.float32, device=dev).view(1, -1, 1, 1)
    std = torch.as_tensor(clip_std, dtype=torch.float32, device=dev).view(1, -1, 1, 1)
    with torch.no_grad():
        x_1 = ae_encode(sd_ae, ((img01 - 0.5) / 0.5).float(), acfg, ae_noise)
    cls, vec, txt = clip_wrapper_forward(sd_tower, sd_wrap, (img01 - mean) / std, tcfg, lora)
    B, _, h, w = x_1.shape
    img_ids = make_img_ids(B, h // 2, w // 2).to(dev)
    txt_ids = torch.zeros(B, txt.shape[1], 3, device=dev)
    x_1 = patchify(x_1)
    x_t = fm_interp(x_1, x_0, t)
    wd = dit_dtype
    sd_d = sd_dit if wd == torch.float32 else {k: v.to(wd) for k, v in sd_dit.items()}
    pred = flux_forward(sd_d, fcfg, x_t.to(wd), img_ids.to(wd), txt.to(wd), txt_ids.to(wd), t.to(wd), vec.to(wd),
                        torch.full((B,), 4.0, device=dev, dtype=wd))
    loss = F.mse_loss(pred.float(), (x_0 - x_1).float(), reduction="mean")
    return StepOut(loss, pred, x_t, x_1, cls, vec, txt)


def stage1_video_step(sd_tower, sd_adapter, sd_dit, sd_ae, cond_frames, target, tcfg: TowerCfg, fcfg: FluxCfg,
                      acfg: AECfg, clip_mean, clip_std, cond_times, target_time, ae_noise, t, x_0,
                      dit_dtype=torch.float32, lora=None) -> StepOut:
    """One video-mode micro-step (R/train_OpenAICLIP_video_stage1.py:355-452 and the nextpredic / use2frames /
    sliding-window variants): per-patch tokens of the conditioning frames -> VisualPromptAdapter ->
    the DiT's txt stream with (time,row,col) ids; vec = mean over frames of visual_projection(pooler_output);
    the target frame's latent gets time index `target_time`."""
    dev = target.device
    mean = torch.as_tensor(clip_mean, dtype=torch.float32, device=dev).view(1, -1, 1, 1)
    std = torch.as_tensor(clip_std, dtype=torch.float32, device=dev).view(1, -1, 1, 1)
    with torch.no_grad():
        x_1 = ae_encode(sd_ae, ((target - 0.5) / 0.5).float(), acfg, ae_noise)
    patches, vecs = [], []
    for f in cond_frames:
        lhs, pooled = tower_forward(sd_tower, (f - mean) / std, tcfg, lora)
        patches.append(lhs[:, 1:, :])
        vecs.append(visual_projection(sd_tower, pooled, lora))
    vec = sum(vecs) / len(vecs)
    txt = adapter_forward(sd_adapter, torch.cat(patches, dim=1))
    B, _, h, w = x_1.shape
    g = int(round(patches[0].shape[1] ** 0.5))
    txt_ids = torch.cat([create_spatio_temporal_ids(B, tt, g, g) for tt in cond_times], dim=1).to(dev)
    img_ids = make_img_ids(B, h // 2, w // 2, float(target_time)).to(dev)
    x_1 = patchify(x_1)
    x_t = fm_interp(x_1, x_0, t)
    wd = dit_dtype
    sd_d = sd_dit if wd == torch.float32 else {k: v.to(wd) for k, v in sd_dit.items()}
    pred = flux_forward(sd_d, fcfg, x_t.to(wd), img_ids.to(wd), txt.to(wd), txt_ids.to(wd), t.to(wd), vec.to(wd),
                        torch.full((B,), 4.0, device=dev, dtype=wd))
    loss = F.mse_loss(pred.float(), (x_0 - x_1).float(), reduction="mean")
    return StepOut(loss, pred, x_t, x_1, vecs[0], vec, txt, extras=dict(txt_ids=txt_ids, img_ids=img_ids))


def lora_merge(w: torch.Tensor, A: torch.Tensor, Bm: torch.Tensor, scaling: float) -> torch.Tensor:
    """peft merge_and_unload: W += scaling * B @ A (R/train_SigLIP_stage2_all.py:307-311)."""
    return w + scaling * (Bm @ A)


# --------------------------------------------------------------------------------------
# flow sampler (R/src/flux/sampling.py)
# --------------------------------------------------------------------------------------


def get_schedule(num_steps: int, image_seq_len: int, base_shift: float = 0.5, max_shift: float = 1.15, shift: bool = True):
    """R/src/flux/sampling.py:66-94: linspace(1, 0, n+1), shifted by mu(image_seq_len) through
    t -> e^mu / (e^mu + (1/t - 1))."""
    ts = torch.linspace(1, 0, num_steps + 1)
    if shift:
        m = (max_shift - base_shift) / (4096 - 256)
        mu = m * image_seq_len + (base_shift - m * 256)
        ts = math.exp(mu) / (math.exp(mu) + (1 / ts - 1) ** 1.0)
    return ts.tolist()


def denoise(sd: dict, c: FluxCfg, img, img_ids, txt, txt_ids, vec, neg_txt, neg_txt_ids, neg_vec, timesteps, guidance=4.0,
            true_gs=1.0, timestep_to_start_cfg=0):
    """R/src/flux/sampling.py:97-150: Euler steps x += (t_prev - t_curr) * v with the true-CFG mix
    v = neg + true_gs * (pred - neg) from step `timestep_to_start_cfg` on."""
    g = torch.full((img.shape[0],), guidance, dtype=img.dtype)
    for i, (tc, tp) in enumerate(zip(timesteps[:-1], timesteps[1:])):
        tv = torch.full((img.shape[0],), tc, dtype=img.dtype)
        pred = flux_forward(sd, c, img, img_ids, txt, txt_ids, tv, vec, g)
        if i >= timestep_to_start_cfg:
            neg = flux_forward(sd, c, img, img_ids, neg_txt, neg_txt_ids, tv, neg_vec, g)
            pred = neg + true_gs * (pred - neg)
        img = img + (tp - tc) * pred
    return img


def unpack(x: torch.Tensor, height: int, width: int) -> torch.Tensor:
    """R/src/flux/sampling.py:234-242: b (h w) (c ph pw) -> b c (h ph) (w pw)."""
    h, w = math.ceil(height / 16), math.ceil(width / 16)
    b, _, d = x.shape
    return x.view(b, h, w, d // 4, 2, 2).permute(0, 3, 1, 4, 2, 5).reshape(b, d // 4, 2 * h, 2 * w)
