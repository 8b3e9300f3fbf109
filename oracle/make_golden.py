"""Pin the oracle against the REAL reference and mint the golden fixtures (build container only).

    python oracle/make_golden.py [--full]

Imports /root/reference/Continuous (read-only) with the two shims from SURVEY.md section 8c
(a stub for the unused `import clip`; HF `from_pretrained` replaced by random-init of the same
architecture), loads deterministic synthetic weights (oracle.synth_state_dict) into the
reference's own modules, runs the reference forward/backward, asserts that
oracle/genhancer_oracle.py reproduces it, and writes small fixtures to tests/golden/*.pt.
Weights are NOT stored: a fixture holds (config, key->shape, seed) and the tests regenerate
the identical tensors.  /root/reference does not exist on the GPU box, so nothing at test or
bench time imports it.
"""
from __future__ import annotations

import argparse
import os
import sys
import types

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = "/root/reference/Continuous"
sys.path.insert(0, REF)
sys.modules.setdefault("clip", types.ModuleType("clip"))  # CLIP_bank.py:1 imports it, never uses it

from einops import rearrange  # noqa: E402
from transformers import CLIPConfig, CLIPModel, SiglipConfig, SiglipModel  # noqa: E402

import clip_models.CLIP_bank as bank  # noqa: E402  (reference)
from clip_models.sampling import prepare_clip  # noqa: E402  (reference)
from src.flux.model import Flux, FluxParams  # noqa: E402  (reference)
from src.flux.modules.autoencoder import AutoEncoder, AutoEncoderParams  # noqa: E402  (reference)

from oracle import genhancer_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
OPENAI_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_STD = (0.26862954, 0.26130258, 0.27577711)


def close(a, b, tol, what):
    err = (a.double() - b.double()).abs().max().item()
    ref = b.double().abs().max().item()
    ok = err <= tol * max(ref, 1e-6)
    print(f"  {'ok ' if ok else 'BAD'} {what}: max|d|={err:.3e} (ref max {ref:.3e})")
    assert ok, what
    return err


def ref_clip_model(tc: O.TowerCfg):
    if tc.kind == "clip":
        vis = dict(hidden_size=tc.hidden, intermediate_size=tc.mlp, num_hidden_layers=tc.layers,
                   num_attention_heads=tc.heads, image_size=tc.image_size, patch_size=tc.patch,
                   hidden_act="quick_gelu", projection_dim=tc.proj_dim, layer_norm_eps=tc.eps)
        txt = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=1, num_attention_heads=2,
                   projection_dim=tc.proj_dim, vocab_size=100, max_position_embeddings=8)
        return CLIPModel(CLIPConfig(text_config=txt, vision_config=vis, projection_dim=tc.proj_dim))
    vis = dict(hidden_size=tc.hidden, intermediate_size=tc.mlp, num_hidden_layers=tc.layers,
               num_attention_heads=tc.heads, image_size=tc.image_size, patch_size=tc.patch,
               hidden_act="gelu_pytorch_tanh", layer_norm_eps=tc.eps)
    txt = dict(hidden_size=tc.hidden, intermediate_size=128, num_hidden_layers=1, num_attention_heads=2,
               vocab_size=100, max_position_embeddings=8)
    return SiglipModel(SiglipConfig(text_config=txt, vision_config=vis))


def load_subset(module, sd):
    missing, unexpected = module.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    return missing


def build_reference_wrapper(tc: O.TowerCfg, clip_dim: int, t5_dim: int, seed: int, metaclip: str | None = None):
    """The reference's own OpenAICLIP / SigLIP / MetaCLIP wrapper around a random-init HF model, then our synthetic
    weights.  ``metaclip`` = "large" | "huge" takes the MetaCLIP class (CLIP_bank.py:76-122) instead of OpenAICLIP."""
    model = ref_clip_model(tc)

    class Cfg:
        clip_image_size = 224 if tc.kind == "clip" else 384  # only selects which from_pretrained path is taken
        pass
    Cfg.clip_dim, Cfg.t5_dim = clip_dim, t5_dim
    shim = type("Shim", (), {"from_pretrained": staticmethod(lambda *a, **k: model)})
    if tc.kind == "clip" and metaclip is not None:
        bank.CLIPModel = shim
        Cfg.clip_type = metaclip
        wrap = bank.MetaCLIP(Cfg())
    elif tc.kind == "clip":
        bank.CLIPModel = shim
        wrap = bank.OpenAICLIP(Cfg())
    else:
        bank.SiglipModel = shim
        wrap = bank.SigLIP(Cfg())
    feat = tc.feat_dim
    # reference hard-codes LayerNorm(768)/(1152); rebuild the projectors for reduced test sizes
    if wrap.project_clip[0].normalized_shape[0] != feat:
        import torch.nn as nn
        wrap.project_clip = nn.Sequential(nn.LayerNorm(feat), nn.Linear(feat, clip_dim), nn.GELU(), nn.Linear(clip_dim, clip_dim))
        wrap.project_t5 = nn.Sequential(nn.LayerNorm(feat), nn.Linear(feat, t5_dim), nn.GELU(), nn.Linear(t5_dim, t5_dim))
    ks_t = O.tower_key_shapes(tc)
    sd_t = O.synth_state_dict(ks_t, seed)
    load_subset(wrap.model, sd_t)
    ks_w = {**O.projector_key_shapes("project_clip", feat, clip_dim), **O.projector_key_shapes("project_t5", feat, t5_dim)}
    sd_w = O.synth_state_dict(ks_w, seed + 1)
    load_subset(wrap, sd_w)
    wrap.float()
    return wrap, sd_t, sd_w, ks_t, ks_w


def golden_tower(name: str, tc: O.TowerCfg, clip_dim=48, t5_dim=96, B=2, seed=11, metaclip: str | None = None):
    print(f"[tower:{name}]")
    wrap, sd_t, sd_w, ks_t, ks_w = build_reference_wrapper(tc, clip_dim, t5_dim, seed, metaclip)
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(B, 3, tc.image_size, tc.image_size, generator=g)
    mean = torch.tensor(OPENAI_MEAN if tc.kind == "clip" else (0.5,) * 3).view(1, 3, 1, 1)
    std = torch.tensor(OPENAI_STD if tc.kind == "clip" else (0.5,) * 3).view(1, 3, 1, 1)
    x = (img - mean) / std
    for p_ in wrap.parameters():
        p_.requires_grad_(False)
    for n_, p_ in wrap.named_parameters():
        if "project_clip" in n_ or "project_t5" in n_:
            p_.requires_grad_(True)
    out = wrap.model.vision_model(x, output_hidden_states=True)
    cls, pc, pt5 = wrap(x)
    (pc.square().mean() + pt5.square().mean()).backward()
    # oracle on the same weights
    sdw = {k: v.clone().requires_grad_(True) for k, v in sd_w.items()}
    lhs, pooled = O.tower_forward(sd_t, x, tc)
    ocls, opc, opt5 = O.clip_wrapper_forward(sd_t, sdw, x, tc)
    (opc.square().mean() + opt5.square().mean()).backward()
    close(lhs, out.last_hidden_state, 2e-5, "last_hidden_state")
    close(pooled, out.pooler_output, 2e-5, "pooler_output")
    close(ocls, cls, 2e-5, "class_token")
    close(opc, pc, 2e-5, "projection_clip")
    close(opt5, pt5, 2e-5, "projection_t5")
    close(sdw["project_t5.1.weight"].grad, wrap.project_t5[1].weight.grad, 1e-4, "grad project_t5.1.weight")
    fx = dict(kind="tower", cfg=tc.__dict__, clip_dim=clip_dim, t5_dim=t5_dim, seed=seed, key_shapes_tower=ks_t,
              key_shapes_wrap=ks_w, img=img, last_hidden_state=out.last_hidden_state.detach(),
              pooler_output=out.pooler_output.detach(), class_token=cls.detach(), projection_clip=pc.detach(),
              projection_t5=pt5.detach(), grad_project_t5_1_weight=wrap.project_t5[1].weight.grad.clone(),
              grad_project_clip_3_bias=wrap.project_clip[3].bias.grad.clone())
    torch.save(fx, os.path.join(GOLD, f"tower_{name}.pt"))


def golden_tower_lora(name: str, tc: O.TowerCfg, all_linear: bool, clip_dim=48, t5_dim=96, B=2, seed=61, r=16, alpha=16.0):
    """Stage-2 tower: the reference wrapper (HF model inside) with peft's LoRA layer restated as forward hooks on the
    wrapped nn.Linear modules (peft 0.14.0 is not installed: y = base(x) + (alpha/r) B(A(x)), dropout off,
    bias='lora_only' -> the wrapped linears' biases train; train_SigLIP_stage2_all.py:134-142).  Gradients of the LoRA
    pairs / biases / projectors from torch autograd over the HF modules are the golden values; the oracle's own LoRA
    path must agree."""
    print(f"[tower_lora:{name}]")
    wrap, sd_t, sd_w, ks_t, ks_w = build_reference_wrapper(tc, clip_dim, t5_dim, seed)
    lora, flat = O.synth_lora(tc, seed + 2, r, alpha, all_linear)
    for p_ in wrap.parameters():
        p_.requires_grad_(False)
    for n_, p_ in wrap.named_parameters():
        if "project_clip" in n_ or "project_t5" in n_:
            p_.requires_grad_(True)
    params = {}
    mods = dict(wrap.model.named_modules())
    for wkey, (A, Bm, sc) in lora.items():
        lin = mods[wkey[:-len(".weight")]]
        A_, B_ = A.clone().requires_grad_(True), Bm.clone().requires_grad_(True)
        params[wkey[:-len(".weight")]] = (A_, B_)
        lin.register_forward_hook(lambda m, inp, out, A_=A_, B_=B_, sc=sc: out + sc * F.linear(F.linear(inp[0], A_), B_))
        if lin.bias is not None:
            lin.bias.requires_grad_(True)
    if tc.kind == "siglip":  # peft also marks the MAP head's out_proj bias trainable (wrapped module, bias='lora_only')
        wrap.model.vision_model.head.attention.out_proj.bias.requires_grad_(True)
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(B, 3, tc.image_size, tc.image_size, generator=g)
    mean = torch.tensor(OPENAI_MEAN if tc.kind == "clip" else (0.5,) * 3).view(1, 3, 1, 1)
    std = torch.tensor(OPENAI_STD if tc.kind == "clip" else (0.5,) * 3).view(1, 3, 1, 1)
    x = (img - mean) / std
    out = wrap.model.vision_model(x, output_hidden_states=True)
    lhs_ref = out.last_hidden_state.detach().clone()
    pooled_ref = out.pooler_output.detach().clone()
    cls, pc, pt5 = wrap(x)
    out2 = wrap.model.vision_model(x, output_hidden_states=True)
    loss = pc.square().mean() + pt5.square().mean() + 0.1 * out2.last_hidden_state[:, 1:].square().mean()
    loss.backward()
    # the oracle's LoRA path on the same tensors
    lo_o = {k: (A.clone().requires_grad_(True), Bm.clone().requires_grad_(True), sc) for k, (A, Bm, sc) in lora.items()}
    sdt = {k: v.clone() for k, v in sd_t.items()}
    bias_keys = [f"{n}.bias" for n in params if f"{n}.bias" in sdt]
    if tc.kind == "siglip":
        bias_keys.append("vision_model.head.attention.out_proj.bias")
    for k in bias_keys:
        sdt[k].requires_grad_(True)
    sdw = {k: v.clone().requires_grad_(True) for k, v in sd_w.items()}
    ocls, opc, opt5 = O.clip_wrapper_forward(sdt, sdw, x, tc, lo_o)
    olhs, _ = O.tower_forward(sdt, x, tc, lo_o)
    (opc.square().mean() + opt5.square().mean() + 0.1 * olhs[:, 1:].square().mean()).backward()
    close(olhs, lhs_ref, 2e-5, "last_hidden_state (LoRA)")
    close(ocls, cls, 2e-5, "class_token (LoRA)")
    close(opt5, pt5, 2e-5, "projection_t5 (LoRA)")
    grads = {}
    for n, (A_, B_) in params.items():
        close(lo_o[f"{n}.weight"][0].grad, A_.grad, 2e-4, f"grad {n}.lora_A")
        close(lo_o[f"{n}.weight"][1].grad, B_.grad, 2e-4, f"grad {n}.lora_B")
        grads[f"{n}.lora_A"], grads[f"{n}.lora_B"] = A_.grad.clone(), B_.grad.clone()
    for k in bias_keys:
        gref = mods[k[:-len(".bias")]].bias.grad
        close(sdt[k].grad, gref, 2e-4, f"grad {k}")
        grads[k] = gref.clone()
    grads["project_t5.1.weight"] = wrap.project_t5[1].weight.grad.clone()
    grads["project_clip.3.bias"] = wrap.project_clip[3].bias.grad.clone()
    fx = dict(kind="tower_lora", cfg=tc.__dict__, clip_dim=clip_dim, t5_dim=t5_dim, seed=seed, key_shapes_tower=ks_t,
              key_shapes_wrap=ks_w, key_shapes_lora=O.lora_key_shapes(tc, r, all_linear), r=r, alpha=alpha,
              all_linear=all_linear, b_scale=0.3, img=img, last_hidden_state=lhs_ref, pooler_output=pooled_ref,
              class_token=cls.detach(), projection_clip=pc.detach(), projection_t5=pt5.detach(), loss=loss.detach(),
              grads=grads)
    torch.save(fx, os.path.join(GOLD, f"tower_lora_{name}.pt"))


def golden_prepare_clip(seed=81):
    """The reference's ``prepare_clip`` (clip_models/sampling.py:9-42) around its own OpenAICLIP wrapper: the dict the
    image-mode scripts feed the DiT with (img patchified, (0,row,col) img_ids, zero txt_ids, txt, vec)."""
    print("[prepare_clip]")
    tc = O.TowerCfg("clip", 128, 2, 2, 512, 56, 14, 64, 1e-5, "quick_gelu")
    wrap, sd_t, sd_w, ks_t, ks_w = build_reference_wrapper(tc, 48, 96, seed)
    g = torch.Generator().manual_seed(seed)
    B = 3
    original_img = torch.randn(B, 3, 56, 56, generator=g)          # already CLIP-normalised, as the scripts pass it
    latent = torch.randn(B, 16, 6, 10, generator=g)                # non-square on purpose: rows / cols must not swap
    with torch.no_grad():
        ref = prepare_clip(clip=wrap, original_img=original_img, img=latent)
    close(O.patchify(latent), ref["img"], 0.0, "img (patchified)")
    close(O.make_img_ids(B, 3, 5), ref["img_ids"], 0.0, "img_ids")
    _, opc, opt5 = O.clip_wrapper_forward(sd_t, sd_w, original_img, tc)
    close(opt5, ref["txt"], 2e-5, "txt")
    close(opc, ref["vec"], 2e-5, "vec")
    torch.save(dict(kind="prepare_clip", cfg=tc.__dict__, clip_dim=48, t5_dim=96, seed=seed, key_shapes_tower=ks_t,
                    key_shapes_wrap=ks_w, original_img=original_img, latent=latent,
                    out={k: v.detach().clone() for k, v in ref.items()}),
               os.path.join(GOLD, "prepare_clip_small.pt"))


def golden_ae(seed=21):
    print("[ae encoder]")
    ac = O.AECfg(ch=64)
    ae = AutoEncoder(AutoEncoderParams(resolution=256, in_channels=3, ch=ac.ch, out_ch=3, ch_mult=list(ac.ch_mult),
                                       num_res_blocks=2, z_channels=16, scale_factor=ac.scale_factor,
                                       shift_factor=ac.shift_factor))
    ks = O.ae_encoder_key_shapes(ac)
    sd = O.synth_state_dict(ks, seed)
    missing, unexpected = ae.encoder.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(2, 3, 48, 48, generator=g)
    x = (img - 0.5) / 0.5
    with torch.no_grad():
        mom = ae.encoder(x)
        torch.manual_seed(seed)
        z = ae.encode(x)  # consumes randn_like(mean)
        torch.manual_seed(seed)
        noise = torch.randn_like(mom[:, :16])
        omom = O.ae_encoder_forward(sd, x, ac)
        oz = O.ae_encode(sd, x, ac, noise)
    close(omom, mom, 2e-5, "moments")
    close(oz, z, 2e-5, "z (given the same noise)")
    torch.save(dict(kind="ae", cfg=ac.__dict__, seed=seed, key_shapes=ks, img=img, noise=noise, moments=mom, z=z),
               os.path.join(GOLD, "ae_small.pt"))


def golden_ae_decoder(seed=23):
    """The reference's Decoder (autoencoder.py:183-259) and AutoEncoder.decode on synthetic weights."""
    print("[ae decoder]")
    ac = O.AECfg(ch=64)
    ae = AutoEncoder(AutoEncoderParams(resolution=256, in_channels=3, ch=ac.ch, out_ch=3, ch_mult=list(ac.ch_mult),
                                       num_res_blocks=2, z_channels=16, scale_factor=ac.scale_factor,
                                       shift_factor=ac.shift_factor))
    ks = O.ae_decoder_key_shapes(ac)
    assert {k: tuple(v.shape) for k, v in ae.decoder.state_dict().items()} == {k: tuple(v) for k, v in ks.items()}
    sd = O.synth_state_dict(ks, seed)
    ae.decoder.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(2, 16, 6, 10, generator=g)
    with torch.no_grad():
        img = ae.decode(z)
        oimg = O.ae_decode(sd, z, ac)
    close(oimg, img, 2e-5, "decoded image")
    torch.save(dict(kind="ae_decoder", cfg=ac.__dict__, seed=seed, key_shapes=ks, z=z, image=img),
               os.path.join(GOLD, "ae_decoder_small.pt"))


def ref_flux(fc: O.FluxCfg):
    return Flux(FluxParams(in_channels=fc.in_channels, vec_in_dim=fc.vec_in_dim, context_in_dim=fc.context_in_dim,
                           hidden_size=fc.hidden_size, mlp_ratio=fc.mlp_ratio, num_heads=fc.num_heads, depth=fc.depth,
                           depth_single_blocks=fc.depth_single_blocks, axes_dim=list(fc.axes_dim), theta=fc.theta,
                           qkv_bias=fc.qkv_bias, guidance_embed=fc.guidance_embed))


def golden_flux(name: str, n_txt: int, seed=31, video_ids=False):
    print(f"[flux:{name}]")
    fc = O.FluxCfg(vec_in_dim=32, context_in_dim=48, hidden_size=256, num_heads=2, depth=1, depth_single_blocks=2)
    dit = ref_flux(fc)
    ks = O.flux_key_shapes(fc)
    assert {k: tuple(v.shape) for k, v in dit.state_dict().items()} == {k: tuple(v) for k, v in ks.items()}
    sd = O.synth_state_dict(ks, seed)
    dit.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(seed)
    B, h2, w2 = 2, 3, 5
    L = h2 * w2
    img = torch.randn(B, L, 64, generator=g)
    txt = torch.randn(B, n_txt, 48, generator=g)
    y = torch.randn(B, 32, generator=g)
    t = torch.sigmoid(torch.randn(B, generator=g))
    guidance = torch.full((B,), 4.0)
    img_ids = O.make_img_ids(B, h2, w2, 1.0 if video_ids else 0.0)
    if video_ids:  # two conditioning frames at t=0 and t=2 (R/train_OpenAICLIP_video_stage1.py:405-420)
        gh = int(round((n_txt // 2) ** 0.5))
        txt_ids = torch.cat([O.create_spatio_temporal_ids(B, 0, gh, gh), O.create_spatio_temporal_ids(B, 2, gh, gh)], 1)
    else:
        txt_ids = torch.zeros(B, n_txt, 3)
    target = torch.randn(B, L, 64, generator=g)
    img_r, txt_r, y_r = (v.clone().requires_grad_(True) for v in (img, txt, y))
    pred = dit(img=img_r, img_ids=img_ids, txt=txt_r, txt_ids=txt_ids, timesteps=t, y=y_r, guidance=guidance)
    loss = F.mse_loss(pred.float(), target)
    loss.backward()
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    img_o, txt_o, y_o = (v.clone().requires_grad_(True) for v in (img, txt, y))
    opred = O.flux_forward(sdo, fc, img_o, img_ids, txt_o, txt_ids, t, y_o, guidance)
    oloss = F.mse_loss(opred.float(), target)
    oloss.backward()
    close(opred, pred, 2e-5, "pred")
    close(oloss, loss, 1e-6, "loss")
    close(img_o.grad, img_r.grad, 1e-4, "d img")
    close(txt_o.grad, txt_r.grad, 1e-4, "d txt")
    close(y_o.grad, y_r.grad, 1e-4, "d y")
    grads = {}
    for k in ("img_in.weight", "double_blocks.0.img_attn.qkv.weight", "double_blocks.0.txt_mod.lin.bias",
              "double_blocks.0.img_attn.norm.query_norm.scale", "single_blocks.1.linear2.weight",
              "single_blocks.0.norm.key_norm.scale", "final_layer.adaLN_modulation.1.weight", "time_in.in_layer.weight"):
        gr = dict(dit.named_parameters())[k].grad
        close(sdo[k].grad, gr, 2e-4, f"d {k}")
        grads[k] = gr.clone()
    # bf16 run of the reference (what the training scripts actually execute: weight_dtype casts)
    dit_bf = ref_flux(fc)
    dit_bf.load_state_dict(sd, strict=True)
    dit_bf = dit_bf.to(torch.bfloat16)
    bf = torch.bfloat16
    txt_b, y_b = txt.to(bf).requires_grad_(True), y.to(bf).requires_grad_(True)
    pred_bf = dit_bf(img=img.to(bf), img_ids=img_ids.to(bf), txt=txt_b, txt_ids=txt_ids.to(bf),
                     timesteps=t.to(bf), y=y_b, guidance=guidance.to(bf))
    F.mse_loss(pred_bf.float(), target).backward()

    def cos(a, b):
        a, b = a.double().flatten(), b.double().flatten()
        return float(a @ b / (a.norm() * b.norm()))

    # how well the reference's OWN bf16 run (what the training scripts execute) tracks its fp32 gradients:
    # the calibration floor for the bf16 CUDA path (tests require: not worse than this by more than 0.01)
    pb = dict(dit_bf.named_parameters())
    bf16_cos = {k: cos(pb[k].grad, g) for k, g in grads.items()}
    bf16_cos["d_txt"] = cos(txt_b.grad, txt_r.grad)
    bf16_cos["d_y"] = cos(y_b.grad, y_r.grad)
    print("  reference bf16-vs-fp32 gradient cosines:", {k: round(v, 4) for k, v in bf16_cos.items()})
    torch.save(dict(kind="flux", cfg=fc.__dict__, seed=seed, key_shapes=ks, img=img, txt=txt, y=y, t=t, guidance=guidance,
                    img_ids=img_ids, txt_ids=txt_ids, target=target, pred=pred.detach(), loss=loss.detach(),
                    d_img=img_r.grad, d_txt=txt_r.grad, d_y=y_r.grad, grads=grads, pred_bf16=pred_bf.detach(),
                    ref_bf16_grad_cos=bf16_cos),
               os.path.join(GOLD, f"flux_{name}.pt"))


def golden_sampler(seed=71):
    """The reference's own sampler (src/flux/sampling.py) driving the reference's Flux: schedule, 4 Euler steps with the
    true-CFG branch switched on from step 1, unpack."""
    import src.flux.sampling as RS  # noqa: E402  (reference)
    print("[sampler]")
    fc = O.FluxCfg(vec_in_dim=32, context_in_dim=48, hidden_size=256, num_heads=2, depth=1, depth_single_blocks=2)
    dit = ref_flux(fc)
    ks = O.flux_key_shapes(fc)
    sd = O.synth_state_dict(ks, seed)
    dit.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(seed)
    B, height, width = 2, 64, 96
    noise = RS.get_noise(B, height, width, torch.device("cpu"), torch.float32, seed)
    h2, w2 = noise.shape[2] // 2, noise.shape[3] // 2
    img = rearrange(noise, "b c (h ph) (w pw) -> b (h w) (c ph pw)", ph=2, pw=2)
    n_txt = 3
    txt, neg_txt = torch.randn(B, n_txt, 48, generator=g), torch.randn(B, n_txt, 48, generator=g)
    vec, neg_vec = torch.randn(B, 32, generator=g), torch.randn(B, 32, generator=g)
    img_ids, txt_ids = O.make_img_ids(B, h2, w2), torch.zeros(B, n_txt, 3)
    sched = RS.get_schedule(4, img.shape[1], shift=True)
    sched_plain = RS.get_schedule(7, 1024, shift=False)
    sched_big = RS.get_schedule(25, 4096)
    with torch.no_grad():
        out = RS.denoise(dit, img, img_ids, txt, txt_ids, vec, neg_txt, txt_ids, neg_vec, sched, guidance=4.0, true_gs=2.5,
                         timestep_to_start_cfg=1)
        oout = O.denoise(sd, fc, img, img_ids, txt, txt_ids, vec, neg_txt, txt_ids, neg_vec, sched, 4.0, 2.5, 1)
    close(oout, out, 5e-5, "denoised latent")
    assert O.get_schedule(4, img.shape[1]) == sched and O.get_schedule(7, 1024, shift=False) == sched_plain
    assert O.get_schedule(25, 4096) == sched_big
    un = RS.unpack(out, height, width)
    assert torch.equal(O.unpack(out, height, width), un)
    torch.save(dict(kind="sampler", cfg=fc.__dict__, seed=seed, key_shapes=ks, height=height, width=width, noise=noise,
                    img=img, txt=txt, neg_txt=neg_txt, vec=vec, neg_vec=neg_vec, img_ids=img_ids, txt_ids=txt_ids,
                    schedule=sched, schedule_plain=sched_plain, schedule_big=sched_big, true_gs=2.5, start_cfg=1,
                    denoised=out, unpacked=un), os.path.join(GOLD, "sampler_small.pt"))


def golden_step_small(seed=41):
    """A whole image-mode stage-1 micro-step with every component at reduced size, through the reference's
    own prepare_clip / Flux / AutoEncoder / OpenAICLIP code, vs oracle.stage1_image_step."""
    print("[stage-1 image step, reduced sizes]")
    tc = O.TowerCfg("clip", 128, 2, 2, 512, 112, 14, 64, 1e-5, "quick_gelu")
    fc = O.FluxCfg(vec_in_dim=32, context_in_dim=48, hidden_size=256, num_heads=2, depth=1, depth_single_blocks=1)
    ac = O.AECfg(ch=64)
    wrap, sd_t, sd_w, ks_t, ks_w = build_reference_wrapper(tc, 32, 48, seed)
    dit = ref_flux(fc)
    ks_d = O.flux_key_shapes(fc)
    sd_d = O.synth_state_dict(ks_d, seed + 2)
    dit.load_state_dict(sd_d, strict=True)
    ae = AutoEncoder(AutoEncoderParams(resolution=256, in_channels=3, ch=ac.ch, out_ch=3, ch_mult=list(ac.ch_mult),
                                       num_res_blocks=2, z_channels=16, scale_factor=ac.scale_factor,
                                       shift_factor=ac.shift_factor))
    ks_a = O.ae_encoder_key_shapes(ac)
    sd_a = O.synth_state_dict(ks_a, seed + 3)
    ae.encoder.load_state_dict(sd_a, strict=True)
    for n_, p_ in wrap.named_parameters():
        p_.requires_grad_("project_clip" in n_ or "project_t5" in n_)
    ae.requires_grad_(False)
    g = torch.Generator().manual_seed(seed)
    B = 2
    img = torch.rand(B, 3, 112, 112, generator=g)
    mean = torch.tensor(OPENAI_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(OPENAI_STD).view(1, 3, 1, 1)
    # --- the reference step body (train_SigLIP_stage1.py:242-263), RNG draws in its order ---
    torch.manual_seed(seed)
    with torch.no_grad():
        x_1 = ae.encode(((img - 0.5) / 0.5).float())                     # draw #1
    inp = prepare_clip(clip=wrap, original_img=(img - mean) / std, img=x_1)
    x_1p = rearrange(x_1, "b c (h ph) (w pw) -> b (h w) (c ph pw)", ph=2, pw=2)
    t = torch.sigmoid(torch.randn((B,)) * 1.0)                            # draw #2
    x_0 = torch.randn_like(x_1p)                                          # draw #3
    x_t = (1 - t[:, None, None]) * x_1p + t[:, None, None] * x_0
    pred = dit(img=x_t, img_ids=inp["img_ids"], txt=inp["txt"], txt_ids=inp["txt_ids"], y=inp["vec"], timesteps=t,
               guidance=torch.full((B,), 4.0))
    loss = F.mse_loss(pred.float(), (x_0 - x_1p).float(), reduction="mean")
    loss.backward()
    # --- oracle with the same three draws ---
    torch.manual_seed(seed)
    noise = torch.randn(B, 16, 14, 14)
    t_o = torch.sigmoid(torch.randn((B,)) * 1.0)
    x0_o = torch.randn_like(x_1p)
    assert torch.equal(t_o, t) and torch.equal(x0_o, x_0)
    sdw = {k: v.clone().requires_grad_(True) for k, v in sd_w.items()}
    sdd = {k: v.clone().requires_grad_(True) for k, v in sd_d.items()}
    out = O.stage1_image_step(sd_t, sdw, sdd, sd_a, img, tc, fc, ac, OPENAI_MEAN, OPENAI_STD, noise, t, x_0)
    out.loss.backward()
    close(out.x_1, x_1p, 2e-5, "x_1 (patchified latent)")
    close(out.x_t, x_t, 2e-5, "x_t")
    close(out.pred, pred, 5e-5, "pred")
    close(out.loss, loss, 1e-5, "loss")
    close(sdw["project_t5.3.weight"].grad, wrap.project_t5[3].weight.grad, 5e-4, "d project_t5.3.weight")
    close(sdd["txt_in.weight"].grad, dit.txt_in.weight.grad, 5e-4, "d txt_in.weight")
    torch.save(dict(kind="step", tower_cfg=tc.__dict__, flux_cfg=fc.__dict__, ae_cfg=ac.__dict__, seed=seed,
                    key_shapes=dict(tower=ks_t, wrap=ks_w, dit=ks_d, ae=ks_a), clip_dim=32, t5_dim=48, img=img,
                    ae_noise=noise, t=t, x_0=x_0, x_1=x_1p.detach(), x_t=x_t.detach(), pred=pred.detach(),
                    loss=loss.detach(), vec=inp["vec"].detach(), txt=inp["txt"].detach(),
                    grad_project_t5_3_weight=wrap.project_t5[3].weight.grad.clone(),
                    grad_txt_in_weight=dit.txt_in.weight.grad.clone(),
                    grad_final_linear_weight=dit.final_layer.linear.weight.grad.clone()),
               os.path.join(GOLD, "step_small.pt"))


def extract_from_reference(script: str, names: list[str]) -> dict:
    """exec the named top-level classes / functions of a reference train script WITHOUT importing the script
    (importing needs accelerate / diffusers / peft / omegaconf, none of which are installed)."""
    import ast
    import random
    import torch.nn as nn
    src = open(os.path.join(REF, script)).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "nn": nn, "random": random, "F": F}
    for node in tree.body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), script, "exec"), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return ns


def golden_video_small(seed=51):
    """Video-mode stage-1 step at reduced size through the reference's OWN pieces: AutoEncoder, HF CLIPModel inside
    the reference OpenAICLIP wrapper, Flux, prepare_clip (for img_ids), and VisualPromptAdapter /
    create_spatio_temporal_ids / build_windows_with_mask lifted out of the reference train scripts by AST.
    Wiring = use2frames next-frame prediction (cond frames t=0,1 -> target t=2), the step body of
    train_OpenAICLIP_use2frames_nextpredic_stage1.py:355-452."""
    print("[video step (use2frames wiring), reduced sizes]")
    ns = extract_from_reference("train_OpenAICLIP_use2frames_nextpredic_stage1.py",
                                ["VisualPromptAdapter", "create_spatio_temporal_ids"])
    ns_w = extract_from_reference("train_OpenAICLIP_sliding_windows_nextpredic_stage1.py", ["build_windows_with_mask"])
    tc = O.TowerCfg("clip", 128, 2, 2, 512, 112, 14, 64, 1e-5, "quick_gelu")
    fc = O.FluxCfg(vec_in_dim=64, context_in_dim=96, hidden_size=256, num_heads=2, depth=1, depth_single_blocks=1)
    ac = O.AECfg(ch=64)
    wrap, sd_t, sd_w, ks_t, ks_w = build_reference_wrapper(tc, 32, 48, seed)
    adapter = ns["VisualPromptAdapter"](in_dim=128, out_dim=96)
    ks_ad = O.adapter_key_shapes(128, 96)
    sd_ad = O.synth_state_dict(ks_ad, seed + 4)
    adapter.load_state_dict(sd_ad, strict=True)
    dit = ref_flux(fc)
    ks_d = O.flux_key_shapes(fc)
    sd_d = O.synth_state_dict(ks_d, seed + 2)
    dit.load_state_dict(sd_d, strict=True)
    ae = AutoEncoder(AutoEncoderParams(resolution=256, in_channels=3, ch=ac.ch, out_ch=3, ch_mult=list(ac.ch_mult),
                                       num_res_blocks=2, z_channels=16, scale_factor=ac.scale_factor,
                                       shift_factor=ac.shift_factor))
    ks_a = O.ae_encoder_key_shapes(ac)
    sd_a = O.synth_state_dict(ks_a, seed + 3)
    ae.encoder.load_state_dict(sd_a, strict=True)
    wrap.requires_grad_(False)
    ae.requires_grad_(False)
    g = torch.Generator().manual_seed(seed)
    B, T, S = 2, 5, 112
    frames = torch.rand(B, T, 3, S, S, generator=g)
    frame_mask = torch.tensor([[1, 1, 1, 1, 1], [1, 1, 1, 1, 0]], dtype=torch.bool)
    # --- window builder (sliding script) ---
    import random
    random.seed(seed)
    c0, c1, c2, tgt, avg_nw, bs_eff = ns_w["build_windows_with_mask"](frames, frame_mask, 3, 1, 8)
    random.seed(seed)
    conds_o, tgt_o, counts = O.build_windows_with_mask(frames, frame_mask.long(), 3, 1, 8, rng=random)
    assert torch.equal(conds_o[0], c0) and torch.equal(conds_o[2], c2) and torch.equal(tgt_o, tgt)
    assert sum(counts) == bs_eff
    # --- the use2frames step body: cond = frames 0 and 1 of each clip, target = frame 2 ---
    cond0, cond1, target = frames[:, 0], frames[:, 1], frames[:, 2]
    mean = torch.tensor(OPENAI_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(OPENAI_STD).view(1, 3, 1, 1)
    torch.manual_seed(seed)
    with torch.no_grad():
        x_1 = ae.encode(((target - 0.5) / 0.5).float())
        out0 = wrap.model.vision_model((cond0 - mean) / std, output_hidden_states=True)
        out1 = wrap.model.vision_model((cond1 - mean) / std, output_hidden_states=True)
        p0, p1 = out0.last_hidden_state[:, 1:, :], out1.last_hidden_state[:, 1:, :]
        vec = (wrap.model.visual_projection(out0.pooler_output) + wrap.model.visual_projection(out1.pooler_output)) / 2
    txt = adapter(torch.cat([p0, p1], dim=1))
    side = int(p0.shape[1] ** 0.5)
    ids0 = ns["create_spatio_temporal_ids"](side, side, time_step=0, device="cpu")
    ids1 = ns["create_spatio_temporal_ids"](side, side, time_step=1, device="cpu")
    txt_ids = torch.cat([ids0, ids1], dim=0)[None].repeat(B, 1, 1).float()
    dummy = prepare_clip(wrap, (cond0 - mean) / std, x_1)
    img_ids = dummy["img_ids"].clone()  # on the GPU the reference's .to(device) materialises the expanded view
    img_ids[..., 0] = 2.0
    x_1p = rearrange(x_1, "b c (h ph) (w pw) -> b (h w) (c ph pw)", ph=2, pw=2)
    t = torch.sigmoid(torch.randn((B,)) * 1.0)
    x_0 = torch.randn_like(x_1p)
    x_t = (1 - t[:, None, None]) * x_1p + t[:, None, None] * x_0
    pred = dit(img=x_t, img_ids=img_ids, txt=txt, txt_ids=txt_ids, y=vec, timesteps=t, guidance=torch.full((B,), 4.0))
    loss = F.mse_loss(pred.float(), (x_0 - x_1p).float(), reduction="mean")
    loss.backward()
    # --- oracle ---
    torch.manual_seed(seed)
    noise = torch.randn(B, 16, S // 8, S // 8)
    sda = {k: v.clone().requires_grad_(True) for k, v in sd_ad.items()}
    sdd = {k: v.clone().requires_grad_(True) for k, v in sd_d.items()}
    out = O.stage1_video_step(sd_t, sda, sdd, sd_a, [cond0, cond1], target, tc, fc, ac, OPENAI_MEAN, OPENAI_STD,
                              (0, 1), 2, noise, t, x_0)
    out.loss.backward()
    close(out.extras["txt_ids"], txt_ids, 0, "txt_ids")
    close(out.extras["img_ids"], img_ids, 0, "img_ids")
    close(out.txt, txt, 5e-5, "txt (adapter output)")
    close(out.vec, vec, 2e-5, "vec")
    close(out.pred, pred, 5e-5, "pred")
    close(out.loss, loss, 1e-5, "loss")
    close(sda["proj.2.weight"].grad, adapter.proj[2].weight.grad, 5e-4, "d adapter.proj.2.weight")
    close(sdd["txt_in.weight"].grad, dit.txt_in.weight.grad, 5e-4, "d txt_in.weight")
    torch.save(dict(kind="video_step", tower_cfg=tc.__dict__, flux_cfg=fc.__dict__, ae_cfg=ac.__dict__, seed=seed,
                    key_shapes=dict(tower=ks_t, adapter=ks_ad, dit=ks_d, ae=ks_a), frames=frames, frame_mask=frame_mask,
                    windows=dict(cond0=c0, cond2=c2, target=tgt, avg_nw=avg_nw, bs_eff=bs_eff),
                    cond_times=(0, 1), target_time=2, ae_noise=noise, t=t, x_0=x_0, x_1=x_1p.detach(), txt=txt.detach(),
                    txt_ids=txt_ids, img_ids=img_ids, vec=vec.detach(), pred=pred.detach(), loss=loss.detach(),
                    grad_adapter_proj2_weight=adapter.proj[2].weight.grad.clone(),
                    grad_adapter_proj3_bias=adapter.proj[3].bias.grad.clone(),
                    grad_txt_in_weight=dit.txt_in.weight.grad.clone()),
               os.path.join(GOLD, "video_step_small.pt"))


def golden_cfg1_full(seed=0):
    """BASELINE config 1 at FULL size (ViT-L/14-224 + full DiT + full AE, B=2, fp32, CPU) through the reference."""
    print("[cfg-1 full-size step through the reference -- ~1 min]")
    tc, fc, ac = O.openai_vit_l14(224), O.FluxCfg(), O.AECfg()
    wrap, sd_t, sd_w, ks_t, ks_w = build_reference_wrapper(tc, 768, 4096, seed)
    dit = ref_flux(fc)
    sd_d = O.synth_state_dict(O.flux_key_shapes(fc), seed + 2)
    dit.load_state_dict(sd_d, strict=True)
    ae = AutoEncoder(AutoEncoderParams(resolution=256, in_channels=3, ch=128, out_ch=3, ch_mult=[1, 2, 4, 4],
                                       num_res_blocks=2, z_channels=16, scale_factor=0.3611, shift_factor=0.1159))
    sd_a = O.synth_state_dict(O.ae_encoder_key_shapes(ac), seed + 3)
    ae.encoder.load_state_dict(sd_a, strict=True)
    for n_, p_ in wrap.named_parameters():
        p_.requires_grad_("project_clip" in n_ or "project_t5" in n_)
    ae.requires_grad_(False)
    g = torch.Generator().manual_seed(seed)
    B = 2
    img = torch.rand(B, 3, 224, 224, generator=g)
    mean = torch.tensor(OPENAI_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(OPENAI_STD).view(1, 3, 1, 1)
    torch.manual_seed(seed)
    with torch.no_grad():
        x_1 = ae.encode(((img - 0.5) / 0.5).float())
    inp = prepare_clip(clip=wrap, original_img=(img - mean) / std, img=x_1)
    cls = wrap(( img - mean) / std)[0].detach()
    x_1p = rearrange(x_1, "b c (h ph) (w pw) -> b (h w) (c ph pw)", ph=2, pw=2)
    t = torch.sigmoid(torch.randn((B,)) * 1.0)
    x_0 = torch.randn_like(x_1p)
    x_t = (1 - t[:, None, None]) * x_1p + t[:, None, None] * x_0
    pred = dit(img=x_t, img_ids=inp["img_ids"], txt=inp["txt"], txt_ids=inp["txt_ids"], y=inp["vec"], timesteps=t,
               guidance=torch.full((B,), 4.0))
    loss = F.mse_loss(pred.float(), (x_0 - x_1p).float(), reduction="mean")
    loss.backward()
    torch.manual_seed(seed)
    noise = torch.randn(B, 16, 28, 28)
    print("  reference loss", loss.item())
    with torch.no_grad():
        out = O.stage1_image_step(sd_t, sd_w, sd_d, sd_a, img, tc, fc, ac, OPENAI_MEAN, OPENAI_STD, noise, t, x_0)
    close(out.loss, loss, 1e-4, "loss (oracle vs reference, full size)")
    close(out.class_token, cls, 1e-4, "class_token")
    torch.save(dict(kind="cfg1", seed=seed, img=img, ae_noise=noise, t=t, x_0=x_0, loss=loss.detach(),
                    class_token=cls, vec=inp["vec"].detach(), x_1=x_1p.detach(), pred=pred.detach(),
                    grad_norm_project_t5_1_weight=wrap.project_t5[1].weight.grad.norm(),
                    grad_final_linear_weight=dit.final_layer.linear.weight.grad.clone(),
                    grad_img_in_weight=dit.img_in.weight.grad.clone()),
               os.path.join(GOLD, "cfg1_full.pt"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="", help="comma-separated subset: tower,metaclip,prepare,lora,ae,decoder,flux,sampler,step,video")
    ap.add_argument("--full", action="store_true", help="also run BASELINE config 1 at full size (~1-2 min, ~12 GB)")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    only = set(filter(None, args.only.split(",")))
    want = lambda k: not only or k in only
    if want("tower"):
        golden_tower("clip_small", O.TowerCfg("clip", 128, 2, 2, 512, 56, 14, 64, 1e-5, "quick_gelu"))
        golden_tower("siglip_small", O.TowerCfg("siglip", 144, 2, 2, 304, 56, 14, 144, 1e-6, "gelu_tanh"))
    if want("metaclip"):   # MetaCLIP-H/14 geometry at reduced size: head_dim 80 (1280 / 16), through the reference's MetaCLIP class
        golden_tower("metaclip_h_small", O.TowerCfg("clip", 160, 2, 2, 320, 56, 14, 64, 1e-5, "quick_gelu"), metaclip="huge",
                     seed=13)
    if want("prepare"):
        golden_prepare_clip()
    if want("lora"):
        golden_tower_lora("clip_small", O.TowerCfg("clip", 128, 2, 2, 512, 56, 14, 64, 1e-5, "quick_gelu"), all_linear=True)
        golden_tower_lora("siglip_small", O.TowerCfg("siglip", 144, 2, 2, 304, 56, 14, 144, 1e-6, "gelu_tanh"), all_linear=False)
    if want("ae"):
        golden_ae()
    if want("decoder"):
        golden_ae_decoder()
    if want("flux"):
        golden_flux("img", n_txt=1)
        golden_flux("video", n_txt=8, video_ids=True)
    if want("sampler"):
        golden_sampler()
    if want("step"):
        golden_step_small()
    if want("video"):
        golden_video_small()
    if args.full:
        golden_cfg1_full()
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
