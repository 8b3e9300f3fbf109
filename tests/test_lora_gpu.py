"""Stage-2 tower on the B200: LoRA folded into the base GEMMs (second operand pair), the tower's gradient chain and
the LoRA / bias wgrads, against golden values from torch autograd over the HF modules inside the reference's wrappers
with peft's LoRA layer restated as forward hooks (oracle/make_golden.py::golden_tower_lora; peft itself is not in
the image)."""
import copy

import pytest
import torch

from conftest import cosine, load_golden

pytestmark = pytest.mark.gpu

OPENAI_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_STD = (0.26862954, 0.26130258, 0.27577711)


def _build(fx, bias="lora_only"):
    from genhancer_b200.clip_models import lora
    from oracle import genhancer_oracle as O
    from test_tower_ae_gpu import _build_wrapper
    wrap = _build_wrapper(fx)
    tgt = "all-linear" if fx["all_linear"] else lora.SIGLIP_TARGETS
    wrap.model = lora.get_peft_model(wrap.model, lora.LoraConfig(r=fx["r"], lora_alpha=fx["alpha"], target_modules=tgt,
                                                                 lora_dropout=0.0, bias=bias))
    flat = O.synth_state_dict(fx["key_shapes_lora"], fx["seed"] + 2)
    with torch.no_grad():
        for key, pair in wrap.model.lora.items():
            n = key.replace("/", ".")
            pair.A.copy_(flat[f"{n}.lora_A"])
            pair.B.copy_(flat[f"{n}.lora_B"] * fx["b_scale"])
    assert {k.replace("/", ".") for k in wrap.model.lora.keys()} == {k.rsplit(".", 1)[0] for k in fx["key_shapes_lora"]}
    return wrap


def _norm_input(fx):
    img = fx["img"].to("cuda")
    clip = fx["cfg"]["kind"] == "clip"
    mean = torch.tensor(OPENAI_MEAN if clip else (0.5,) * 3, device="cuda").view(1, 3, 1, 1)
    std = torch.tensor(OPENAI_STD if clip else (0.5,) * 3, device="cuda").view(1, 3, 1, 1)
    return (img - mean) / std


def _run(wrap, x):
    cls, pc, pt5 = wrap(x)
    out = wrap.model.vision_model(x, output_hidden_states=True)
    loss = pc.float().square().mean() + pt5.float().square().mean() + 0.1 * out.last_hidden_state[:, 1:].float().square().mean()
    return cls, pc, pt5, out, loss


@pytest.mark.parametrize("name", ["clip_small", "siglip_small"])
def test_lora_tower_forward_and_gradients_match_reference(name):
    fx = load_golden(f"tower_lora_{name}.pt")
    wrap = _build(fx)
    x = _norm_input(fx)
    cls, pc, pt5, out, loss = _run(wrap, x)
    assert cosine(out.last_hidden_state, fx["last_hidden_state"]) >= 0.999
    assert cosine(out.pooler_output, fx["pooler_output"]) >= 0.999
    assert cosine(cls, fx["class_token"]) >= 0.999
    assert cosine(pc, fx["projection_clip"]) >= 0.999 and cosine(pt5, fx["projection_t5"]) >= 0.999
    assert abs(loss.item() - fx["loss"].item()) / fx["loss"].item() < 1e-2      # north_star: loss within 1e-2 in bf16
    loss.backward()
    P = dict(wrap.model.named_parameters())
    checked = 0
    for k, gref in fx["grads"].items():
        if k.startswith("project_"):
            seq, idx, leaf = k.split(".")
            g = getattr(getattr(wrap, seq)[int(idx)], leaf).grad
        elif k.endswith(("lora_A", "lora_B")):
            pair = wrap.model.lora[k.rsplit(".", 1)[0].replace(".", "/")]
            g = pair.A.grad if k.endswith("lora_A") else pair.B.grad
        else:
            g = P[k].grad
        assert g is not None, k
        if "k_proj.bias" in k:  # mathematically zero (softmax is shift-invariant): the golden value is rounding noise
            assert g.abs().max().item() < 1e-3 * fx["grads"][k.replace("k_proj", "v_proj")].abs().max().item() + 1e-6
            continue
        assert cosine(g, gref) >= 0.99, (k, cosine(g, gref))
        assert abs(g.float().norm().item() / gref.norm().item() - 1) < 0.05, k
        checked += 1
    assert checked >= 30


def test_lora_gradient_accumulates_and_zero_b_is_identity():
    """Size-independent properties: two backward passes double every LoRA gradient; with B = 0 (peft's init) the
    wrapped tower reproduces the frozen tower exactly and dA = 0."""
    fx = load_golden("tower_lora_siglip_small.pt")
    wrap = _build(fx)
    x = _norm_input(fx)
    _run(wrap, x)[-1].backward()
    g1 = {k: p.A.grad.clone() for k, p in wrap.model.lora.items()}
    _run(wrap, x)[-1].backward()
    for k, p in wrap.model.lora.items():
        assert torch.allclose(p.A.grad, 2 * g1[k], rtol=1e-3, atol=1e-7), k
    with torch.no_grad():
        base = wrap.model.vision_model(x).pooler_output.clone()
        for p in wrap.model.lora.values():
            p.B.zero_()
        zero_b = wrap.model.vision_model(x).pooler_output
        from test_tower_ae_gpu import _build_wrapper
        frozen = _build_wrapper(fx).model.vision_model(x).pooler_output
    assert not torch.equal(base, zero_b)
    assert torch.equal(zero_b, frozen)  # B = 0 adds exact zeros in the fp32 accumulator
    for p in wrap.model.lora.values():
        p.A.grad = None
        p.B.grad = None
    _run(wrap, x)[-1].backward()
    for k, p in wrap.model.lora.items():
        assert p.A.grad.abs().max().item() == 0.0, k
        assert p.B.grad.abs().max().item() > 0.0, k


def test_merge_and_unload_matches_the_wrapped_tower():
    """merge_and_unload (train_SigLIP_stage2_all.py:305-311): W += (alpha/r) B A; the merged plain tower gives the
    wrapped tower's features, and save_pretrained writes HF key names."""
    from genhancer_b200.clip_models import lora, vision_tower as vt
    fx = load_golden("tower_lora_clip_small.pt")
    wrap = _build(fx)
    x = _norm_input(fx)
    with torch.no_grad():
        ref = wrap.model.vision_model(x).last_hidden_state
        sd = lora.merged_state_dict(wrap.model)
        assert not any("lora" in k for k in sd)
        plain = vt.VisionLanguageModel(wrap.model.config).to("cuda")
        missing, unexpected = plain.load_state_dict(sd, strict=False)
        assert not unexpected and all(k.startswith("text_projection") for k in missing)
        got = plain.vision_model(x).last_hidden_state
    assert cosine(got, ref) >= 0.9999
    assert copy.deepcopy(wrap.model) is not None  # the engine cache does not break deepcopy (reference: deepcopy(...).merge_and_unload())


def test_lora_dropout_mask_is_regenerated_and_group_matches_torch():
    """lora_dropout (0.1 in every stage-2 YAML): the mask is a pure function of (seed, offset, index); the folded
    forward / backward of one wrapped linear agrees with torch autograd given that mask."""
    from genhancer_b200 import kernels as K
    from genhancer_b200.clip_models import lora
    from genhancer_b200.clip_models.tower_engine import LinGroup
    torch.manual_seed(0)
    p, seed, off = 0.1, 1234, 7
    ones = torch.ones(64, 256, device="cuda", dtype=torch.bfloat16)
    m1 = K.dropout_fwd(ones, p, seed, off)
    assert torch.equal(m1, K.dropout_fwd(ones, p, seed, off))            # reproducible
    assert not torch.equal(m1, K.dropout_fwd(ones, p, seed, off + 1))    # new offset, new mask
    keep = (m1 != 0).float().mean().item()
    assert abs(keep - 0.9) < 0.02
    assert torch.allclose(m1[m1 != 0].float(), torch.tensor(1 / 0.9, device="cuda"), rtol=4e-3)
    acc = torch.zeros_like(ones)
    K.dropout_bwd_add(ones, acc, p, seed, off)
    assert torch.equal(acc, m1)                                            # backward regenerates the same mask
    # the device-resident part of the offset (what a CUDA-graph replay advances): offset + *base
    base = torch.tensor([5], dtype=torch.int64, device="cuda")
    assert torch.equal(K.dropout_fwd(ones, p, seed, off - 5, base), m1)
    base += 1 << 20
    m2 = K.dropout_fwd(ones, p, seed, off - 5, base)
    assert not torch.equal(m2, m1)
    acc.zero_()
    K.dropout_bwd_add(ones, acc, p, seed, off - 5, base)
    assert torch.equal(acc, m2)

    lin = torch.nn.Linear(256, 384).cuda()
    lin.bias.requires_grad_(True)
    lin.weight.requires_grad_(False)
    pair = lora.LoraPair(256, 384, 16).cuda()
    with torch.no_grad():
        pair.B.normal_(0, 0.05)
    g = LinGroup(["l"], [lin], [pair], 1.0)
    flat = torch.zeros(g.staging_numel(), device="cuda")
    g.carve(flat)
    tab = K.CopyTable("cuda")
    g.fill_frozen(tab)
    g.fill_trainable(tab)
    tab.run()
    x = torch.randn(64, 256, device="cuda").to(torch.bfloat16)
    dy = torch.randn(64, 384, device="cuda").to(torch.bfloat16)
    y, saved = g.fwd(x, drop=(p, seed, off))
    dx = g.bwd(dy, x, saved)
    mask = (m1.float() * 0.9).round()
    xr = x.float().requires_grad_(True)
    A, Bm = pair.A.detach().clone().requires_grad_(True), pair.B.detach().clone().requires_grad_(True)
    bias = lin.bias.detach().clone().requires_grad_(True)
    yr = xr @ lin.weight.float().t() + bias + ((xr * mask / 0.9) @ A.t()) @ Bm.t()
    yr.backward(dy.float())
    assert cosine(y, yr) >= 0.9999
    assert cosine(dx, xr.grad) >= 0.999
    assert cosine(g.gA, A.grad) >= 0.999 and cosine(g.gB, Bm.grad) >= 0.999 and cosine(g.gb, bias.grad) >= 0.9999


def test_lora_dropout_trains_and_eval_mode_is_deterministic():
    fx = load_golden("tower_lora_clip_small.pt")
    wrap = _build(fx)
    wrap.model.lora_config.lora_dropout = 0.1
    from genhancer_b200.clip_models import tower_engine
    wrap.model._engine = None
    x = _norm_input(fx)
    wrap.train()
    a = wrap.model.vision_model(x).pooler_output.detach().clone()
    b = wrap.model.vision_model(x).pooler_output.detach().clone()
    assert not torch.equal(a, b)                       # fresh masks per call in training mode
    _run(wrap, x)[-1].backward()
    assert all(p.A.grad is not None and torch.isfinite(p.A.grad).all() for p in wrap.model.lora.values())
    wrap.eval()
    c = wrap.model.vision_model(x).pooler_output
    assert torch.equal(c, wrap.model.vision_model(x).pooler_output)
    assert cosine(c, fx["pooler_output"]) >= 0.999     # eval: dropout off -> the golden features
