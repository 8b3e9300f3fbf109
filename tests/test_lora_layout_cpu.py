"""Host-side layout of one LoRA-wrapped linear group (no GPU): the fp32 gradient staging with the bias gradient riding as
a column of the dB block (DESIGN.md finding 29), the zero-padded A operand the u GEMM reads, and the scatter table."""
import torch

from genhancer_b200.clip_models import lora
from genhancer_b200.clip_models.tower_engine import LinGroup


def _group(train_bias=True, r=16, n_members=1):
    mods, pairs = [], []
    for _ in range(n_members):
        lin = torch.nn.Linear(64, 48)
        lin.weight.requires_grad_(False)
        lin.bias.requires_grad_(train_bias)
        mods.append(lin)
        pairs.append(lora.LoraPair(64, 48, r))
    return LinGroup([f"l{i}" for i in range(n_members)], mods, pairs, 2.0), mods, pairs


def test_bias_gradient_is_a_column_of_the_dB_staging():
    g, _, _ = _group(train_bias=True, r=16, n_members=3)      # a fused q/k/v group: R = 48
    assert (g.R, g.RX, g.N, g.K) == (48, 16, 144, 64)
    assert g.A_ext.shape == (64, 64) and g.A.data_ptr() == g.A_ext.data_ptr() and g.A.shape == (48, 64)
    assert float(g.A_ext[48:].abs().max()) == 0.0              # zero rows: those columns of u are the GEMM's "bias" alone
    assert g.u_bias.tolist() == [0.0] * 48 + [1.0] + [0.0] * 15
    n = g.staging_numel()
    assert n % 64 == 0 and n >= 48 * 64 + 144 * 64
    flat = torch.zeros(n)
    g.carve(flat)
    assert g.gA.shape == (48, 64) and g.gBx.shape == (144, 64)
    assert g.gB.shape == (144, 48) and g.gB.stride() == (64, 1) and g.gB.data_ptr() == g.gBx.data_ptr()
    assert g.gb.shape == (144,) and g.gb.stride() == (64,) and g.gb.data_ptr() == g.gBx[:, 48].data_ptr()
    # what the wgrad GEMM writes there is [dB | db | 0]: the views pick the right pieces
    g.gBx.copy_(torch.arange(144 * 64, dtype=torch.float32).view(144, 64))
    assert torch.equal(g.gb, g.gBx[:, 48]) and torch.equal(g.gB, g.gBx[:, :48])


def test_without_trainable_bias_or_lora_the_layout_is_the_plain_one():
    g, _, _ = _group(train_bias=False)
    assert g.RX == 0 and g.u_bias is None and g.A_ext.shape == (16, 64)
    flat = torch.zeros(g.staging_numel())
    g.carve(flat)
    assert g.gBx.shape == (48, 16) and g.gb is None
    lin = torch.nn.Linear(64, 48)
    lin.weight.requires_grad_(False)
    g2 = LinGroup(["l"], [lin], [None], 1.0)                   # no LoRA pair, trainable bias: column-sum path
    assert g2.R == 0 and g2.RX == 0 and g2.A is None
    flat = torch.zeros(g2.staging_numel())
    g2.carve(flat)
    assert g2.gb.shape == (48,) and g2.gb.is_contiguous()
