"""Hand-scheduled forward / backward of the CLIP / SigLIP vision tower on the sm_100a kernels.

Stage 1 runs the tower frozen (forward only, nothing kept).  Stage 2 (``train_*_stage2_{all,only}.py``) trains LoRA
pairs (peft r=16, /root/reference/Continuous/train_SigLIP_stage2_all.py:134-142) and, with ``bias="lora_only"``, the
biases of the wrapped linears: the tower then needs its activation gradient chain (dgrad through frozen weights)
plus the LoRA / bias wgrads.  This module is that schedule -- the tower-side counterpart of ``flux/engine.py`` --
replacing HF ``CLIPVisionTransformer.forward`` / ``SiglipVisionTransformer.forward`` (transformers
modeling_clip.py:667-696, modeling_siglip.py:586-654), peft's ``lora.Linear.forward`` and torch autograd over them.

LoRA is folded into the base GEMM instead of running two extra GEMMs per wrapped linear:

    forward   u = s x A^T  [M, r]                 (skinny GEMM, r = 16 per wrapped member of a fused group)
              y = x W^T + u B^T + b               (ONE tcgen05 GEMM with a second operand pair: one extra MMA k-step)
    backward  du = s dy B   [M, r]
              dx = dy W + du A                    (ONE GEMM, second operand pair again)
              dA = du^T x,  dB = dy^T u,  db = colsum(dy)    (fp32 staging -> .grad by one batched scatter launch)

q/k/v share one fused GEMM: their three A matrices are stacked ([48, D]) and their B matrices sit block-diagonally
in a [3 D', 48] operand.  Heads of 72 / 80 features (SigLIP-so400m, ViT-H) live in 128-wide slots so that the
D=128 attention kernels serve them (zero lanes contribute nothing); the pack / scatter tables do the
slot mapping.  Everything saved for backward stays resident in HBM (nothing is recomputed).
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

import os

from .. import kernels as K
from ..kernels import ACT_GELU_TANH, ACT_NONE, ACT_QUICK_GELU, BF16, F32

# patch-embed conv as ONE implicit-GEMM kernel (gh_patch_embed_fwd: the operand tile is gathered from the image into
# the MMA's shared-memory layout); GH_PATCH_IMPLICIT=0 selects the older gather-to-HBM + plain GEMM pair
IMPLICIT_PATCH_EMBED = os.environ.get("GH_PATCH_IMPLICIT", "1") != "0"
# A/B switches of the round-2 LoRA work (profiles/r02_same_box_ab.txt): the fused dropout kernels of csrc/lora_fused.cu, and
# the bias gradient as a ones column of u
LORA_FUSED = os.environ.get("GH_LORA_FUSED", "1") != "0"
LORA_ONES_COLUMN = os.environ.get("GH_LORA_ONES", "1") != "0"


def _alias_f32(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    return t if (t.dtype == F32 and t.is_contiguous()) else t.float().contiguous()


class LinGroup:
    """One or more frozen ``nn.Linear`` sharing an input, fused along the output features, with optional LoRA pairs.

    in_slots / out_slots = (d, dp, H): the input / every member's output is H heads of d features stored in dp-wide
    slots (None: dense)."""

    def __init__(self, names, mods, pairs, scaling, in_slots=None, out_slots=None):
        self.names, self.mods, self.pairs, self.s = names, mods, pairs, float(scaling)
        self.in_slots, self.out_slots = in_slots, out_slots
        w0 = mods[0].weight
        dev = w0.device
        self.K_log = w0.shape[1]
        self.K = in_slots[1] * in_slots[2] if in_slots else self.K_log
        self.N_log = [m.weight.shape[0] for m in mods]
        self.N_each = [out_slots[1] * out_slots[2] if out_slots else n for n in self.N_log]
        self.N = sum(self.N_each)
        self.row0 = [sum(self.N_each[:i]) for i in range(len(mods))]
        self.lora_idx = [i for i, p in enumerate(pairs) if p is not None]
        self.R = sum(pairs[i].A.shape[0] for i in self.lora_idx)
        self.has_bias = any(m.bias is not None for m in mods)
        self.train_bias = [m.bias is not None and m.bias.requires_grad for m in mods]
        self.w = torch.zeros(self.N, self.K, dtype=BF16, device=dev)
        self.bias = torch.zeros(self.N, dtype=F32, device=dev) if self.has_bias else None
        # Trainable biases of a LoRA-wrapped group: db = colsum(dy) = dy^T 1 rides in the dB = dy^T u GEMM as one more column.
        # u gets RX extra columns, the first of them a constant one (the u GEMM's "bias" on a zero row of A), so the wgrad
        # GEMM's output is [dB | db | 0]: no separate pass over dy (it was 145 column-sum launches, 7.4 ms of the SigLIP
        # stage-2 step).  An MMA with N = 32 costs what N = 16 costs.
        self.RX = 16 if (self.R and any(self.train_bias) and LORA_ONES_COLUMN) else 0
        self.A_ext = torch.zeros(self.R + self.RX, self.K, dtype=BF16, device=dev) if self.R else None
        self.A = self.A_ext[:self.R] if self.R else None
        self.u_bias = None
        if self.RX:
            self.u_bias = torch.zeros(self.R + self.RX, dtype=F32, device=dev)
            self.u_bias[self.R] = 1.0
        self.Bm = torch.zeros(self.N, self.R, dtype=BF16, device=dev) if self.R else None
        self.gA = self.gB = self.gb = self.gBx = None  # fp32 gradient staging, carved out of the engine's flat buffer

    # ---- slot mapping helpers -----------------------------------------------------------------------------
    def _rg(self):
        return (self.out_slots[0], self.out_slots[1]) if self.out_slots else None

    def _cg(self):
        return (self.in_slots[0], self.in_slots[1]) if self.in_slots else None

    def fill_frozen(self, tab: K.CopyTable) -> None:
        """bf16 copies of the frozen weights and fp32 copies of the frozen biases (slot layout)."""
        for i, m in enumerate(self.mods):
            r0, n = self.row0[i], self.N_each[i]
            tab.add(m.weight.detach(), self.w[r0:r0 + n], self.N_log[i], self.K_log, dst_rows=self._rg(), dst_cols=self._cg())
            if m.bias is not None and not self.train_bias[i]:
                tab.add(m.bias.detach().unsqueeze(1), self.bias[r0:r0 + n].unsqueeze(1), self.N_log[i], 1, dst_rows=self._rg())

    def fill_trainable(self, tab: K.CopyTable) -> None:
        """Per-step refresh: LoRA A / B (fp32 parameters) -> bf16 operands; trainable biases -> the fused bias."""
        col = 0
        for i in self.lora_idx:
            pr = self.pairs[i]
            r = pr.A.shape[0]
            tab.add(pr.A.detach(), self.A[col:col + r], r, self.K_log, dst_cols=self._cg())
            r0, n = self.row0[i], self.N_each[i]
            tab.add(pr.B.detach(), self.Bm[r0:r0 + n, col:col + r], self.N_log[i], r, dst_rows=self._rg())
            col += r
        for i, m in enumerate(self.mods):
            if self.train_bias[i]:
                r0, n = self.row0[i], self.N_each[i]
                tab.add(m.bias.detach().unsqueeze(1), self.bias[r0:r0 + n].unsqueeze(1), self.N_log[i], 1, dst_rows=self._rg())

    def staging_numel(self) -> int:
        n = 0
        if self.R:
            n += self.R * self.K + self.N * (self.R + self.RX)
        if any(self.train_bias) and not self.RX:
            n += self.N
        return (n + 63) // 64 * 64

    def carve(self, flat: torch.Tensor) -> None:
        o = 0
        if self.R:
            self.gA = flat[o:o + self.R * self.K].view(self.R, self.K)
            o += self.R * self.K
            self.gBx = flat[o:o + self.N * (self.R + self.RX)].view(self.N, self.R + self.RX)   # [dB | db | 0]
            self.gB = self.gBx[:, :self.R]
            o += self.N * (self.R + self.RX)
        if self.RX:
            self.gb = self.gBx[:, self.R]
        elif any(self.train_bias):
            self.gb = flat[o:o + self.N]

    def fill_scatter(self, tab: K.CopyTable) -> None:
        """fp32 staging -> ``.grad`` of the LoRA / bias parameters (accumulating, slot layout undone)."""
        col = 0
        for i in self.lora_idx:
            pr = self.pairs[i]
            r = pr.A.shape[0]
            r0, n = self.row0[i], self.N_each[i]
            tab.add(self.gA[col:col + r], pr.A.grad, r, self.K_log, accumulate=True, src_cols=self._cg())
            tab.add(self.gB[r0:r0 + n, col:col + r], pr.B.grad, self.N_log[i], r, accumulate=True, src_rows=self._rg())
            col += r
        for i, m in enumerate(self.mods):
            if self.train_bias[i]:
                r0, n = self.row0[i], self.N_each[i]
                tab.add(self.gb[r0:r0 + n].unsqueeze(1), m.bias.grad.unsqueeze(1), self.N_log[i], 1, accumulate=True,
                        src_rows=self._rg())

    def trainable(self):
        out = []
        for i in self.lora_idx:
            out += [self.pairs[i].A, self.pairs[i].B]
        out += [m.bias for i, m in enumerate(self.mods) if self.train_bias[i]]
        return out

    # ---- compute ------------------------------------------------------------------------------------------
    def fwd(self, x2d, drop=None, **epi):
        """-> (y, saved).  ``drop`` = (p, seed, offset) when lora_dropout is active: the LoRA branch then reads the
        dropped-out copy of x (peft: lora_B(lora_A(dropout(x)))); q/k/v of a fused group share one mask.
        ``saved`` = (u, xd, drop) is what ``bwd`` needs besides x."""
        if not self.R:
            return K.gemm(x2d, self.w, bias=self.bias, **epi), None
        xd = None
        if drop is not None and LORA_FUSED and self.R in (16, 32, 48) and self.K % 8 == 0:
            # the mask sits between x and A: dropout and the skinny product in ONE pass over x (csrc/lora_fused.cu)
            xd, ux = K.lora_dropout_fwd(x2d.contiguous(), self.A_ext, self.R, self.s, *drop)
            return K.gemm(x2d, self.w, bias=self.bias, a2=ux[:, :self.R], b2=self.Bm, **epi), (ux, xd, drop)
        if drop is not None:
            xd = K.dropout_fwd(x2d.contiguous(), *drop)
        ux = K.gemm(x2d if xd is None else xd, self.A_ext, alpha=self.s, bias=self.u_bias)   # [s x A^T | 1 | 0]
        return K.gemm(x2d, self.w, bias=self.bias, a2=ux[:, :self.R], b2=self.Bm, **epi), (ux, xd, drop)

    def bwd(self, dy2d, x2d, saved, need_dx=True, **dx_epi):
        """dx (or None).  Parameter gradients go to the fp32 staging (the engine scatters them once per backward)."""
        dx = None
        if self.R:
            u, xd, drop = saved
            du = K.gemm(dy2d, self.Bm, b_mn=True, alpha=self.s)
            if need_dx:
                if drop is None:
                    dx = K.gemm(dy2d, self.w, b_mn=True, a2=du, b2=self.A, **dx_epi)
                else:  # the mask sits between x and A: dx = dy W + mask * (du A) / (1 - p), then the epilogue math
                    dx = K.gemm(dy2d, self.w, b_mn=True)
                    fused = LORA_FUSED and self.R in (16, 32, 48) and self.K % 8 == 0
                    ag = bool(dx_epi.get("act_grad"))
                    if dx_epi and not ag:
                        raise NotImplementedError(f"LoRA dropout with dgrad epilogue {sorted(dx_epi)}")
                    if fused:      # du A never exists in memory; act'(pre) applied in the same read-modify-write
                        pre = dx_epi["aux_in"] if ag else None
                        if pre is not None and not (pre.is_contiguous() and pre.shape == dx.shape):
                            pre = None
                        K.lora_dropout_bwd(du, self.A, dx, *drop, act_pre=pre, act=dx_epi["act"] if pre is not None else ACT_NONE)
                        if ag and pre is None:
                            dx = K.act_bwd(dx, dx_epi["aux_in"], dx_epi["act"])
                    else:
                        K.dropout_bwd_add(K.gemm(du, self.A, b_mn=True), dx, *drop)
                        if ag:
                            dx = K.act_bwd(dx, dx_epi["aux_in"], dx_epi["act"])
            # skinny outputs reduced over every token of the batch: split-K, partial products added into the staging
            # (which is all-zero between backward calls, see TowerEngine._scatter_grads)
            K.gemm(du, x2d if xd is None else xd, a_mn=True, b_mn=True, out=self.gA, k_splits=-1)
            K.gemm(dy2d, u, a_mn=True, b_mn=True, out=self.gBx, k_splits=-1)   # [dB | db | 0]
        elif need_dx:
            dx = K.gemm(dy2d, self.w, b_mn=True, **dx_epi)
        if self.gb is not None and not self.RX:
            K.colsum(dy2d, self.gb)
        return dx


class TowerEngine:
    """Prepared operands + schedules for one ``VisionLanguageModel`` (vision side)."""

    def __init__(self, model):
        self.model = model
        vm = model.vision_model
        self.vm = vm
        c = vm.config
        self.c = c
        self.D, self.H, self.T = c.hidden_size, c.num_attention_heads, c.num_tokens
        self.d, self.dp = self.D // self.H, vm.head_dim_padded
        self.act = ACT_QUICK_GELU if c.hidden_act == "quick_gelu" else ACT_GELU_TANH
        self.lora_cfg = getattr(model, "lora_config", None)
        self.s = self.lora_cfg.scaling if self.lora_cfg is not None else 1.0
        self.p_drop = float(getattr(self.lora_cfg, "lora_dropout", 0.0) or 0.0)
        self._drop_seed = None
        self._drop_calls = 0
        self._drop_base = None      # int64 device scalar: advances by _DROP_STRIDE per training forward (graph-replayable)
        self._frozen_key = None
        self._train_key = None
        self._grad_key = None
        self._build()

    def __deepcopy__(self, memo):  # prepared operands are per-object caches: a copied model rebuilds its own
        return None

    # ---- operand preparation ------------------------------------------------------------------------------
    def _pair(self, dotted: str):
        lo = getattr(self.model, "lora", None)
        if lo is None:
            return None
        key = dotted.replace(".", "/")
        return lo[key] if key in lo else None

    def _group(self, prefix, leafs, in_slots=None, out_slots=None):
        mods, names = [], []
        for leaf in leafs:
            name = f"{prefix}.{leaf}" if prefix else leaf
            m = self.model
            for part in name.split("."):
                m = getattr(m, part)
            mods.append(m)
            names.append(name)
        return LinGroup(names, mods, [self._pair(n) for n in names], self.s, in_slots, out_slots)

    def _build(self):
        c, vm = self.c, self.vm
        D, H, d, dp = self.D, self.H, self.d, self.dp
        slots = (d, dp, H) if d != dp else None
        dev = vm.post_layernorm.weight.device
        self.dev = dev
        self.layers = []
        for i in range(c.num_hidden_layers):
            p = f"vision_model.encoder.layers.{i}"
            self.layers.append(SimpleNamespace(
                qkv=self._group(f"{p}.self_attn", ["q_proj", "k_proj", "v_proj"], out_slots=slots),
                o=self._group(f"{p}.self_attn", ["out_proj"], in_slots=slots),
                fc1=self._group(f"{p}.mlp", ["fc1"]), fc2=self._group(f"{p}.mlp", ["fc2"]),
                mod=vm.encoder.layers[i]))
        self.groups = [g for L in self.layers for g in (L.qkv, L.o, L.fc1, L.fc2)]
        if c.kind == "siglip":
            self.h_fc1 = self._group("vision_model.head.mlp", ["fc1"])
            self.h_fc2 = self._group("vision_model.head.mlp", ["fc2"])
            self.groups += [self.h_fc1, self.h_fc2]
        else:
            self.proj = self._group("", ["visual_projection"]) if hasattr(self.model, "visual_projection") else None
            if self.proj is not None:
                self.groups.append(self.proj)
        n_stage = sum(g.staging_numel() for g in self.groups)
        self.staging = torch.zeros(max(n_stage, 1), dtype=F32, device=dev)
        o = 0
        for g in self.groups:
            g.carve(self.staging[o:o + g.staging_numel()])
            o += g.staging_numel()
        self.has_trainable = any(g.trainable() for g in self.groups)
        self._head_gbo = torch.zeros(D, dtype=F32, device=dev) if c.kind == "siglip" else None
        self.sig = _signature(self.model)
        # embeddings / norms / MAP-head attention: small frozen tensors kept as plain copies
        kdim = 3 * c.patch_size * c.patch_size
        self.patch_ld = (kdim + 7) // 8 * 8
        self.kdim = kdim

    def _refresh_frozen(self):
        vm, c = self.vm, self.c
        frozen = [p for p in self.model.parameters() if not p.requires_grad]
        key = tuple((p.data_ptr(), p._version) for p in frozen)
        if key == self._frozen_key:
            return
        self._frozen_key = key
        D, H, d, dp = self.D, self.H, self.d, self.dp
        tab = K.CopyTable(self.dev)
        for g in self.groups:
            g.fill_frozen(tab)
        W = {}
        pw = torch.zeros(D, self.patch_ld, dtype=BF16, device=self.dev)
        pw[:, :self.kdim] = vm.embeddings.patch_embedding.weight.detach().reshape(D, self.kdim).to(BF16)
        W["patch_w"] = pw[:, :self.kdim]
        pb = vm.embeddings.patch_embedding.bias
        W["patch_b"] = _alias_f32(pb) if pb is not None else None
        W["cls"] = _alias_f32(vm.embeddings.class_embedding) if c.kind == "clip" else None
        W["pos"] = _alias_f32(vm.embeddings.position_embedding.weight)
        ln = lambda m: (_alias_f32(m.weight), _alias_f32(m.bias))
        if c.kind == "clip":
            W["pre_ln"] = ln(vm.pre_layrnorm)
        W["post_ln"] = ln(vm.post_layernorm)
        W["ln1"] = [ln(L.mod.layer_norm1) for L in self.layers]
        W["ln2"] = [ln(L.mod.layer_norm2) for L in self.layers]
        if c.kind == "siglip":  # MAP head: nn.MultiheadAttention packs q,k,v in in_proj_{weight,bias}
            hd = vm.head
            Wi, bi = hd.attention.in_proj_weight.detach(), hd.attention.in_proj_bias.detach()
            rg, cg = ((d, dp), (d, dp)) if d != dp else (None, None)
            Hd = dict(probe=hd.probe.detach().reshape(1, D).to(BF16).contiguous(),
                      wq=torch.zeros(H * dp, D, dtype=BF16, device=self.dev), bq=torch.zeros(H * dp, dtype=F32, device=self.dev),
                      wkv=torch.zeros(2 * H * dp, D, dtype=BF16, device=self.dev),
                      bkv=torch.zeros(2 * H * dp, dtype=F32, device=self.dev),
                      wo=torch.zeros(D, H * dp, dtype=BF16, device=self.dev), ln=ln(hd.layernorm))
            tab.add(Wi[:D], Hd["wq"], D, D, dst_rows=rg)
            tab.add(bi[:D].unsqueeze(1), Hd["bq"].unsqueeze(1), D, 1, dst_rows=rg)
            for j in range(2):
                tab.add(Wi[(1 + j) * D:(2 + j) * D], Hd["wkv"][j * H * dp:(j + 1) * H * dp], D, D, dst_rows=rg)
                tab.add(bi[(1 + j) * D:(2 + j) * D].unsqueeze(1), Hd["bkv"][j * H * dp:(j + 1) * H * dp].unsqueeze(1), D, 1,
                        dst_rows=rg)
            tab.add(hd.attention.out_proj.weight.detach(), Hd["wo"], D, D, dst_cols=cg)
            ob = hd.attention.out_proj.bias
            Hd["bo"] = _alias_f32(ob)  # fp32 parameter storage itself: always current, also when trainable
            Hd["bo_param"] = ob if ob.requires_grad else None
            W["head"] = Hd
        tab.run(blocks_per_desc=64)
        self.W = W
        self._frozen_tab = tab  # keeps the sources alive until the copy has run

    def _refresh_trainable(self):
        tr = [p for g in self.groups for p in g.trainable()]
        key = tuple(p.data_ptr() for p in tr)
        if key != self._train_key:
            self._train_key = key
            self._pack = K.CopyTable(self.dev)
            for g in self.groups:
                g.fill_trainable(self._pack)
        self._pack.run()

    def prepare(self):
        self._refresh_frozen()
        if self.has_trainable:
            self._refresh_trainable()
        return self.W

    def _scatter_grads(self):
        tr = [p for g in self.groups for p in g.trainable()]
        hb = self.W.get("head", {}).get("bo_param") if self.c.kind == "siglip" else None
        if hb is not None:
            tr.append(hb)
        for p in tr:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        key = tuple(p.grad.data_ptr() for p in tr)
        if key != self._grad_key:
            self._grad_key = key
            self._scatter = K.CopyTable(self.dev)
            for g in self.groups:
                g.fill_scatter(self._scatter)
            if hb is not None:
                self._scatter.add(self._head_gbo.unsqueeze(0), hb.grad.unsqueeze(0), 1, self.D, accumulate=True)
        self._scatter.run()
        # the staging is all-zero between backward calls: the tower node and the visual_projection node of one step
        # each scatter the whole table, and a stale block must not be accumulated twice
        self.staging.zero_()
        if self._head_gbo is not None:
            self._head_gbo.zero_()

    _DROP_STRIDE = 1 << 20

    def _drop_active(self) -> bool:
        training = self.model.training if hasattr(self.model, "training") else self.vm.training
        return self.p_drop > 0.0 and training

    def _drop_new_step(self, device) -> None:
        """Start of a training forward: the device-resident part of the mask offset moves on (a stream-ordered add, so a
        replay of the captured step draws fresh masks), the call index restarts."""
        if not self._drop_active():
            return
        if self._drop_base is None:
            self._drop_base = torch.zeros(1, dtype=torch.int64, device=device)
        self._drop_base += self._DROP_STRIDE
        self._drop_calls = 0

    def _drop(self):
        """(p, seed, offset, offset_base) of the next LoRA-dropout mask, or None (eval mode / p = 0).  The seed is taken
        from torch's generator once (``torch.manual_seed`` reproduces a run; the trainer re-seeds per rank); the offset is
        the call index inside the step (host counter) + a device-resident per-step base: no device sync, and the
        backward regenerates the forward's masks from the same triple."""
        if not self._drop_active():
            return None
        if self._drop_seed is None:
            self._drop_seed = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF
        self._drop_calls += 1
        return (self.p_drop, self._drop_seed, self._drop_calls, self._drop_base)

    # ---- forward ------------------------------------------------------------------------------------------
    def forward(self, pixel_values, _norm=None, save=False):
        """-> (last_hidden_state [B,T,D] bf16, pooler_output [B,D] bf16, ctx or None)."""
        c, W = self.c, self.prepare()
        B = pixel_values.shape[0]
        D, H, T, d, dp = self.D, self.H, self.T, self.d, self.dp
        eps, act = c.layer_norm_eps, self.act
        S = {} if save else None
        self._drop_new_step(pixel_values.device)
        if K.is_u8_image(pixel_values):    # the decoded uint8 HWC batch itself: u8 / 255 happens inside the gather
            if _norm is None:
                raise ValueError("a uint8 image batch needs `_norm=(mean3, std3)`: the tower takes NORMALISED pixels, and "
                                 "raw [0, 255] bytes are not that")
            img = pixel_values.contiguous()
        else:
            img = pixel_values.float().contiguous()
        mean, std = _norm if _norm is not None else (None, None)
        if IMPLICIT_PATCH_EMBED:   # image bytes -> tokens in one kernel: the im2col matrix never exists in HBM
            patch = K.patch_embed(img, W["patch_w"], c.patch_size, bias=W["patch_b"], mean=mean, std=std)
        else:
            A = K.patch_im2col(img, c.patch_size, self.patch_ld, mean, std)
            patch = K.gemm(A, W["patch_w"], bias=W["patch_b"])
        x = K.embed_assemble(patch, W["cls"], W["pos"], B, T, D)
        if c.kind == "clip":
            x, _, _ = K.layernorm_fwd(x, weight=W["pre_ln"][0], bias=W["pre_ln"][1], eps=eps, save_stats=False)
        for li, L in enumerate(self.layers):
            h, m1, r1 = K.layernorm_fwd(x, weight=W["ln1"][li][0], bias=W["ln1"][li][1], eps=eps, save_stats=save)
            h2d = h.view(-1, D)
            qkv, u_qkv = L.qkv.fwd(h2d, drop=self._drop())
            qkv = qkv.view(B, T, 3, H, dp)
            q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))
            attn = torch.empty(B, T, H * dp, dtype=BF16, device=x.device)
            lse = K.flash_attn_fwd(q, k, v, d ** -0.5, attn, want_lse=save, d_valid=d)
            xm, u_o = L.o.fwd(attn.view(-1, H * dp), drop=self._drop(), residual=x.view(-1, D))
            xm = xm.view(B, T, D)
            g, m2, r2 = K.layernorm_fwd(xm, weight=W["ln2"][li][0], bias=W["ln2"][li][1], eps=eps, save_stats=save)
            g2d = g.view(-1, D)
            pre = torch.empty(B * T, c.intermediate_size, dtype=BF16, device=x.device) if save else None
            a, u_1 = L.fc1.fwd(g2d, drop=self._drop(), act=act, aux_out=pre)
            xo, u_2 = L.fc2.fwd(a, drop=self._drop(), residual=xm.view(-1, D))
            if save:
                S[li] = dict(x=x, m1=m1, r1=r1, h=h2d, u_qkv=u_qkv, qkv=qkv, attn=attn, lse=lse, u_o=u_o, xm=xm, m2=m2,
                             r2=r2, g=g2d, pre=pre, a=a, u_1=u_1, u_2=u_2)
            x = xo.view(B, T, D)
        if c.kind == "clip":  # last_hidden_state is NOT post-layernormed; pooler = post_layernorm(h[:, 0])
            pooled, mp, rp = K.layernorm_fwd(x[:, 0:1], weight=W["post_ln"][0], bias=W["post_ln"][1], eps=eps,
                                             save_stats=save)
            if save:
                S["x_last"], S["mp"], S["rp"] = x, mp, rp
            return x, pooled[:, 0], (SimpleNamespace(t=S, B=B) if save else None)
        # SigLIP: post_layernorm on all tokens, then the MAP pooling head (modeling_siglip.py:586-654)
        xp, mp, rp = K.layernorm_fwd(x, weight=W["post_ln"][0], bias=W["post_ln"][1], eps=eps, save_stats=save)
        Hd = W["head"]
        q1 = K.gemm(Hd["probe"], Hd["wq"], bias=Hd["bq"])                                  # [1, H*dp], same for all samples
        q = q1.view(1, 1, H, dp).expand(B, 1, H, dp).contiguous().permute(0, 2, 1, 3)     # [B, H, 1, dp]
        kv = K.gemm(xp.view(-1, D), Hd["wkv"], bias=Hd["bkv"]).view(B, T, 2, H, dp)
        k, v = (kv[:, :, i].permute(0, 2, 1, 3) for i in range(2))
        o = torch.empty(B, 1, H * dp, dtype=BF16, device=x.device)
        lse = K.flash_attn_fwd(q, k, v, d ** -0.5, o, want_lse=save, d_valid=d)
        a = K.gemm(o.view(B, H * dp), Hd["wo"], bias=Hd["bo"])                              # [B, D]
        y, my, ry = K.layernorm_fwd(a, weight=Hd["ln"][0], bias=Hd["ln"][1], eps=eps, save_stats=save)
        pre = torch.empty(B, c.intermediate_size, dtype=BF16, device=x.device) if save else None
        y1, u_1 = self.h_fc1.fwd(y, drop=self._drop(), act=act, aux_out=pre)
        pooled, u_2 = self.h_fc2.fwd(y1, drop=self._drop(), residual=a)
        if save:
            S["x_last"], S["mp"], S["rp"] = x, mp, rp
            S["head"] = dict(q=q, kv=kv, o=o, lse=lse, a=a, my=my, ry=ry, y=y, pre=pre, y1=y1, u_1=u_1, u_2=u_2)
        return xp, pooled, (SimpleNamespace(t=S, B=B) if save else None)

    def project(self, pooled, save=False):
        """visual_projection (no bias), with its LoRA pair under target_modules='all-linear'."""
        x2 = pooled.to(BF16).contiguous()
        self.prepare()
        y, u = self.proj.fwd(x2, drop=self._drop())
        return y, ((x2, u) if save else None)

    def project_backward(self, saved, dy):
        x2, u = saved
        dx = self.proj.bwd(dy.to(BF16).contiguous(), x2, u)
        self._scatter_grads()
        return dx

    # ---- backward -----------------------------------------------------------------------------------------
    def backward(self, ctx, d_lhs, d_pooled):
        """Gradient chain of ``forward``: d_lhs [B,T,D] / d_pooled [B,D] (either may be None).  LoRA / bias gradients
        are accumulated into ``.grad``; nothing is returned (pixels and embeddings are frozen)."""
        c, W, S, B = self.c, self.W, ctx.t, ctx.B
        D, H, T, d, dp = self.D, self.H, self.T, self.d, self.dp
        dev = self.dev
        act = self.act
        bf = lambda t: t.to(BF16).contiguous()
        if c.kind == "clip":
            if d_lhs is not None:
                dx = bf(d_lhs).clone() if d_lhs.dtype == BF16 and d_lhs.is_contiguous() else bf(d_lhs)
            else:
                dx = torch.zeros(B, T, D, dtype=BF16, device=dev)
            if d_pooled is not None:
                dp_ = bf(d_pooled).view(B, 1, D)
                K.layernorm_bwd_dx(dp_, S["x_last"][:, 0:1], S["mp"], S["rp"], weight=W["post_ln"][0], dres=dx[:, 0:1],
                                   out=dx[:, 0:1])
        else:
            dxp = bf(d_lhs) if d_lhs is not None else None
            if d_pooled is not None:
                Hs, Hd = S["head"], W["head"]
                dpool = bf(d_pooled)
                dpre = self.h_fc2.bwd(dpool, Hs["y1"], Hs["u_2"], act=act, act_grad=True, aux_in=Hs["pre"])
                dy = self.h_fc1.bwd(dpre, Hs["y"], Hs["u_1"])
                da = K.layernorm_bwd_dx(dy, Hs["a"], Hs["my"], Hs["ry"], weight=Hd["ln"][0], dres=dpool)
                if Hd["bo_param"] is not None:
                    K.colsum(da, self._head_gbo)
                do = K.gemm(da, Hd["wo"], b_mn=True).view(B, 1, H * dp)
                dq = torch.empty(B, H, 1, dp, dtype=BF16, device=dev)
                dkv = torch.empty_like(Hs["kv"])
                k, v = (Hs["kv"][:, :, i].permute(0, 2, 1, 3) for i in range(2))
                dk, dv = (dkv[:, :, i].permute(0, 2, 1, 3) for i in range(2))
                K.flash_attn_bwd(Hs["q"], k, v, Hs["lse"], d ** -0.5, Hs["o"], do, dq, dk, dv, d_valid=d)
                kw = dict(residual=dxp.view(-1, D)) if dxp is not None else {}
                dxp = K.gemm(dkv.view(-1, 2 * H * dp), Hd["wkv"], b_mn=True, **kw).view(B, T, D)
            if dxp is None:
                dxp = torch.zeros(B, T, D, dtype=BF16, device=dev)
            dx = K.layernorm_bwd_dx(dxp, S["x_last"], S["mp"], S["rp"], weight=W["post_ln"][0])
        for li in reversed(range(len(self.layers))):
            L, s = self.layers[li], S[li]
            dx2d = dx.view(-1, D)
            dpre = L.fc2.bwd(dx2d, s["a"], s["u_2"], act=act, act_grad=True, aux_in=s["pre"])
            dg = L.fc1.bwd(dpre, s["g"], s["u_1"])
            dxm = K.layernorm_bwd_dx(dg.view(B, T, D), s["xm"], s["m2"], s["r2"], weight=W["ln2"][li][0], dres=dx)
            dattn = L.o.bwd(dxm.view(-1, D), s["attn"].view(-1, H * dp), s["u_o"])
            dqkv = torch.empty_like(s["qkv"])
            q, k, v = (s["qkv"][:, :, i].permute(0, 2, 1, 3) for i in range(3))
            dq, dk, dv = (dqkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))
            K.flash_attn_bwd(q, k, v, s["lse"], d ** -0.5, s["attn"], dattn.view(B, T, H * dp), dq, dk, dv, d_valid=d)
            need_dx = li > 0  # nothing trainable below the first layer's LN1
            dh = L.qkv.bwd(dqkv.view(-1, 3 * H * dp), s["h"], s["u_qkv"], need_dx=need_dx)
            if need_dx:
                dx = K.layernorm_bwd_dx(dh.view(B, T, D), s["x"], s["m1"], s["r1"], weight=W["ln1"][li][0], dres=dxm)
            S[li] = None  # free this layer's activations
        self._scatter_grads()


class _TowerFn(torch.autograd.Function):
    """One autograd node for the whole tower (LoRA / bias gradients are written into ``.grad`` directly)."""

    @staticmethod
    def forward(ctx, eng, pixel_values, norm, *trainable):
        lhs, pooled, ectx = eng.forward(pixel_values, norm, save=True)
        ctx.eng, ctx.ectx = eng, ectx
        return lhs, pooled

    @staticmethod
    def backward(ctx, d_lhs, d_pooled):
        ctx.eng.backward(ctx.ectx, d_lhs, d_pooled)
        ctx.ectx = None
        return (None, None, None) + (None,) * (len(ctx.needs_input_grad) - 3)


class _ProjFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng, pooled, *trainable):
        y, saved = eng.project(pooled, save=True)
        ctx.eng, ctx.saved, ctx.dt = eng, saved, pooled.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        dx = ctx.eng.project_backward(ctx.saved, dy)
        return (None, dx.to(ctx.dt)) + (None,) * (len(ctx.needs_input_grad) - 2)


def _signature(model):
    """What the prepared operands depend on besides parameter VALUES: the LoRA wrap and the requires_grad flags."""
    return (id(getattr(model, "lora_config", None)), tuple(p.requires_grad for p in model.parameters()),
            next(iter(model.parameters())).device)


def engine_of(model) -> TowerEngine:
    eng = getattr(model, "_engine", None)
    if eng is None or eng.sig != _signature(model):
        eng = TowerEngine(model)
        if isinstance(model, torch.nn.Module):
            object.__setattr__(model, "_engine", eng)  # plain attribute: not a submodule, not in the state_dict
        else:
            model._engine = eng
    return eng


def run_tower(model, pixel_values, _norm=None):
    """-> (last_hidden_state, pooler_output).  Saves activations and wires autograd only when something in the tower
    trains and gradients are enabled (stage 2); otherwise the frozen forward keeps nothing."""
    eng = engine_of(model)
    if eng.has_trainable and torch.is_grad_enabled():
        tr = [p for g in eng.groups if g is not getattr(eng, "proj", None) for p in g.trainable()]
        hb = model.vision_model.head.attention.out_proj.bias if eng.c.kind == "siglip" else None
        if hb is not None and hb.requires_grad:
            tr.append(hb)
        if tr:
            return _TowerFn.apply(eng, pixel_values, _norm, *tr)
    lhs, pooled, _ = eng.forward(pixel_values, _norm, save=False)
    return lhs, pooled


def run_projection(model, pooled):
    eng = engine_of(model)
    tr = eng.proj.trainable()
    if torch.is_grad_enabled() and (tr or pooled.requires_grad):
        return _ProjFn.apply(eng, pooled, *tr)
    return eng.project(pooled)[0]
