"""Hand-scheduled forward / backward of the lightweight FLUX DiT on the sm_100a kernels.

This is the B200-native replacement of ``Flux.forward`` + autograd for
/root/reference/Continuous/src/flux/model.py:137-228 and the live processors of
src/flux/modules/layers.py:303-337 (double stream), :485-501 (single stream), :561-572 (last layer).

Instead of ~300 small ATen kernels chained by autograd, the step is an explicit schedule of
  * tcgen05 GEMMs with fused epilogues (bias, GELU/SiLU, gate*u+residual, act' in dgrad),
  * one fused AdaLN kernel per modulated LayerNorm,
  * one QK-RMSNorm+RoPE scatter kernel per stream and one flash-attention launch per block,
with every activation needed by backward kept resident in HBM (180 GB: nothing is recomputed).
Parameter gradients are written straight into the caller's grad buffers by the wgrad GEMM epilogue
(beta = 1 when accumulating), small vectors go through one fp32 scratch that is flushed once.
"""
from __future__ import annotations

import contextlib
import os
from dataclasses import dataclass, field

import torch

from .. import kernels as K
from ..kernels import ACT_GELU_TANH, ACT_NONE, ACT_SILU, BF16, F32


@dataclass
class FluxDims:
    B: int
    Li: int
    Lt: int
    C: int
    H: int
    D: int
    mlp: int
    in_ch: int

    @property
    def L(self) -> int:
        return self.Li + self.Lt


class GradSink:
    """Where parameter gradients go.  ``big(name)`` returns the bf16/fp32 [N,K] grad tensor a wgrad GEMM writes
    (accumulating when ``accumulate``); ``small(name)`` returns an fp32 accumulator view flushed by ``flush()``."""

    # Weight gradients are leaves of the backward: dW = dY^T X (and db = colsum dY) of a layer is needed by nothing
    # until the optimizer, while dX = dY W is the critical path.  With ``overlap_wgrad`` they are issued on a side
    # stream (forked after dY is ready, joined before the block is flushed / announced to the reducer): the persistent
    # GEMM grids of the two streams then overlap at their edges -- an SM that has finished its share of one kernel
    # starts on the other's instead of idling until the kernel boundary (the dgrad with 9.08 waves of tiles left 92 % of
    # the SMs idle for its last wave).  Inside the step graph the side stream is a forked branch.
    overlap_wgrad = os.environ.get("GH_WGRAD_STREAM", "1") != "0"
    _side_streams: dict = {}

    def __init__(self, params: dict[str, torch.Tensor], accumulate: bool, on_ready=None):
        self.params = params
        self.accumulate = accumulate
        self.on_ready = on_ready
        self._keep: list = []       # operands of side-stream kernels, kept alive until join()
        self._forked = False
        self._flushed: set[str] = set()
        self._discard = None
        self._small: dict[str, torch.Tensor] = {}
        self._touched: set[str] = set()
        n_small = sum(p.numel() for n, p in params.items() if p.dim() == 1 and p.requires_grad)
        dev = next(iter(params.values())).device
        self._scratch = torch.zeros(max(n_small, 1), dtype=F32, device=dev)
        off = 0
        for n, p in params.items():
            if p.dim() == 1 and p.requires_grad:
                self._small[n] = self._scratch[off:off + p.numel()]
                off += p.numel()

    def wants(self, name: str) -> bool:
        return self.params[name].requires_grad

    def side(self, *operands):
        """Context in which the caller launches leaf work (wgrad GEMM, bias column sum) that reads ``operands``."""
        dev = self._scratch.device
        if not GradSink.overlap_wgrad or dev.type != "cuda":
            return contextlib.nullcontext()
        st = GradSink._side_streams.get(dev.index)
        if st is None:
            st = GradSink._side_streams[dev.index] = torch.cuda.Stream(device=dev)
        st.wait_stream(torch.cuda.current_stream(dev))     # dY (and the fp32 scratch) are ready
        self._keep.extend(operands)
        self._forked = True
        return torch.cuda.stream(st)

    def join(self) -> None:
        """The current stream waits for the side stream; the kept operands may be freed."""
        if self._forked:
            dev = self._scratch.device
            torch.cuda.current_stream(dev).wait_stream(GradSink._side_streams[dev.index])
            self._forked = False
        self._keep.clear()

    def big(self, name: str):
        """-> (grad tensor, residual-or-None) for a weight matrix."""
        p = self.params[name]
        first = p.grad is None or (not self.accumulate and name not in self._touched)
        if p.grad is None:
            p.grad = torch.empty_like(p)
        self._touched.add(name)
        return p.grad, (None if first else p.grad)

    def small(self, name: str) -> torch.Tensor:
        acc = self._small.get(name)
        if acc is None:  # frozen parameter (stage2_only trains the tower alone): its gradient goes to a discard buffer
            n = self.params[name].numel()
            if self._discard is None or self._discard.numel() < n:
                self._discard = torch.zeros(n, dtype=F32, device=self._scratch.device)
            acc = self._discard[:n]
        return acc

    def flush(self, prefix: str | None = None) -> None:
        """Write the fp32 accumulators of the 1-D parameters under ``prefix`` (all when None) into ``.grad`` and
        tell ``on_ready`` (the data-parallel bucket launcher) that every gradient under ``prefix`` is final."""
        self.join()
        for n, acc in self._small.items():
            if n in self._flushed or (prefix is not None and not n.startswith(prefix)):
                continue
            self._flushed.add(n)
            p = self.params[n]
            first = p.grad is None or not self.accumulate
            if p.grad is None:
                p.grad = torch.empty_like(p)
            K.accum_cast(acc, p.grad, 1.0, accumulate=not first)
        if self.on_ready is not None and prefix is not None:
            self.on_ready(prefix)


def _lin_fwd(x2d, w, b, **kw):
    return K.gemm(x2d, w, bias=b, **kw)


# The conditioning side of the DiT -- the 11 Modulation GEMMs (all functions of `vec` alone) and the txt stream of the
# double blocks (B rows in image mode) -- is a chain of 32-row GEMMs that each stream a whole weight matrix through
# 48-144 CTAs in 25-70 us: HBM/latency-bound work that leaves the tensor pipe idle.  It runs on its own stream beside
# the img stream (which is tensor-bound) and meets it only where the reference's dataflow does: the joint attention
# (q/k/v of both streams) and the concat before the single blocks.  Inside the step graph it is a forked branch.
# GH_COND_STREAM=0 switches it off (everything on the caller's stream).
OVERLAP_COND = os.environ.get("GH_COND_STREAM", "1") != "0"
_COND_STREAMS: dict = {}


def _cond_stream(dev):
    st = _COND_STREAMS.get(dev.index)
    if st is None:
        st = _COND_STREAMS[dev.index] = torch.cuda.Stream(device=dev)
    return st


def _bias_acc(sink: GradSink, b_name: str):
    """The fp32 accumulator of a bias gradient when the kernel that PRODUCES dY also sums its columns (gh_gate_bwd)."""
    return sink.small(b_name) if sink.wants(b_name) else None


def _lin_bwd(dy2d, x2d, w_name, b_name, P, sink: GradSink, need_dx=True, dx_kw=None, bias_done=False):
    """dx = dy @ W ; dW (+)= dy^T x ; db += colsum(dy)  (``bias_done``: db already came out of the producer of dy)."""
    w = P[w_name]
    want_w, want_b = sink.wants(w_name), b_name is not None and sink.wants(b_name) and not bias_done
    if want_w or want_b:
        with sink.side(dy2d, x2d):      # leaves of the backward: off the critical path
            if want_w:
                gw, res = sink.big(w_name)
                K.gemm(dy2d, x2d, a_mn=True, b_mn=True, out=gw, residual=res)
            if want_b:
                K.colsum(dy2d, sink.small(b_name))
    dx = None
    if need_dx:
        dx = K.gemm(dy2d, w, b_mn=True, **(dx_kw or {}))
    return dx


@dataclass
class FluxCtx:
    dims: FluxDims
    t: dict = field(default_factory=dict)  # saved tensors by name


def flux_forward(P: dict, cfg, img, img_ids, txt, txt_ids, timesteps, y, guidance, save: bool = True):
    """P: name -> bf16 parameter tensor (reference state_dict names).  Returns (pred [B,Li,in] bf16, ctx)."""
    B, Li, in_ch = img.shape
    Lt = txt.shape[1]
    C, H = cfg.hidden_size, cfg.num_heads
    D = C // H
    mlp = int(C * cfg.mlp_ratio)
    dm = FluxDims(B, Li, Lt, C, H, D, mlp, in_ch)
    L = dm.L
    ctx = FluxCtx(dm)
    S = ctx.t
    dev = img.device
    scale = D ** -0.5

    img2d = img.reshape(B * Li, in_ch).to(BF16).contiguous()
    txt2d = txt.reshape(B * Lt, -1).to(BF16).contiguous()
    y2d = y.to(BF16).contiguous()
    S["img2d"], S["txt2d"], S["y2d"] = img2d, txt2d, y2d

    x_img = _lin_fwd(img2d, P["img_in.weight"], P["img_in.bias"])
    x_txt = _lin_fwd(txt2d, P["txt_in.weight"], P["txt_in.bias"])

    # vec = time_in(temb(t)) + guidance_in(temb(g)) + vector_in(y)      (model.py:155-160)
    def embedder(name, inp, residual):
        pre = torch.empty(B, C, dtype=BF16, device=dev)
        h = _lin_fwd(inp, P[f"{name}.in_layer.weight"], P[f"{name}.in_layer.bias"], act=ACT_SILU, aux_out=pre)
        S[f"{name}.in"], S[f"{name}.pre"], S[f"{name}.h"] = inp, pre, h
        return _lin_fwd(h, P[f"{name}.out_layer.weight"], P[f"{name}.out_layer.bias"], residual=residual)

    temb = K.timestep_embedding(timesteps.float(), round_bf16=(timesteps.dtype != F32))
    vec = embedder("time_in", temb, None)
    if cfg.guidance_embed:
        if guidance is None:
            raise ValueError("Didn't get guidance strength for guidance distilled model.")
        gemb = K.timestep_embedding(guidance.float(), round_bf16=(guidance.dtype != F32))
        vec = embedder("guidance_in", gemb, vec)
    vec = embedder("vector_in", y2d, vec)
    svec = K.act_fwd(vec, ACT_SILU)
    S["vec"], S["svec"] = vec, svec

    ids = torch.cat((txt_ids, img_ids), dim=1)
    cs = K.rope_table(ids, tuple(cfg.axes_dim), float(cfg.theta))  # [B, L, D/2, 2]
    S["cs"] = cs

    def adaln(name, x3d, shift, scl):
        h, mean, rstd = K.layernorm_fwd(x3d, shift=shift, scale=scl, eps=1e-6)
        S[f"{name}.x"], S[f"{name}.h"], S[f"{name}.mean"], S[f"{name}.rstd"] = x3d, h, mean, rstd
        return h

    x_img = x_img.view(B, Li, C)
    x_txt = x_txt.view(B, Lt, C)

    # ---- every Modulation of the network: functions of svec alone (layers.py:162-175) ----
    main = torch.cuda.current_stream(dev) if dev.type == "cuda" else None
    # the txt stream only pays as a separate branch while it is a handful of rows (image mode: Lt = 1); in the video
    # modes it is 1152-1728 tokens per sample, tensor-bound like the img stream
    side = _cond_stream(dev) if (OVERLAP_COND and main is not None) else None
    txt_side = side if (side is not None and B * Lt <= 256) else None
    mod_names = [(f"double_blocks.{bi}.{s}_mod", f"double_blocks.{bi}.{s}_mod.lin", 6) for bi in range(cfg.depth)
                 for s in ("img", "txt")]
    mod_names += [(f"single_blocks.{bi}.mod", f"single_blocks.{bi}.modulation.lin", 3) for bi in range(cfg.depth_single_blocks)]
    mod_names += [("final.mod", "final_layer.adaLN_modulation.1", 2)]
    mod_out = {key: torch.empty(B, n * C, dtype=BF16, device=dev) for key, _, n in mod_names}   # (allocated on the caller's stream)
    mod_ready = {}
    if side is not None:
        side.wait_stream(main)                      # svec
    with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
        for key, lin, _ in mod_names:
            _lin_fwd(svec, P[f"{lin}.weight"], P[f"{lin}.bias"], out=mod_out[key])
            S[key] = mod_out[key]
            if side is not None:
                mod_ready[key] = torch.cuda.Event()
                mod_ready[key].record(side)

    def mod_of(key, stream=None):
        """The modulation tensor ``key``; the consumer's stream waits for the GEMM that produces it."""
        ev = mod_ready.get(key)
        if ev is not None:
            (stream or torch.cuda.current_stream(dev)).wait_event(ev)
        return mod_out[key]

    def on_txt_stream():
        return torch.cuda.stream(txt_side) if txt_side is not None else contextlib.nullcontext()

    if txt_side is not None:
        txt_side.wait_stream(main)                  # x_txt (txt_in ran on the caller's stream)
    for bi in range(cfg.depth):
        p = f"double_blocks.{bi}"
        mods = {"img": mod_of(f"{p}.img_mod")}
        with on_txt_stream():
            mods["txt"] = mod_of(f"{p}.txt_mod")
        q = torch.empty(B, H, L, D, dtype=BF16, device=dev)
        k = torch.empty_like(q)
        v = torch.empty_like(q)
        if txt_side is not None:
            txt_side.wait_stream(main)              # q/k/v are fresh allocations: whatever last used that memory is done
        xs = {"img": x_img, "txt": x_txt}
        for s, l_off in (("txt", 0), ("img", Lt)):
            with (on_txt_stream() if s == "txt" else contextlib.nullcontext()):
                m = mods[s]
                h = adaln(f"{p}.{s}_norm1", xs[s], m[:, 0:C], m[:, C:2 * C])
                qkv = _lin_fwd(h.view(-1, C), P[f"{p}.{s}_attn.qkv.weight"], P[f"{p}.{s}_attn.qkv.bias"])
                qkv = qkv.view(B, -1, 3 * C)
                S[f"{p}.{s}_qkv"] = qkv
                K.qk_norm_rope_fwd(qkv, H, P[f"{p}.{s}_attn.norm.query_norm.scale"],
                                   P[f"{p}.{s}_attn.norm.key_norm.scale"], cs, q, k, v, l_off)
        attn_t = torch.empty(B, Lt, C, dtype=BF16, device=dev)
        attn_i = torch.empty(B, Li, C, dtype=BF16, device=dev)
        if txt_side is not None:
            main.wait_stream(txt_side)              # the txt rows of q / k / v
        lse = K.flash_attn_fwd(q, k, v, scale, attn_i, attn_t, Lt)
        if txt_side is not None:
            txt_side.wait_stream(main)              # attn_t
        S[f"{p}.q"], S[f"{p}.k"], S[f"{p}.v"], S[f"{p}.lse"] = q, k, v, lse
        S[f"{p}.attn_t"], S[f"{p}.attn_i"] = attn_t, attn_i
        def after_attention(s, attn, rpb):
            m = mods[s]
            x = xs[s]
            u1 = torch.empty(B * rpb, C, dtype=BF16, device=dev)
            x = _lin_fwd(attn.view(-1, C), P[f"{p}.{s}_attn.proj.weight"], P[f"{p}.{s}_attn.proj.bias"],
                         gate=m[:, 2 * C:3 * C], rows_per_batch=rpb, residual=x.view(-1, C), aux_out=u1).view(B, rpb, C)
            S[f"{p}.{s}_u1"] = u1
            h2 = adaln(f"{p}.{s}_norm2", x, m[:, 3 * C:4 * C], m[:, 4 * C:5 * C])
            pre = torch.empty(B * rpb, mlp, dtype=BF16, device=dev)
            a = _lin_fwd(h2.view(-1, C), P[f"{p}.{s}_mlp.0.weight"], P[f"{p}.{s}_mlp.0.bias"], act=ACT_GELU_TANH,
                         aux_out=pre)
            u2 = torch.empty(B * rpb, C, dtype=BF16, device=dev)
            x = _lin_fwd(a, P[f"{p}.{s}_mlp.2.weight"], P[f"{p}.{s}_mlp.2.bias"], gate=m[:, 5 * C:6 * C],
                         rows_per_batch=rpb, residual=x.view(-1, C), aux_out=u2).view(B, rpb, C)
            S[f"{p}.{s}_mlp_pre"], S[f"{p}.{s}_mlp_a"], S[f"{p}.{s}_u2"] = pre, a, u2
            xs[s] = x

        after_attention("img", attn_i, Li)
        with on_txt_stream():
            after_attention("txt", attn_t, Lt)
        x_img, x_txt = xs["img"], xs["txt"]

    if txt_side is not None:
        main.wait_stream(txt_side)                  # x_txt of the last double block
    x = torch.empty(B, L, C, dtype=BF16, device=dev)
    x[:, :Lt].copy_(x_txt)
    x[:, Lt:].copy_(x_img)
    for bi in range(cfg.depth_single_blocks):
        p = f"single_blocks.{bi}"
        m = mod_of(f"{p}.mod")  # [B, 3C]
        h = adaln(f"{p}.pre_norm", x, m[:, 0:C], m[:, C:2 * C])
        w1, b1 = P[f"{p}.linear1.weight"], P[f"{p}.linear1.bias"]
        qkv = _lin_fwd(h.view(-1, C), w1[:3 * C], b1[:3 * C]).view(B, L, 3 * C)
        catb = torch.empty(B * L, C + mlp, dtype=BF16, device=dev)  # [attn | gelu(mlp)]  (layers.py:499)
        pre = torch.empty(B * L, mlp, dtype=BF16, device=dev)
        _lin_fwd(h.view(-1, C), w1[3 * C:], b1[3 * C:], act=ACT_GELU_TANH, aux_out=pre, out=catb[:, C:])
        q = torch.empty(B, H, L, D, dtype=BF16, device=dev)
        k = torch.empty_like(q)
        v = torch.empty_like(q)
        K.qk_norm_rope_fwd(qkv, H, P[f"{p}.norm.query_norm.scale"], P[f"{p}.norm.key_norm.scale"], cs, q, k, v, 0)
        attn_view = catb.view(B, L, C + mlp)[:, :, :C]
        lse = K.flash_attn_fwd(q, k, v, scale, attn_view)
        u = torch.empty(B * L, C, dtype=BF16, device=dev)
        x = _lin_fwd(catb, P[f"{p}.linear2.weight"], P[f"{p}.linear2.bias"], gate=m[:, 2 * C:3 * C], rows_per_batch=L,
                     residual=x.view(-1, C), aux_out=u).view(B, L, C)
        S[f"{p}.qkv"], S[f"{p}.cat"], S[f"{p}.mlp_pre"], S[f"{p}.u"] = qkv, catb, pre, u
        S[f"{p}.q"], S[f"{p}.k"], S[f"{p}.v"], S[f"{p}.lse"] = q, k, v, lse

    mf = mod_of("final.mod")
    if side is not None:
        main.wait_stream(side)                      # (every branch of the forward has joined the caller's stream)
    hf = adaln("final.norm", x[:, Lt:], mf[:, 0:C], mf[:, C:2 * C])  # chunk order: shift, scale (layers.py:569)
    pred = _lin_fwd(hf.view(-1, C), P["final_layer.linear.weight"], P["final_layer.linear.bias"])
    return pred.view(B, Li, in_ch), ctx


def flux_backward(P: dict, cfg, ctx: FluxCtx, dpred: torch.Tensor, sink: GradSink,
                  need_dtxt: bool = True, need_dy: bool = True, need_dimg: bool = False):
    """Backward of flux_forward.  Returns (d_img, d_txt, d_y) (bf16, shaped like the inputs) or None each."""
    dm, S = ctx.dims, ctx.t
    B, Li, Lt, C, H, D, mlp = dm.B, dm.Li, dm.Lt, dm.C, dm.H, dm.D, dm.mlp
    L = dm.L
    dev = dpred.device
    scale = D ** -0.5
    svec = S["svec"]
    dsvec = torch.zeros(B, C, dtype=F32, device=dev)  # every Modulation feeds it

    def mod_bwd(dmod_acc, w_name, b_name):
        """dmod_acc fp32 [B, n] -> param grads of the modulation linear, dsvec += dmod @ W.  All of it is a leaf of the
        backward (dsvec is read only after the last block), so it rides the wgrad side stream."""
        dmod = torch.empty(dmod_acc.shape, dtype=BF16, device=dev)
        with sink.side(dmod_acc, dmod):
            K.accum_cast(dmod_acc, dmod)
            # [B, n] x [n, C] with B = 32 rows: one M tile and 12-48 N tiles would leave most SMs idle while each CTA
            # streams the whole K = n (up to 18432) alone -> split-K over the machine, partials red.add-ed into dsvec
            K.gemm(dmod, P[w_name], b_mn=True, out=dsvec, k_splits=-1)
            if sink.wants(w_name):
                gw, res = sink.big(w_name)
                K.gemm(dmod, svec, a_mn=True, b_mn=True, out=gw, residual=res)
            if sink.wants(b_name):
                K.colsum(dmod, sink.small(b_name))

    def adaln_bwd(name, dh3d, mod, sh_off, dmod_acc, dres=None, out=None):
        x, mean, rstd = S[f"{name}.x"], S[f"{name}.mean"], S[f"{name}.rstd"]
        K.layernorm_bwd_params(dh3d, x, mean, rstd, dmod_acc[:, sh_off:sh_off + C], dmod_acc[:, sh_off + C:sh_off + 2 * C])
        return K.layernorm_bwd_dx(dh3d, x, mean, rstd, scale=mod[:, sh_off + C:sh_off + 2 * C], dres=dres, out=out)

    # ---- last layer ----
    dpred2d = dpred.reshape(B * Li, -1).to(BF16).contiguous()
    hf = S["final.norm.h"]
    dhf = _lin_bwd(dpred2d, hf.view(-1, C), "final_layer.linear.weight", "final_layer.linear.bias", P, sink)
    dmodf = torch.zeros(B, 2 * C, dtype=F32, device=dev)
    dx = torch.zeros(B, L, C, dtype=BF16, device=dev)  # txt rows of the last single block get no gradient
    adaln_bwd("final.norm", dhf.view(B, Li, C), S["final.mod"], 0, dmodf, out=dx[:, Lt:])
    mod_bwd(dmodf, "final_layer.adaLN_modulation.1.weight", "final_layer.adaLN_modulation.1.bias")
    sink.flush("final_layer.")

    # ---- single-stream blocks ----
    for bi in reversed(range(cfg.depth_single_blocks)):
        p = f"single_blocks.{bi}"
        m = S[f"{p}.mod"]
        dmod = torch.zeros(B, 3 * C, dtype=F32, device=dev)
        catb, pre, qkv = S[f"{p}.cat"], S[f"{p}.mlp_pre"], S[f"{p}.qkv"]
        # out = x + gate * u
        du = K.gate_bwd(dx, S[f"{p}.u"].view(B, L, C), m[:, 2 * C:3 * C], dmod[:, 2 * C:3 * C],
                        dbias_acc=_bias_acc(sink, f"{p}.linear2.bias"))
        # linear2: u = cat @ W2^T + b2 ;  dcat[:, :C] = dattn, dcat[:, C:] = da -> * gelu'(pre) fused in the dgrad epilogue
        du2d = du.view(-1, C)
        w2 = P[f"{p}.linear2.weight"]
        dl1 = torch.empty(B * L, 3 * C + mlp, dtype=BF16, device=dev)  # d(linear1 output) = [dqkv | dpre]
        with sink.side(du2d, catb):
            if sink.wants(f"{p}.linear2.weight"):
                gw, res = sink.big(f"{p}.linear2.weight")
                K.gemm(du2d, catb, a_mn=True, b_mn=True, out=gw, residual=res)
        dattn = K.gemm(du2d, w2[:, :C], b_mn=True)
        K.gemm(du2d, w2[:, C:], b_mn=True, act=ACT_GELU_TANH, act_grad=True, aux_in=pre, out=dl1[:, 3 * C:])
        # attention
        dq = torch.empty(B, H, L, D, dtype=BF16, device=dev)
        dk = torch.empty_like(dq)
        dv = torch.empty_like(dq)
        attn_view = catb.view(B, L, C + mlp)[:, :, :C]
        K.flash_attn_bwd(S[f"{p}.q"], S[f"{p}.k"], S[f"{p}.v"], S[f"{p}.lse"], scale, attn_view, dattn.view(B, L, C),
                         dq, dk, dv)
        gq = sink.small(f"{p}.norm.query_norm.scale")
        gk = sink.small(f"{p}.norm.key_norm.scale")
        K.qk_norm_rope_bwd(dq, dk, dv, qkv, H, P[f"{p}.norm.query_norm.scale"], P[f"{p}.norm.key_norm.scale"], S["cs"], 0,
                           dl1.view(B, L, 3 * C + mlp), gq, gk)
        h = S[f"{p}.pre_norm.h"]
        dh = _lin_bwd(dl1, h.view(-1, C), f"{p}.linear1.weight", f"{p}.linear1.bias", P, sink)
        dx = adaln_bwd(f"{p}.pre_norm", dh.view(B, L, C), m, 0, dmod, dres=dx)
        mod_bwd(dmod, f"{p}.modulation.lin.weight", f"{p}.modulation.lin.bias")
        sink.flush(p + ".")

    # ---- split back into the two streams ----
    dxs = {"txt": dx[:, :Lt], "img": dx[:, Lt:]}
    for bi in reversed(range(cfg.depth)):
        p = f"double_blocks.{bi}"
        dmods = {s: torch.zeros(B, 6 * C, dtype=F32, device=dev) for s in ("img", "txt")}
        dattn = {}
        dx_mid = {}
        for s, rpb in (("img", Li), ("txt", Lt)):
            m = S[f"{p}.{s}_mod"]
            dxo = dxs[s]
            # x_out = x_mid + gate2 * u2 ; u2 = a @ W2^T + b2 ; a = gelu(pre) ; pre = h2 @ W1^T + b1
            du2 = K.gate_bwd(dxo, S[f"{p}.{s}_u2"].view(B, rpb, C), m[:, 5 * C:6 * C], dmods[s][:, 5 * C:6 * C],
                             dbias_acc=_bias_acc(sink, f"{p}.{s}_mlp.2.bias"))
            dpre = _lin_bwd(du2.view(-1, C), S[f"{p}.{s}_mlp_a"], f"{p}.{s}_mlp.2.weight", f"{p}.{s}_mlp.2.bias", P, sink,
                            dx_kw=dict(act=ACT_GELU_TANH, act_grad=True, aux_in=S[f"{p}.{s}_mlp_pre"]), bias_done=True)
            h2 = S[f"{p}.{s}_norm2.h"]
            dh2 = _lin_bwd(dpre, h2.view(-1, C), f"{p}.{s}_mlp.0.weight", f"{p}.{s}_mlp.0.bias", P, sink)
            dxm = adaln_bwd(f"{p}.{s}_norm2", dh2.view(B, rpb, C), m, 3 * C, dmods[s], dres=dxo)
            # x_mid = x_in + gate1 * u1 ; u1 = attn @ Wp^T + bp
            du1 = K.gate_bwd(dxm, S[f"{p}.{s}_u1"].view(B, rpb, C), m[:, 2 * C:3 * C], dmods[s][:, 2 * C:3 * C],
                             dbias_acc=_bias_acc(sink, f"{p}.{s}_attn.proj.bias"))
            attn = S[f"{p}.attn_i"] if s == "img" else S[f"{p}.attn_t"]
            dattn[s] = _lin_bwd(du1.view(-1, C), attn.view(-1, C), f"{p}.{s}_attn.proj.weight",
                                f"{p}.{s}_attn.proj.bias", P, sink, bias_done=True).view(B, rpb, C)
            dx_mid[s] = dxm
            if bi == 0 and sink.on_ready is not None:
                # the LAST block of the backward: announce its sub-modules as their gradients become final, so the
                # exposed tail of the data-parallel exchange is one Modulation's worth instead of the whole block
                sink.flush(f"{p}.{s}_mlp.")
        dq = torch.empty(B, H, L, D, dtype=BF16, device=dev)
        dk = torch.empty_like(dq)
        dv = torch.empty_like(dq)
        K.flash_attn_bwd(S[f"{p}.q"], S[f"{p}.k"], S[f"{p}.v"], S[f"{p}.lse"], scale, S[f"{p}.attn_i"], dattn["img"],
                         dq, dk, dv, S[f"{p}.attn_t"], dattn["txt"], Lt)
        for s, l_off, rpb in (("txt", 0, Lt), ("img", Lt, Li)):
            m = S[f"{p}.{s}_mod"]
            qkv = S[f"{p}.{s}_qkv"]
            dqkv = torch.empty_like(qkv)
            K.qk_norm_rope_bwd(dq, dk, dv, qkv, H, P[f"{p}.{s}_attn.norm.query_norm.scale"],
                               P[f"{p}.{s}_attn.norm.key_norm.scale"], S["cs"], l_off, dqkv,
                               sink.small(f"{p}.{s}_attn.norm.query_norm.scale"),
                               sink.small(f"{p}.{s}_attn.norm.key_norm.scale"))
            h = S[f"{p}.{s}_norm1.h"]
            dh = _lin_bwd(dqkv.view(-1, 3 * C), h.view(-1, C), f"{p}.{s}_attn.qkv.weight", f"{p}.{s}_attn.qkv.bias", P, sink)
            if bi == 0 and sink.on_ready is not None:
                sink.flush(f"{p}.{s}_attn.")
            dxs[s] = adaln_bwd(f"{p}.{s}_norm1", dh.view(B, rpb, C), m, 0, dmods[s], dres=dx_mid[s])
            mod_bwd(dmods[s], f"{p}.{s}_mod.lin.weight", f"{p}.{s}_mod.lin.bias")
            if bi == 0 and sink.on_ready is not None:
                sink.flush(f"{p}.{s}_mod.")
        sink.flush(p + ".")

    # ---- input projections and the conditioning vector ----
    d_img = _lin_bwd(dxs["img"].reshape(-1, C), S["img2d"], "img_in.weight", "img_in.bias", P, sink, need_dx=need_dimg)
    d_txt = _lin_bwd(dxs["txt"].reshape(-1, C), S["txt2d"], "txt_in.weight", "txt_in.bias", P, sink, need_dx=need_dtxt)
    sink.join()                                     # every Modulation's contribution to dsvec
    dsv = torch.empty(B, C, dtype=BF16, device=dev)
    K.accum_cast(dsvec, dsv)
    dvec = K.act_bwd(dsv, S["vec"], ACT_SILU)  # vec = sum of the three embedders -> same gradient to each
    d_y = None
    names = ["time_in", "vector_in"] + (["guidance_in"] if cfg.guidance_embed else [])
    for name in names:
        dpre = _lin_bwd(dvec, S[f"{name}.h"], f"{name}.out_layer.weight", f"{name}.out_layer.bias", P, sink,
                        dx_kw=dict(act=ACT_SILU, act_grad=True, aux_in=S[f"{name}.pre"]))
        want_in = name == "vector_in" and need_dy
        din = _lin_bwd(dpre, S[f"{name}.in"], f"{name}.in_layer.weight", f"{name}.in_layer.bias", P, sink,
                       need_dx=want_in)
        if want_in:
            d_y = din
    sink.flush("")
    return (d_img.view(B, Li, -1) if d_img is not None else None,
            d_txt.view(B, Lt, -1) if d_txt is not None else None, d_y)
