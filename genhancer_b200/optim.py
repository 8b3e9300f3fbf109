"""Flat parameter / gradient storage and the fused optimizer step.

Replaces, for the hot path, what the reference gets from ``torch.optim.AdamW`` + ``accelerator.clip_grad_norm_``
(/root/reference/Continuous/train_SigLIP_stage1.py:147-153,271-275): ~110 foreach launches over 1.33 B
parameters become ONE norm kernel + ONE AdamW kernel per dtype group, streaming contiguous HBM.

Layout in HBM (one ``FlatGroup`` per dtype, in module registration order so that a DiT block's parameters are
one contiguous range -- the unit of the data-parallel gradient all-reduce, see ``parallel.py``):

    params   [ p0 | pad | p1 | pad | ... ]      every tensor starts on a 128-byte boundary
    grads    same offsets; ``param.grad`` is a permanent view (never None, never reallocated)
    exp_avg / exp_avg_sq   same offsets, same dtype as the parameters (bf16 states for the bf16 DiT, as the
                           reference's non-DeepSpeed path: SURVEY.md Q7)
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch

from . import kernels as K

_ALIGN_BYTES = 128


@dataclass(eq=False)
class FlatGroup:
    dtype: torch.dtype
    names: list[str] = field(default_factory=list)
    params: list[torch.nn.Parameter] = field(default_factory=list)
    offsets: list[int] = field(default_factory=list)  # element offsets
    numel: int = 0
    flat_p: torch.Tensor | None = None
    flat_g: torch.Tensor | None = None
    exp_avg: torch.Tensor | None = None
    exp_avg_sq: torch.Tensor | None = None

    def range_of(self, prefix: str) -> tuple[int, int] | None:
        """[start, end) element range covering every parameter whose name starts with ``prefix``."""
        idx = [i for i, n in enumerate(self.names) if n.startswith(prefix)]
        if not idx:
            return None
        lo, hi = min(idx), max(idx)
        return self.offsets[lo], self.offsets[hi] + self.params[hi].numel()


def flatten(named_params, device=None) -> list[FlatGroup]:
    """Move the given trainable parameters into flat per-dtype buffers (param.data and param.grad become views).
    Call AFTER the final ``.to(device/dtype)`` of the modules."""
    groups: dict[torch.dtype, FlatGroup] = {}
    for name, p in named_params:
        if not p.requires_grad:
            continue
        g = groups.setdefault(p.dtype, FlatGroup(p.dtype))
        align = _ALIGN_BYTES // p.element_size()
        g.numel = (g.numel + align - 1) // align * align
        g.names.append(name)
        g.params.append(p)
        g.offsets.append(g.numel)
        g.numel += p.numel()
    out = []
    for g in groups.values():
        dev = device or g.params[0].device
        total = (g.numel + 63) // 64 * 64
        g.flat_p = torch.zeros(total, dtype=g.dtype, device=dev)
        g.flat_g = torch.zeros(total, dtype=g.dtype, device=dev)
        with torch.no_grad():
            for p, off in zip(g.params, g.offsets):
                view = g.flat_p[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                p.grad = g.flat_g[off:off + p.numel()].view(p.shape)
        g.numel = total
        out.append(g)
    return out


class FusedAdamW:
    """AdamW + clip-by-global-norm on flat groups, all on the device (no host sync, no per-tensor launches).

    ``step(grad_scale)``: grad_scale multiplies every gradient first (1/gradient_accumulation_steps, or 1/world
    when the all-reduce summed).  ``engine_managed`` modules (Flux) overwrite their gradients on the first backward
    after ``zero_grad()``; the other groups are memset."""

    def __init__(self, groups: list[FlatGroup], lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.01, max_grad_norm: float = 1.0, engine_managed=()):
        self.groups = groups
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.max_grad_norm = max_grad_norm
        self.engine_managed = list(engine_managed)
        self.step_count = 0
        for g in groups:
            g.exp_avg = torch.zeros_like(g.flat_p)
            g.exp_avg_sq = torch.zeros_like(g.flat_p)
        self.gnorm_sq = torch.zeros(1, dtype=torch.float32, device=groups[0].flat_p.device)
        self._sumsq_ws = None       # scratch of the deterministic norm reduction (allocated on the first step, on the device)
        # device-resident [step count, update pending]: what a captured update reads (see step_captured)
        self.dev_state = torch.zeros(2, dtype=torch.int32, device=groups[0].flat_p.device)
        self._managed_dtypes = {p.dtype for m in self.engine_managed for p in m.parameters()}
        self.zero_grad()

    def _ws(self) -> torch.Tensor:
        if self._sumsq_ws is None:   # private to this optimizer: its calls are ordered on whichever stream runs the update
            from . import _lib
            self._sumsq_ws = torch.zeros(_lib.lib().gh_sumsq_workspace_bytes() // 4, dtype=torch.float32,
                                         device=self.gnorm_sq.device)
        return self._sumsq_ws

    def zero_grad(self) -> None:
        for m in self.engine_managed:
            m._grad_overwrite = True
        for g in self.groups:
            if g.dtype not in self._managed_dtypes:
                g.flat_g.zero_()

    def grad_norm(self) -> torch.Tensor:
        """Device scalar: global L2 norm of the (unscaled) gradients -- valid after ``step``."""
        return self.gnorm_sq.sqrt()

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0) -> None:
        self.step_count += 1
        self.gnorm_sq.zero_()
        for g in self.groups:
            K.sumsq_accum(g.flat_g, self.gnorm_sq, self._ws())
        for g in self.groups:
            K.adamw_step(g.flat_p, g.flat_g, g.exp_avg, g.exp_avg_sq, self.lr, self.betas[0], self.betas[1], self.eps,
                         self.weight_decay, self.step_count, self.gnorm_sq, self.max_grad_norm, grad_scale)

    @torch.no_grad()
    def step_captured(self, grad_scale: float = 1.0) -> None:
        """The same update in a form a CUDA graph can replay (``graph.PipelinedTrainStep``): the step count lives in
        ``dev_state[0]`` and advances by ``dev_state[1]`` (1 = the gradients of a finished backward are waiting,
        0 = nothing pending: every kernel below is then a no-op on the parameters).  The caller captures
        ``mark_pending()`` after the backward and keeps ``step_count`` (the host mirror) in step with the replays."""
        self.dev_state[0:1] += self.dev_state[1:2]
        self.gnorm_sq.zero_()
        for g in self.groups:
            K.sumsq_accum(g.flat_g, self.gnorm_sq, self._ws())
        for g in self.groups:
            K.adamw_step(g.flat_p, g.flat_g, g.exp_avg, g.exp_avg_sq, self.lr, self.betas[0], self.betas[1], self.eps,
                         self.weight_decay, 0, self.gnorm_sq, self.max_grad_norm, grad_scale, dev_state=self.dev_state)

    def mark_pending(self, pending: bool = True) -> None:
        """dev_state[1] = pending (a stream-ordered fill: capturable)."""
        self.dev_state[1:2].fill_(1 if pending else 0)

    def sync_device_state(self) -> None:
        """After an eager ``step()``: the device copy of the step count follows the host's, nothing is pending."""
        self.dev_state.copy_(torch.tensor([self.step_count, 0], dtype=torch.int32), non_blocking=False)

    # ---- checkpoint layout of the reference: optimizer-state-{N}.bin = torch optimizer.state_dict() ----------
    def state_dict(self) -> dict:
        """torch.optim.AdamW.state_dict() layout (tensors on the host).  Parameter indices follow the flat groups
        (per dtype, in module order: bf16 DiT first, then the fp32 projectors / adapter), not the reference's
        ``super_model.parameters()`` order: the file round-trips through ``load_state_dict`` here, it is not
        index-compatible with a torch optimizer built over the reference's module."""
        state, idx = {}, 0
        for g in self.groups:
            for p, off in zip(g.params, g.offsets):
                sl = slice(off, off + p.numel())
                state[idx] = {"step": torch.tensor(float(self.step_count)),
                              "exp_avg": g.exp_avg[sl].view(p.shape).cpu(),      # straight to the host: no
                              "exp_avg_sq": g.exp_avg_sq[sl].view(p.shape).cpu()}  # second copy of the moments in HBM
                idx += 1
        pg = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay,
              "amsgrad": False, "params": list(range(idx))}
        return {"state": state, "param_groups": [pg]}

    def load_state_dict(self, sd: dict) -> None:
        idx = 0
        for g in self.groups:
            for p, off in zip(g.params, g.offsets):
                st = sd["state"].get(idx)
                if st is not None:
                    sl = slice(off, off + p.numel())
                    g.exp_avg[sl].copy_(st["exp_avg"].reshape(-1))
                    g.exp_avg_sq[sl].copy_(st["exp_avg_sq"].reshape(-1))
                    self.step_count = int(float(st["step"]))
                idx += 1
        self.sync_device_state()
