"""Data-parallel gradient exchange: bucketed all-reduce over NCCL/NVLink, overlapped with the DiT backward.

The reference gets data parallelism from accelerate -> DeepSpeed ZeRO-0 (``accelerator.backward``,
/root/reference/Continuous/train_SigLIP_stage1.py:270; train_configs/accelerate_config_4gpu.yaml:5-19): coarse
5e8-element buckets reduced after the fact.  Here the fused DiT engine knows its own backward schedule, so it
tells the reducer the moment a block's gradients are final (``Flux._on_grads_ready(prefix)``, fired from
``engine.flux_backward``); because ``optim.flatten`` lays the parameters out in module order, a block is ONE
contiguous slice of the flat gradient buffer and its all-reduce is issued immediately, on NCCL's own stream,
while the compute stream carries on with the next block's dgrad/wgrad GEMMs.  Only the last bucket (input
embedders, ~0.1 % of the bytes) and the small fp32 projector group are exposed.

One process per GPU; plumbing is ``torch.distributed`` (backend nccl on the GPUs, gloo in the CPU tests).
The sum is NOT divided here: ``FusedAdamW.step(grad_scale=1/world)`` folds the mean into the update.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .optim import FlatGroup


class GradReducer:
    def __init__(self, groups: list[FlatGroup], engine_modules=(), process_group=None,
                 bucket_cap_bytes: int = 256 << 20):
        self.groups = groups
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.bucket_cap_bytes = bucket_cap_bytes
        self.enabled = True          # False on gradient-accumulation micro-steps that do not synchronise
        # While buckets are in flight NCCL's kernels hold some SMs: the persistent GEMM grids switch to the dynamic
        # tile schedule (gh_gemm_args::dynamic_tiles, a per-launch field) from the first bucket until wait(), so CTAs that start late find the
        # tile queue drained instead of owing a full static share (CUDA tensors only; the CPU tests run on gloo).
        self.dynamic_tiles = bool(groups) and groups[0].flat_g.is_cuda
        self._dyn_on = False
        self._works = []
        self._done: dict[int, list[tuple[int, int]]] = {}
        self.log: list[tuple[str, int, int]] = []   # (prefix, lo, hi) in issue order, for tests / tracing
        managed_dtypes = {p.dtype for m in engine_modules for p in m.parameters()}
        self._managed = {id(g) for g in groups if g.dtype in managed_dtypes}
        for m in engine_modules:
            m._on_grads_ready = self.on_ready

    # ---- called from the backward schedule -------------------------------------------------------------
    def on_ready(self, prefix: str) -> None:
        """Every gradient whose parameter name starts with ``prefix`` is final: all-reduce the part of that range
        that has not been issued yet (prefixes may nest: the last block announces its sub-modules one by one so that
        only a small tail of the exchange is left when the backward ends, then the whole block)."""
        if not self.enabled or self.world == 1:
            return
        for gi, g in enumerate(self.groups):
            if id(g) not in self._managed:
                continue
            if prefix == "":
                self._reduce_rest(gi, g, prefix)
                continue
            rng = g.range_of(prefix)
            if rng is not None:
                self._reduce_rest(gi, g, prefix, rng[0], rng[1])

    def _issue(self, gi: int, g: FlatGroup, lo: int, hi: int, prefix: str) -> None:
        if hi <= lo:
            return
        self._done.setdefault(gi, []).append((lo, hi))
        if self.dynamic_tiles and not self._dyn_on:
            self._set_dynamic(True)
        cap = max(1, self.bucket_cap_bytes // g.flat_g.element_size())
        for s in range(lo, hi, cap):
            e = min(hi, s + cap)
            self.log.append((prefix, s, e))
            self._works.append(dist.all_reduce(g.flat_g[s:e], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))

    def _reduce_rest(self, gi: int, g: FlatGroup, prefix: str = "", lo: int = 0, hi: int | None = None) -> None:
        """All-reduce every slice of [lo, hi) of group ``gi`` (default: the whole group) not reduced yet."""
        hi = g.numel if hi is None else hi
        cur = lo
        gaps = []
        for a, b in sorted(self._done.get(gi, [])):
            if b <= cur or a >= hi:
                continue
            if a > cur:
                gaps.append((cur, a))
            cur = max(cur, b)
        if cur < hi:
            gaps.append((cur, hi))
        for a, b in gaps:
            self._issue(gi, g, a, b, prefix or "<rest>")

    # ---- called by the trainer after loss.backward() ---------------------------------------------------
    def issue_rest(self) -> None:
        """Issue (without waiting) the all-reduce of whatever the backward schedule did not announce: the
        autograd-managed groups such as the fp32 projectors / adapter.  Call right after ``loss.backward()``."""
        if self.enabled and self.world > 1:
            for gi, g in enumerate(self.groups):
                self._reduce_rest(gi, g)

    def wait(self) -> None:
        """Make the current stream wait for every bucket issued so far."""
        for w in self._works:
            w.wait()
        self._works.clear()
        self._done.clear()
        if self._dyn_on:
            self._set_dynamic(False)

    def _set_dynamic(self, on: bool) -> None:
        from . import kernels as K
        K.DYNAMIC_TILES = bool(on)      # launch policy of the host wrapper; the library itself holds no such state
        self._dyn_on = on

    def finish(self) -> None:
        """``issue_rest()`` + ``wait()``.  The data-parallel loops call the two halves separately: the rest is issued
        right after backward, the wait is deferred into the next step (``before_trainable``), so the tail of the
        exchange hides under the frozen AE / tower forward."""
        self.issue_rest()
        self.wait()

    @property
    def grad_scale(self) -> float:
        return 1.0 / self.world


def broadcast_parameters(groups: list[FlatGroup], src: int = 0, process_group=None) -> None:
    """Make every rank start from rank ``src``'s parameters (flat buffers -> one broadcast per dtype group)."""
    if dist.is_initialized() and dist.get_world_size(process_group) > 1:
        for g in groups:
            dist.broadcast(g.flat_p, src=src, group=process_group)
