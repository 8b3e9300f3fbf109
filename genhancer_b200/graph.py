"""Whole-step CUDA graph: forward + backward of one training micro-step captured once, replayed per step.

The step is ~780 kernel launches issued from Python (ctypes + autograd bookkeeping: ~60 ms of host time per step
against ~100 ms of GPU time); where consecutive kernels are short the GPU waits for the host.  Shapes, buffers and
the schedule are static, the three RNG draws use torch's graph-safe Philox generator, and nothing in the step
synchronises with the host, so the whole micro-step can be a graph ("CUDA streams and graphs instead of a tracing
compiler").  The optimizer update and ``zero_grad`` stay outside (the AdamW kernel takes the step count by value).

The reference has no analogue (eager PyTorch through accelerate, train_SigLIP_stage1.py:238-275).
"""
from __future__ import annotations

import torch


class GraphedMicroStep:
    """``loss = step_fn(img); loss.backward()`` as one CUDA graph over a static input buffer.

    ``prepare()`` (optional) is run before capture and must leave the gradients in the state every replay starts
    from (``FusedAdamW.zero_grad()``: the DiT overwrites its flat gradient buffer on the first backward after it).
    The caller runs ``prepare`` / ``zero_grad`` and the optimizer itself around each call, exactly as in eager mode.

    ``after_backward()`` (optional) is captured right after the backward: data parallelism passes
    ``GradReducer.finish`` -- the per-block NCCL all-reduces the backward schedule issues on NCCL's stream become a
    forked branch of the graph and ``finish`` joins it, so a replay carries forward, backward AND the overlapped
    gradient exchange with no Python between them (the eager data-parallel step was host-paced: ~105 ms on 2 GPUs
    without any exchange against 99.5 ms graphed on one).
    """

    def __init__(self, step_fn, example_img: torch.Tensor, prepare=None, warmup: int = 2, after_backward=None):
        if not example_img.is_cuda:
            raise RuntimeError("GraphedMicroStep needs CUDA tensors (there is no CPU fallback)")
        self.static_img = example_img.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # lazily-built state (operand caches, id tensors, TMEM/TMA setup) before capture
            for _ in range(warmup):
                if prepare is not None:
                    prepare()
                step_fn(self.static_img).backward()
                if after_backward is not None:
                    after_backward()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if prepare is not None:
            prepare()
        from . import kernels as K
        n0 = K.LAUNCHES
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: NCCL's watchdog thread may query its events while this thread captures
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.static_loss = step_fn(self.static_img)
            self.static_loss.backward()
            if after_backward is not None:
                after_backward()
        self.launches_per_replay = K.LAUNCHES - n0
        self.replays = 0

    def __call__(self, img: torch.Tensor) -> torch.Tensor:
        """Copies ``img`` (device or pinned host) into the static buffer, replays, returns the (static) loss tensor."""
        self.static_img.copy_(img, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        return self.static_loss
