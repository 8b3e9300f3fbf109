"""Whole-step CUDA graph: forward + backward of one training micro-step captured once, replayed per step.

The step is ~780 kernel launches issued from Python (ctypes + autograd bookkeeping: ~60 ms of host time per step
against ~100 ms of GPU time); where consecutive kernels are short the GPU waits for the host.  Shapes, buffers and
the schedule are static, the three RNG draws use torch's graph-safe Philox generator, and nothing in the step
synchronises with the host, so the whole micro-step can be a graph ("CUDA streams and graphs instead of a tracing
compiler").  The optimizer update and ``zero_grad`` stay outside (the AdamW kernel takes the step count by value).

The reference has no analogue (eager PyTorch through accelerate, train_SigLIP_stage1.py:238-275).
"""
from __future__ import annotations

import os

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


class PipelinedTrainStep:
    """The WHOLE training step -- optimizer update included -- as one CUDA graph per micro-step variant.

    A replay of the graph is

        opt stream :  [update of the PREVIOUS optimizer step: step tick, sumsq, clip + AdamW, zero_grad]
        main stream:  frozen AE encode (side stream) | frozen tower forward  --join opt stream-->  projectors / adapter
                      -> DiT forward -> loss -> backward (weight gradients on the wgrad side stream, NCCL buckets on
                      NCCL's stream) -> [join the exchange] -> mark "update pending"

    i.e. the ~3.5 ms HBM-bound ``sumsq`` + AdamW of step n run on a forked branch UNDER the tensor-bound frozen forward
    of step n+1 instead of alone between two replays (``before_trainable`` of the step objects is the join point: the
    first kernel that reads a trainable parameter waits for the update, so the arithmetic is that of the sequential
    loop ``backward; clip; step; zero_grad`` of train_SigLIP_stage1.py:270-275).  The update reads its step count and
    a "pending" flag from device memory (``FusedAdamW.step_captured``), so the very first replay -- and the one after
    an eager ``flush()`` (checkpoint, end of training) -- applies nothing.

    Gradient accumulation: ``grad_accum`` micro-steps per optimizer step are up to three captured variants
    (first: update + overwrite gradients; middle: accumulate; last: accumulate + exchange + mark pending) sharing one
    memory pool and replayed in a fixed cycle; the loss of every micro-step is divided by ``grad_accum`` before
    backward, as ``accelerator.backward`` does.

    ``step_fn(*inputs, before_trainable=cb) -> loss`` is ``Stage1ImageStep`` / ``VideoStep`` (wrapped by the caller when
    the inputs are not a flat tuple of tensors)."""

    def __init__(self, step_fn, example_inputs, opt, reducer=None, grad_accum: int = 1, warmup: int = 2):
        ex = tuple(example_inputs)
        if not ex or not all(t.is_cuda for t in ex):
            raise RuntimeError("PipelinedTrainStep needs CUDA tensors (there is no CPU fallback)")
        from . import kernels as K
        self.step_fn, self.opt, self.reducer = step_fn, opt, reducer
        self.ga = max(1, int(grad_accum))
        self.grad_scale = reducer.grad_scale if reducer is not None else 1.0
        self.static_in = tuple(t.clone() for t in ex)
        self.shapes = tuple((tuple(t.shape), t.dtype) for t in ex)
        dev = ex[0].device
        self.opt_stream = torch.cuda.Stream(device=dev)
        self.overlap_update = os.environ.get("GH_OPT_OVERLAP", "1") != "0"
        self.micro = 0                 # position inside the accumulation cycle
        self.host_pending = False      # a finished backward whose update has not been applied yet
        self.replays = 0
        self.launches_per_replay = {}
        self.graphs, self.static_loss = {}, {}
        variants = list(dict.fromkeys((m == 0, m == self.ga - 1) for m in range(self.ga)))   # in cycle order
        pool = torch.cuda.graph_pool_handle() if len(variants) > 1 else None
        opt.mark_pending(False)
        for v in variants:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):       # lazily-built state before capture; updates are no-ops (nothing pending)
                for _ in range(warmup):
                    self._micro_step(*v)
                    opt.mark_pending(False)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            n0 = K.LAUNCHES
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool, capture_error_mode="thread_local"):
                self.static_loss[v] = self._micro_step(*v)
            self.graphs[v] = g
            self.launches_per_replay[v] = K.LAUNCHES - n0
            opt.mark_pending(False)
        torch.cuda.synchronize(dev)
        opt.zero_grad()

    # one micro-step, as captured ------------------------------------------------------------------------------
    def _micro_step(self, first: bool, last: bool):
        opt, red = self.opt, self.reducer
        cur = torch.cuda.current_stream()
        join = None
        if first and not self.overlap_update:       # A/B switch: the update alone at the head of the replay
            opt.step_captured(self.grad_scale)
            opt.zero_grad()
        elif first:
            self.opt_stream.wait_stream(cur)
            with torch.cuda.stream(self.opt_stream):
                opt.step_captured(self.grad_scale)
                opt.zero_grad()                     # engine-managed groups: "overwrite on the next backward"
            join = lambda: torch.cuda.current_stream().wait_stream(self.opt_stream)   # noqa: E731
        if red is not None:
            red.enabled = last
        done = {"j": False}

        def before_trainable():
            if join is not None and not done["j"]:
                join()
                done["j"] = True

        loss = self.step_fn(*self.static_in, before_trainable=before_trainable)
        before_trainable()                          # (a step object that never called it)
        (loss / self.ga if self.ga > 1 else loss).backward()
        if last:
            if red is not None:
                red.finish()
            opt.mark_pending(True)
        return loss.detach()

    # public ---------------------------------------------------------------------------------------------------
    def matches(self, *inputs) -> bool:
        return len(inputs) == len(self.shapes) and all((tuple(t.shape), t.dtype) == s for t, s in zip(inputs, self.shapes))

    def __call__(self, *inputs) -> torch.Tensor:
        """One micro-step: copies the inputs (device or pinned host) into the static buffers and replays the variant
        of this position in the accumulation cycle.  Returns the (static) loss tensor of that variant."""
        for dst, src in zip(self.static_in, inputs):
            dst.copy_(src, non_blocking=True)
        v = (self.micro == 0, self.micro == self.ga - 1)
        self.graphs[v].replay()
        self.replays += 1
        if v[0] and self.host_pending:              # the replay applied the pending update on the device
            self.opt.step_count += 1
            self.host_pending = False
        if v[1]:
            self.host_pending = True
        self.micro = (self.micro + 1) % self.ga
        return self.static_loss[v]

    @property
    def optimizer_steps(self) -> int:
        """Optimizer steps whose backward has been issued (applied or pending)."""
        return self.opt.step_count + int(self.host_pending)

    def flush(self) -> None:
        """Apply the pending update now (eagerly): before a checkpoint, an eager fallback step, or at the end."""
        if self.host_pending:
            if self.micro != 0:
                raise RuntimeError("flush() in the middle of a gradient-accumulation cycle")
            self.opt.step(self.grad_scale)
            self.opt.zero_grad()
            self.opt.sync_device_state()
            self.host_pending = False

    def reset(self) -> None:
        """Drop the graphs (a graph that holds NCCL kernels must go before the communicator does)."""
        for g in self.graphs.values():
            g.reset()
        self.graphs.clear()
