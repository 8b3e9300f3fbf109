"""Host logic of the data-parallel path on CPU: flat parameter layout, per-block bucket ranges, and a world_size-2
gloo run in which two ranks with different gradients end up with identical, averaged flat gradients."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from genhancer_b200 import optim
from genhancer_b200.parallel import GradReducer, broadcast_parameters


class _Blk(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.lin = torch.nn.Linear(24, 40)
        self.scale = torch.nn.Parameter(torch.ones(7))


class _Net(torch.nn.Module):
    """Same registration-order shape as Flux: input layers, blocks, final layer."""

    def __init__(self):
        super().__init__()
        self.img_in = torch.nn.Linear(8, 24)
        self.double_blocks = torch.nn.ModuleList([_Blk(), _Blk()])
        self.single_blocks = torch.nn.ModuleList([_Blk(), _Blk(), _Blk()])
        self.final_layer = torch.nn.Linear(24, 8)
        self._on_grads_ready = None
        self._grad_overwrite = False


def test_flatten_layout_views_and_ranges():
    net = _Net()
    ref = {n: p.detach().clone() for n, p in net.named_parameters()}
    extra = torch.nn.Parameter(torch.randn(5, dtype=torch.float64))  # a second dtype group
    groups = optim.flatten(list(net.named_parameters()) + [("project.w", extra)])
    assert len(groups) == 2
    g = next(x for x in groups if x.dtype == torch.float32)
    for n, p in net.named_parameters():
        assert torch.equal(p.detach(), ref[n])                      # values preserved
        assert p.grad is not None and p.grad.shape == p.shape       # permanent grad views
        off = g.offsets[g.names.index(n)]
        assert off % 32 == 0                                         # 128-byte alignment (fp32)
        assert p.data_ptr() == g.flat_p.data_ptr() + 4 * off
        assert p.grad.data_ptr() == g.flat_g.data_ptr() + 4 * off
    lo, hi = g.range_of("single_blocks.1.")
    mine = [(g.offsets[i], g.offsets[i] + g.params[i].numel()) for i, n in enumerate(g.names)
            if n.startswith("single_blocks.1.")]
    assert len(mine) == 3 and lo == min(a for a, _ in mine) and hi == max(b for _, b in mine)
    others = [g.offsets[i] for i, n in enumerate(g.names) if not n.startswith("single_blocks.1.")]
    assert all(o < lo or o >= hi for o in others)                # a block is one contiguous slice
    assert g.range_of("nope.") is None
    # writing through the flat buffer is visible in the module and vice versa
    g.flat_p.zero_()
    assert all(float(p.detach().abs().sum()) == 0 for p in net.parameters())


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)  # different init per rank: broadcast must fix it
        net = _Net()
        groups = optim.flatten(list(net.named_parameters()))
        broadcast_parameters(groups)
        red = GradReducer(groups, engine_modules=[net], bucket_cap_bytes=1024)
        g = groups[0]
        g.flat_g.copy_(torch.arange(g.numel, dtype=torch.float32) * (rank + 1))
        # the engine's backward schedule announces blocks in reverse execution order, then "" for the rest
        # (the last block announces its sub-modules first, then itself: nested prefixes must not reduce twice)
        for prefix in ["final_layer.", "single_blocks.2.", "single_blocks.1.", "single_blocks.0.", "double_blocks.1.",
                       "double_blocks.0.lin.", "double_blocks.0.scale", "double_blocks.0.", ""]:
            net._on_grads_ready(prefix)
        red.finish()
        expect = torch.arange(g.numel, dtype=torch.float32) * sum(r + 1 for r in range(world))
        ok_sum = torch.equal(g.flat_g, expect)
        # every element reduced exactly once, first bucket issued is the last block of the network
        covered = torch.zeros(g.numel)
        for _, lo, hi in red.log:
            covered[lo:hi] += 1
        res = dict(ok_sum=bool(ok_sum), once=bool((covered == 1).all()), first=red.log[0][0],
                         n_buckets=len(red.log), p0=float(g.flat_p.double().sum()), scale=red.grad_scale)
        # second round through the two halves the data-parallel loops use (issue right after backward, wait deferred
        # into the next step); on CPU tensors the dynamic tile schedule of the CUDA GEMMs is never touched
        res["dyn_off_on_cpu"] = (red.dynamic_tiles is False) and (red._dyn_on is False)
        g.flat_g.copy_(torch.ones(g.numel) * (rank + 1))
        net._on_grads_ready("single_blocks.0.")
        red.issue_rest()
        red.wait()
        res["two_halves"] = bool(torch.equal(g.flat_g, torch.ones(g.numel) * sum(r + 1 for r in range(world))))
        # accumulation micro-step: nothing is exchanged
        red.enabled = False
        before = g.flat_g.clone()
        net._on_grads_ready("final_layer.")
        red.finish()
        res["no_sync_untouched"] = bool(torch.equal(before, g.flat_g))
        out[rank] = res
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_world2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    for r in (r0, r1):
        assert r["ok_sum"] and r["once"] and r["no_sync_untouched"] and r["two_halves"] and r["dyn_off_on_cpu"]
        assert r["first"] == "final_layer." and r["n_buckets"] > 7   # 1 KiB cap splits the blocks into several buckets
        assert r["scale"] == 0.5
    assert r0["p0"] == r1["p0"]                                       # broadcast made the replicas identical


def test_gradsink_side_stream_is_inert_without_cuda():
    """The wgrad side stream of the DiT backward (flux/engine.py) only exists for CUDA tensors: on CPU parameters
    ``side()`` is a no-op context and ``join()`` just releases the kept operands."""
    from genhancer_b200.flux.engine import GradSink
    params = {"lin.weight": torch.nn.Parameter(torch.zeros(4, 4)), "lin.bias": torch.nn.Parameter(torch.zeros(4))}
    sink = GradSink(params, accumulate=False)
    a = torch.ones(2, 4)
    with sink.side(a):
        pass
    assert sink._forked is False
    sink.join()
    assert sink._keep == []
