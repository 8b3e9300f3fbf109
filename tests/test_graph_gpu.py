"""The whole micro-step (forward + backward) as one CUDA graph gives the eager step's loss and gradients."""
import pytest
import torch

from conftest import cosine, load_golden

pytestmark = pytest.mark.gpu


def test_graphed_micro_step_matches_eager():
    from genhancer_b200 import optim
    from genhancer_b200.graph import GraphedMicroStep
    from test_step_gpu import build_step
    fx = load_golden("step_small.pt")
    step, wrap, dit = build_step(fx["tower_cfg"], fx["flux_cfg"], fx["ae_cfg"], fx["key_shapes"], fx["seed"],
                                 fx["clip_dim"], fx["t5_dim"])
    groups = optim.flatten(list(dit.named_parameters()) + [(f"clip_vis.{n}", p) for n, p in wrap.named_parameters()])
    opt = optim.FusedAdamW(groups, lr=1e-4, engine_managed=[dit])
    draws = dict(ae_noise=fx["ae_noise"].cuda(), t=fx["t"].cuda(), x_0=fx["x_0"].cuda())
    img = fx["img"].cuda()
    fn = lambda x: step(x, **draws)     # fixed draws: eager and graph must agree exactly on what they compute
    opt.zero_grad()
    loss_e = fn(img)
    loss_e.backward()
    le = loss_e.item()
    del loss_e                           # no eager autograd graph (AccumulateGrad nodes of the legacy stream) may outlive this
    g_e = [g.flat_g.clone() for g in groups]
    opt.zero_grad()
    gs = GraphedMicroStep(fn, img, prepare=opt.zero_grad)
    assert gs.launches_per_replay > 50
    for _ in range(2):                   # replays are repeatable
        opt.zero_grad()
        loss_g = gs(img)
        torch.cuda.synchronize()
        assert abs(loss_g.item() - le) <= 1e-6 * abs(le)
        for a, b in zip(groups, g_e):
            assert cosine(a.flat_g, b) >= 0.99999     # (fp32 atomics in the bias / LN-parameter reductions: not bit-stable)
    # a different input through the same graph changes the result, and the optimizer runs outside the graph
    opt.zero_grad()
    l2 = gs(torch.rand_like(img))
    opt.step()
    torch.cuda.synchronize()
    assert l2.item() != le and torch.isfinite(l2)


def test_graphed_step_draws_fresh_noise_each_replay():
    """Without fixed draws the step's three torch.randn calls sit inside the graph: torch's graph-safe Philox
    generator advances per replay, so two replays on the same image see different t / x_0."""
    from genhancer_b200 import optim
    from genhancer_b200.graph import GraphedMicroStep
    from test_step_gpu import build_step
    fx = load_golden("step_small.pt")
    step, wrap, dit = build_step(fx["tower_cfg"], fx["flux_cfg"], fx["ae_cfg"], fx["key_shapes"], fx["seed"],
                                 fx["clip_dim"], fx["t5_dim"])
    groups = optim.flatten(list(dit.named_parameters()) + [(f"clip_vis.{n}", p) for n, p in wrap.named_parameters()])
    opt = optim.FusedAdamW(groups, lr=1e-4, engine_managed=[dit])
    img = fx["img"].cuda()
    gs = GraphedMicroStep(step, img, prepare=opt.zero_grad)
    opt.zero_grad()
    a = gs(img).item()
    opt.zero_grad()
    b = gs(img).item()
    assert a != b and a > 0 and b > 0
