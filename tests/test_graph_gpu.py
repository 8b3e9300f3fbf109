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


def _small_training_setup(lr=1e-3):
    from genhancer_b200 import optim
    from test_step_gpu import build_step
    fx = load_golden("step_small.pt")
    step, wrap, dit = build_step(fx["tower_cfg"], fx["flux_cfg"], fx["ae_cfg"], fx["key_shapes"], fx["seed"],
                                 fx["clip_dim"], fx["t5_dim"])
    groups = optim.flatten(list(dit.named_parameters()) + [(f"clip_vis.{n}", p) for n, p in wrap.named_parameters()])
    opt = optim.FusedAdamW(groups, lr=lr, engine_managed=[dit])
    draws = dict(ae_noise=fx["ae_noise"].cuda(), t=fx["t"].cuda(), x_0=fx["x_0"].cuda())
    g = torch.Generator(device="cuda").manual_seed(7)
    imgs = [torch.rand(fx["img"].shape, device="cuda", generator=g) for _ in range(6)]
    fn = lambda x, before_trainable=None: step(x, before_trainable=before_trainable, **draws)   # noqa: E731
    return fn, groups, opt, imgs


@pytest.mark.parametrize("ga", [1, 2])
def test_pipelined_train_step_equals_the_sequential_loop(ga):
    """PipelinedTrainStep (the update of step n captured on a forked branch at the head of replay n+1, reading its step
    count from device memory) trains exactly like ``loss.backward(); opt.step(); opt.zero_grad()``: same losses, same
    weights after 4 optimizer steps (+ a flush), with and without gradient accumulation."""
    from genhancer_b200.graph import PipelinedTrainStep
    n_opt = 4
    # --- the sequential eager loop (train_SigLIP_stage1.py:238-275) ---
    fn, groups, opt, imgs = _small_training_setup()
    p0 = [g.flat_p.clone() for g in groups]
    losses_e = []
    for i in range(n_opt * ga):
        loss = fn(imgs[i % len(imgs)])
        (loss / ga).backward()
        losses_e.append(loss.item())
        del loss
        if (i + 1) % ga == 0:
            opt.step()
            opt.zero_grad()
    torch.cuda.synchronize()
    p_eager = [g.flat_p.clone() for g in groups]
    m_eager = [g.exp_avg.clone() for g in groups]
    assert opt.step_count == n_opt
    # --- the pipelined graph, same initial weights ---
    fn, groups, opt, imgs = _small_training_setup()
    for g, p in zip(groups, p0):
        assert torch.equal(g.flat_p, p)
    pipe = PipelinedTrainStep(fn, (imgs[0],), opt, reducer=None, grad_accum=ga)
    for g, p in zip(groups, p0):
        assert torch.equal(g.flat_p, p), "capture / warm-up must not touch the weights"
    assert opt.step_count == 0 and len(pipe.graphs) == ga
    losses_g = []
    for i in range(n_opt * ga):
        losses_g.append(pipe(imgs[i % len(imgs)]).item())
    assert opt.step_count == n_opt - 1 and pipe.host_pending and pipe.optimizer_steps == n_opt
    pipe.flush()
    torch.cuda.synchronize()
    assert opt.step_count == n_opt and not pipe.host_pending
    assert int(opt.dev_state[0]) == n_opt and int(opt.dev_state[1]) == 0
    for a, b in zip(losses_g, losses_e):
        assert abs(a - b) <= 2e-3 * abs(b), (losses_g, losses_e)
    for g, pe, me in zip(groups, p_eager, m_eager):
        assert cosine(g.flat_p, pe) >= 0.999999 and cosine(g.exp_avg, me) >= 0.999
        assert (g.flat_p.float() - pe.float()).abs().max().item() <= 3e-2 * 1e-3 * n_opt + 1e-2 * pe.float().abs().max().item()
    # a flush followed by more replays: the first of them must NOT apply a stale update
    p_before = [g.flat_p.clone() for g in groups]
    pipe(imgs[0])
    torch.cuda.synchronize()
    for g, p in zip(groups, p_before):
        assert torch.equal(g.flat_p, p), "the replay after a flush applied an update although nothing was pending"
    pipe.reset()
