"""GPU suite: epsilon-greedy selection on the device equals the reference's
``EpsilonGreedyActionSelector.select_action`` (src/components/action_selectors.py:44-68) when both consume the same
injected uniforms -- ``th.rand_like`` (line 62) returns ``u_pick`` and ``th.multinomial`` over the 0/1 mask (line 66) picks
the floor(u_act * #available)-th available action, which is the same uniform choice among available actions."""
import types

import numpy as np
import pytest

from baseline import refloop

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _args(**kw):
    a = dict(epsilon_start=1.0, epsilon_finish=0.05, epsilon_anneal_time=50000, epsilon_zero=None, seed=7)
    a.update(kw)
    return types.SimpleNamespace(**a)


def _kth_available(avail2d, u):
    n_av = (avail2d != 0).sum(-1)
    k = torch.clamp((u * n_av.float()).long(), max=n_av - 1)
    order = torch.cumsum((avail2d != 0).long(), -1) - 1                       # rank of each available action
    hit = (order == k.unsqueeze(-1)) & (avail2d != 0)
    return hit.float().argmax(-1, keepdim=True)


@pytest.mark.skipif(not refloop.available(), reason="reference sources absent (baseline/_ref)")
@pytest.mark.parametrize("shape,A,t_env", [((64, 5), 9, 0), ((33, 10), 8, 30000), ((16, 3, 3), 3, 49000), ((7, 5), 9, 10 ** 7)])
def test_device_selector_equals_reference_selector_under_injected_uniforms(shape, A, t_env):
    refloop.import_reference()
    import components.action_selectors as ref_sel
    from homophily_marl_b200.selectors import DeviceEpsilonGreedySelector
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(A * 100 + len(shape))
    q = torch.randn(*shape, A, generator=g)
    q[..., 1] = q[..., 0]                                                       # ties: the first maximum must win
    avail = (torch.rand(*shape, A, generator=g) < 0.7).int()
    avail[..., 4 % A] = 1                                                       # at least one action available
    if A == 3:
        avail = torch.ones(*shape, A)                                           # the incentive head's float mask of ones
    u_pick, u_act = torch.rand(*shape, generator=g), torch.rand(*shape, generator=g)
    ref = ref_sel.EpsilonGreedyActionSelector(_args())
    saved = (torch.rand_like, torch.multinomial)
    try:
        torch.rand_like = lambda x, *a, **k: u_pick.to(x.device)
        torch.multinomial = lambda w, n, replacement=False: _kth_available(w, u_act.reshape(-1).to(w.device))
        want = ref.select_action(q.clone(), avail.clone(), t_env, test_mode=False)
        want_test = ref.select_action(q.clone(), avail.clone(), t_env, test_mode=True)
    finally:
        torch.rand_like, torch.multinomial = saved
    sel = DeviceEpsilonGreedySelector(_args())
    got = sel.select_action(q.to(dev), avail.to(dev), t_env, test_mode=False, u_pick=u_pick, u_act=u_act)
    assert sel.epsilon == ref.schedule.eval(t_env) and got.dtype == torch.int64 and got.shape == tuple(shape)
    assert torch.equal(got.cpu(), want)
    got_test = sel.select_action(q.to(dev), avail.to(dev), t_env, test_mode=True, u_pick=u_pick, u_act=u_act)
    assert sel.epsilon == 0.0 and torch.equal(got_test.cpu(), want_test)
    if t_env == 0:
        assert (got.cpu() != want_test).any()                                  # epsilon = 1: exploration really happened


def test_philox_mode_respects_the_mask_and_epsilon_statistics():
    from homophily_marl_b200.selectors import select_actions
    dev = torch.device("cuda:0")
    R, A = 200000, 9
    q = torch.randn(R, A, device=dev)
    avail = torch.ones(R, A, dtype=torch.int32, device=dev)
    avail[:, 5:8] = 0                                                           # yaml mask: no rotations, no fire
    greedy = select_actions(q, avail, 0.0)
    assert torch.equal(greedy, q.masked_fill(avail == 0, -float("inf")).argmax(-1))
    a = select_actions(q, avail, 0.3, seed=11, counter=1)
    b = select_actions(q, avail, 0.3, seed=11, counter=1)
    c = select_actions(q, avail, 0.3, seed=11, counter=2)
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert bool((avail.gather(1, a.unsqueeze(1)) == 1).all())
    # P(action != greedy) = eps * (1 - 1/6) = 0.25
    frac = (a != greedy).float().mean().item()
    assert abs(frac - 0.25) < 0.01
    rnd = select_actions(q, avail, 1.0, seed=3, counter=9)
    hist = torch.bincount(rnd, minlength=A).float() / R
    assert torch.allclose(hist[[0, 1, 2, 3, 4, 8]], torch.full((6,), 1 / 6, device=dev), atol=0.01) and hist[5:8].sum() == 0
