"""CPU suite: host-side logic of the policy-side rows -- the epsilon schedule mirrors the reference's
DecayThenFlatSchedule (src/components/epsilon_schedules.py:5-26), the selector fails loudly without CUDA, and the registries
that INTEGRATION.md section 6 extends exist with the documented names."""
import types

import pytest

from baseline import refloop

torch = pytest.importorskip("torch")


def test_schedule_equals_the_reference_schedule():
    from homophily_marl_b200.selectors import DecayThenFlatSchedule
    ours = DecayThenFlatSchedule(1.0, 0.05, 50000, decay="linear")
    for T in (0, 1, 100, 25000, 49999, 50000, 50001, 10 ** 7):
        assert ours.eval(T) == max(0.05, 1.0 - (1.0 - 0.05) / 50000 * T)
    if refloop.available():
        refloop.import_reference()
        from components.epsilon_schedules import DecayThenFlatSchedule as Ref
        ref = Ref(1.0, 0.05, 50000, decay="linear")
        for T in (0, 7, 12345, 50000, 99999):
            assert ours.eval(T) == ref.eval(T)
    with pytest.raises(ValueError):
        DecayThenFlatSchedule(1.0, 0.05, 50000, decay="exp")


def test_selector_and_front_end_have_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from homophily_marl_b200.frontend import ObsFrontEnd
    from homophily_marl_b200.selectors import DeviceEpsilonGreedySelector
    sel = DeviceEpsilonGreedySelector(types.SimpleNamespace(epsilon_start=1.0, epsilon_finish=0.05, epsilon_anneal_time=100, epsilon_zero=None))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sel.select_action(torch.zeros(2, 3, 9), torch.ones(2, 3, 9), 0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ObsFrontEnd(torch.zeros(6, 3, 3, 3), torch.zeros(6), torch.zeros(32, 6 * 13 * 13), torch.zeros(32), view=7)


def test_registries_named_in_integration_md():
    from homophily_marl_b200 import learner, selectors
    assert set(selectors.REGISTRY) == {"epsilon_greedy_b200"} and set(learner.REGISTRY) == {"homophily_learner_b200"}
    if refloop.available() and not torch.cuda.is_available():
        # register_b200 extends the reference's live registries without touching its files (the env entries need CUDA to be USED,
        # not to be registered)
        refloop.register_b200()
        import components.action_selectors as a
        import envs
        import learners
        import runners
        assert "epsilon_greedy_b200" in a.REGISTRY and "homophily_learner_b200" in learners.REGISTRY and "batched" in runners.REGISTRY
        assert envs.REGISTRY["cleanup"].keywords["env"].__module__ == "homophily_marl_b200.pymarl_env"
        envs.REGISTRY.update(envs._reference_registry)                     # leave the reference's own constructors in place
        assert envs.REGISTRY["cleanup"].keywords["env"].__module__ == "envs.ssd.cleanup"
