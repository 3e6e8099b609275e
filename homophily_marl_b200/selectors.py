"""Epsilon-greedy action selection on the device.

Mirror of ``EpsilonGreedyActionSelector`` (src/components/action_selectors.py:44-68) with the same constructor and
``select_action(agent_inputs, avail_actions, t_env, test_mode)`` signature, so that ``HomophilyMAC`` can use it unchanged
(``action_selectors.REGISTRY['epsilon_greedy_b200']``, ``args.action_selector='epsilon_greedy_b200'``).  The reference spends
seven torch launches per call (clone, masked fill, rand_like, compare, multinomial, max, blend); here it is one kernel
behind ``ssd_select_actions``.  Randomness: Philox4x32-10 keyed (seed; row, call counter) or injected uniforms.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _capi


class DecayThenFlatSchedule:
    """Linear decay to ``finish`` over ``time_length`` steps, then flat (epsilon_schedules.py:5-26, decay='linear')."""

    def __init__(self, start, finish, time_length, decay="linear"):
        if decay != "linear":
            raise ValueError("only the linear schedule is used by the reference's selectors")
        self.start, self.finish, self.time_length = start, finish, time_length
        self.delta = (self.start - self.finish) / self.time_length

    def eval(self, T):
        return max(self.finish, self.start - self.delta * T)


def select_actions(q: torch.Tensor, avail: torch.Tensor | None, epsilon: float, u_pick=None, u_act=None, seed=0, counter=0):
    """q [..., A] float32 CUDA; avail [..., A] (any dtype, 0 = unavailable) or None.  Returns int64 [...]."""
    if not q.is_cuda:
        raise RuntimeError("select_actions needs CUDA tensors; there is no CPU fallback")
    A = q.shape[-1]
    qf = q.detach().reshape(-1, A).to(torch.float32).contiguous()
    av = None if avail is None else (avail.reshape(-1, A) != 0).to(torch.int32).contiguous()
    rows = qf.shape[0]
    out = torch.empty(rows, dtype=torch.int64, device=q.device)
    up = None if u_pick is None else u_pick.reshape(-1).to(device=q.device, dtype=torch.float32).contiguous()
    ua = None if u_act is None else u_act.reshape(-1).to(device=q.device, dtype=torch.float32).contiguous()
    lib = _capi.load()
    with torch.cuda.device(q.device):
        _capi.check(lib.ssd_select_actions(qf.data_ptr(), None if av is None else av.data_ptr(), rows, A, float(epsilon),
                                           None if up is None else up.data_ptr(), None if ua is None else ua.data_ptr(),
                                           int(seed) & (2 ** 64 - 1), int(counter) & (2 ** 64 - 1), out.data_ptr(),
                                           C.c_void_p(torch.cuda.current_stream(q.device).cuda_stream)))
    return out.view(q.shape[:-1])


class DeviceEpsilonGreedySelector:
    def __init__(self, args):
        self.args = args
        self.schedule = DecayThenFlatSchedule(args.epsilon_start, args.epsilon_finish, args.epsilon_anneal_time, decay="linear")
        self.epsilon = self.schedule.eval(0)
        self.seed = int(getattr(args, "seed", 0) or 0)
        self.calls = 0

    def select_action(self, agent_inputs, avail_actions, t_env, test_mode=False, u_pick=None, u_act=None):
        self.epsilon = self.schedule.eval(t_env)                                   # action_selectors.py:47-49
        if getattr(self.args, "epsilon_zero", None) is not None and t_env > self.args.epsilon_zero:
            self.epsilon = 0.0
        if test_mode:
            self.epsilon = 0.0                                                     # 51-53: greedy only
        self.calls += 1
        return select_actions(agent_inputs, avail_actions, self.epsilon, u_pick=u_pick, u_act=u_act, seed=self.seed,
                              counter=self.calls)


REGISTRY = {"epsilon_greedy_b200": DeviceEpsilonGreedySelector}
