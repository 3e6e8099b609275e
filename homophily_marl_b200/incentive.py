"""Incentive bookkeeping of the homophily learner on the device.

Restates ``src/learners/homophily_learner.py:98-115`` (and the ``sign(receive_value)`` feature of
``src/controllers/homophily_controller.py:154-164``) as one kernel behind ``ssd_incentive``:

    give[i]   = #{j != i : a[i][j] != 0}          recv+/-[j] = #{i != j : a[i][j] == 1 / 2}
    rewards_for_env = (r + (recv+ - recv-) * ratio * incentive) / T
    rewards_for_inc = (r - give * cost * incentive) / T

fp32, same operation order as the torch expressions.  ``recip=True`` multiplies by ``1/T`` (what torch's CUDA kernels do
for a scalar divisor), ``recip=False`` divides (torch on the CPU).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _capi


def incentive_rewards(actions_inc: torch.Tensor, rewards: torch.Tensor, incentive: float, cost: float, ratio: float,
                      max_seq_length: int, recip: bool = True):
    """actions_inc: int64 [..., n, n, 1] (or [..., n, n]); rewards: float32 [..., n].
    Returns (rewards_for_env, rewards_for_inc, sign_of_receive_value), each float32 [..., n]."""
    if not (actions_inc.is_cuda and rewards.is_cuda):
        raise RuntimeError("incentive_rewards needs CUDA tensors; there is no CPU fallback")
    n = rewards.shape[-1]
    a = actions_inc.reshape(-1, n, n).to(torch.int64).contiguous()
    r = rewards.reshape(-1, n).to(torch.float32).contiguous()
    if a.shape[0] != r.shape[0]:
        raise ValueError("actions_inc and rewards disagree on the number of rows")
    outs = [torch.empty_like(r) for _ in range(3)]
    lib = _capi.load()
    with torch.cuda.device(r.device):
        _capi.check(lib.ssd_incentive(a.data_ptr(), r.data_ptr(), a.shape[0], n, float(incentive), float(cost), float(ratio),
                                      -int(max_seq_length) if recip else int(max_seq_length),
                                      outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(),
                                      C.c_void_p(torch.cuda.current_stream(r.device).cuda_stream)))
    return tuple(o.view(rewards.shape) for o in outs)
