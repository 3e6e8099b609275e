"""Fused observation front end: u8 obs planes -> conv3x3(3->6)+LeakyReLU -> Linear(->32)+LeakyReLU.

Forward-only replacement, for rollouts, of ``HomophilyAgent.rgb_preprocess`` (src/modules/agents/homophily_agent.py:19-27,
213-214) consuming the env's u8 observation buffer directly (``ssd_frontend_forward``: conv on the CUDA cores, the Linear on
tcgen05 tensor cores with a tf32 hi/lo split and fp32 accumulation in TMEM).  The learner keeps the reference's autograd
module; ``attach_to_mac`` only reroutes the no-grad rollout calls.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _capi


class ObsFrontEnd:
    def __init__(self, conv_w, conv_b, fc_w, fc_b, view, negative_slope=0.01, device="cuda:0"):
        if not torch.cuda.is_available():
            raise RuntimeError("ObsFrontEnd needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device)
        self.view, self.N, self.P = int(view), 2 * int(view) + 1, 2 * int(view) - 1
        arrs = [np.ascontiguousarray(torch.as_tensor(t).detach().cpu().numpy(), dtype=np.float32) for t in (conv_w, conv_b, fc_w, fc_b)]
        if arrs[0].shape != (6, 3, 3, 3) or arrs[1].shape != (6,) or arrs[2].shape != (32, 6 * self.P * self.P) or arrs[3].shape != (32,):
            raise ValueError("expected Conv2d(3,6,3) and Linear(6*(N-2)^2, 32) parameters (config/default.yaml defaults)")
        self.lib = _capi.load()
        self._h = C.c_void_p()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _capi.check(self.lib.ssd_frontend_create(self.view, arrs[0].ctypes.data, arrs[1].ctypes.data, arrs[2].ctypes.data,
                                                 arrs[3].ctypes.data, float(negative_slope), idx, C.byref(self._h)))

    @classmethod
    def from_module(cls, conv_to_fc, view, device=None):
        """conv_to_fc: the reference's nn.Sequential(Conv2d, LeakyReLU, Flatten, Linear, LeakyReLU)."""
        conv, act, fc = conv_to_fc[0], conv_to_fc[1], conv_to_fc[3]
        device = device or conv.weight.device
        return cls(conv.weight, conv.bias, fc.weight, fc.bias, view, negative_slope=getattr(act, "negative_slope", 0.01), device=device)

    def forward(self, obs_buf: torch.Tensor, rows: int, agent_stride: int, plane_stride: int, row_stride: int, out=None):
        """obs_buf: u8 CUDA buffer holding `rows` agent views `agent_stride` bytes apart.  Returns f32 [rows, 32]."""
        if obs_buf.dtype != torch.uint8 or not obs_buf.is_cuda or not obs_buf.is_contiguous():
            raise ValueError("obs_buf must be a contiguous uint8 CUDA tensor")
        if obs_buf.numel() < rows * agent_stride:
            raise ValueError("obs_buf is smaller than rows * agent_stride")
        if out is None:
            out = torch.empty((rows, 32), dtype=torch.float32, device=obs_buf.device)
        with torch.cuda.device(obs_buf.device):
            _capi.check(self.lib.ssd_frontend_forward(self._h, obs_buf.data_ptr(), int(rows), int(agent_stride), int(plane_stride),
                                                      int(row_stride), out.data_ptr(),
                                                      C.c_void_p(torch.cuda.current_stream(obs_buf.device).cuda_stream)))
        return out

    def forward_env(self, env, obs_buf=None, out=None):
        """All B*n agent views of an ``SSDBatchEnv`` (row order [B][n], as ``batch['obs'].reshape(bs * n, ...)``)."""
        lay = env.layout
        return self.forward(env.obs_buf if obs_buf is None else obs_buf, env.B * env.n, lay.obs_agent_stride,
                            lay.obs_plane_stride, lay.obs_row_stride, out=out)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.ssd_frontend_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MacFrontEnd:
    """Reroutes ``mac.agent.rgb_preprocess`` (homophily_controller.py:132-136) to the fused kernel while a rollout is in
    progress.  The packed weights follow the module's parameters (re-packed when an optimiser step bumped their versions)."""

    def __init__(self, mac, env):
        self.mac, self.env = mac, env
        self.module = mac.agent.conv_to_fc
        self.original = mac.agent.rgb_preprocess
        self.active = False
        self.rows = None                                      # (first agent view, count) of the env range being rolled out
        self._fe, self._stamp = None, None
        mac.agent.rgb_preprocess = self

    def _front_end(self):
        stamp = tuple((p.data_ptr(), p._version) for p in self.module.parameters())
        if self._fe is None or stamp != self._stamp:
            if self._fe is not None:
                self._fe.close()
            self._fe, self._stamp = ObsFrontEnd.from_module(self.module, self.env.spec.view, device=self.env.device), stamp
        return self._fe

    def __call__(self, x):
        first, count = self.rows if self.rows is not None else (0, self.env.B * self.env.n)
        if not self.active or x.shape[0] != count:            # learner path (autograd) / foreign batch: the module
            return self.original(x)
        lay = self.env.layout
        return self._front_end().forward(self.env.obs_buf.view(-1)[first * lay.obs_agent_stride:], count, lay.obs_agent_stride,
                                         lay.obs_plane_stride, lay.obs_row_stride)

    def __deepcopy__(self, memo):
        """``copy.deepcopy(mac)`` (the learner's target network, homophily_learner.py:47) must not copy the kernel handle: the
        copy of the agent gets its own, unpatched ``rgb_preprocess`` back."""
        import types
        agent_copy = memo.get(id(self.mac.agent))
        if agent_copy is None:
            return self.original
        return types.MethodType(type(self.mac.agent).rgb_preprocess, agent_copy)

    def detach(self):
        self.mac.agent.rgb_preprocess = self.original
        if self._fe is not None:
            self._fe.close()
