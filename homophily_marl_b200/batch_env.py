"""Batched, HBM-resident SSD env (Cleanup / Harvest): the tensor-level API.

``SSDBatchEnv`` owns B env instances on one B200.  State lives in PyTorch-owned
device tensors (packed u8 grids + one u32 per agent); every transition and every
observation is produced by the sm_100a kernels behind the C ABI
(``include/ssd_b200.h``).  PyTorch is only the allocator / stream provider here.

Replaces, batched: MapEnv.reset / step / get_obs / get_state
(src/envs/ssd/map_env.py:874-957, 986-993 of the reference).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _capi, mapspec


def _require_cuda(device):
    if not torch.cuda.is_available():
        raise RuntimeError("homophily_marl_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"device must be a CUDA device, got {device!r}")
    return torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class SSDBatchEnv:
    def __init__(self, name, n_envs, num_agents, map="default", view_size=7, episode_limit=100,
                 extra_args=None, seed=0, device="cuda:0", env_gid_base=0, rows=None, params=None,
                 fire_cost=1, hit_penalty=0, want_state=False, check_actions=False):
        extra = dict(random_spawn_point=False, random_spawn_rotation=0, disable_rotation_action=True,
                     disable_fire_action=True, obs_color="simplified")
        extra.update(extra_args or {})
        self.extra_args = extra
        self.name = name
        self.spec = mapspec.compile_map(name, map, num_agents, view_size, episode_limit,
                                        obs_color=extra["obs_color"], rows=rows, params=params,
                                        fire_cost=fire_cost, hit_penalty=hit_penalty)
        self.device = _require_cuda(device)
        # debug switch (one device sync per step): out-of-range action codes raise KeyError like the reference's action_map
        self.check_actions = bool(check_actions) or os.environ.get("SSD_B200_CHECK_ACTIONS") == "1"
        self.lib = _capi.load()
        s = self.spec
        self.B, self.n, self.H, self.W, self.G, self.N = int(n_envs), s.n_agents, s.H, s.W, s.G, s.N
        self.n_actions, self.episode_limit = s.n_actions, s.episode_limit
        self.seed, self.env_gid_base = int(seed), int(env_gid_base)

        cfg = _capi.SsdConfig()
        cfg.kind, cfg.n_envs, cfg.n_agents = s.kind, self.B, s.n_agents
        cfg.height, cfg.width, cfg.view, cfg.episode_limit = s.H, s.W, s.view, s.episode_limit
        cfg.fire_cost, cfg.hit_penalty, cfg.beam_len = s.fire_cost, s.hit_penalty, s.beam_len
        cfg.random_spawn_point = int(bool(extra["random_spawn_point"]))
        rot = extra["random_spawn_rotation"]
        cfg.spawn_rotation = -1 if rot is None else int(rot)
        cfg.device = self.device.index
        cfg.seed, cfg.env_gid_base = self.seed & (2 ** 64 - 1), self.env_gid_base & 0xFFFFFFFF
        self._ascii = s.ascii_bytes()
        cfg.ascii_map = self._ascii
        self._thr_a = np.ascontiguousarray(s.thr_apple, dtype=np.uint32)
        self._thr_w = np.ascontiguousarray(s.thr_waste, dtype=np.uint32)
        cfg.n_waste_lut = len(self._thr_a)
        cfg.thr_apple, cfg.thr_waste = self._thr_a.ctypes.data, self._thr_w.ctypes.data
        for k in range(4):
            cfg.thr_harvest[k] = int(s.thr_harvest[k])
        for i in range(16):
            for ch in range(3):
                cfg.color_lut[i][ch] = int(s.lut[i, ch])
        self._h = C.c_void_p()
        _capi.check(self.lib.ssd_create(C.byref(cfg), C.byref(self._h)))
        lay = _capi.SsdLayout()
        _capi.check(self.lib.ssd_get_layout(self._h, C.byref(lay)))
        self.layout = lay
        assert lay.n_actions == self.n_actions and lay.n_cells == self.G

        dev, B, n = self.device, self.B, self.n
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)  # noqa: E731
        self.grid_buf = z((B, lay.grid_stride), torch.uint8)
        self.agent_buf = z((B, lay.agent_stride), torch.int32)
        self.ep_ret_buf = z((B, lay.agent_stride), torch.int32)
        self.t_buf = z((B,), torch.int32)
        self.tick_buf = z((B,), torch.int32)
        self.counts_buf = z((B,), torch.int32)                 # #apple cells | #waste cells << 16 of every grid
        self.reward = z((B, n), torch.int8)
        self.clean = z((B, n), torch.uint8)
        self.apple_cnt = z((B,), torch.int16)
        self.done = z((B,), torch.uint8)
        self.obs_buf = self.new_obs_buffer()
        self.state_rgb = z((B, 3, self.H, self.W), torch.uint8) if want_state else None
        self._st = _capi.SsdState(*[t.data_ptr() for t in (self.grid_buf, self.agent_buf, self.ep_ret_buf,
                                                            self.t_buf, self.tick_buf, self.counts_buf)])
        self._keep = None

    # ------------------------------------------------------------------ buffers / views
    def new_obs_buffer(self):
        """Flat u8 [B, obs_env_stride] buffer in the kernel's padded plane layout."""
        return torch.zeros((self.B, self.layout.obs_env_stride), dtype=torch.uint8, device=self.device)

    def obs_view(self, buf=None):
        """[B, n, 3, N, N] u8 view (no copy) of an obs buffer; ``view / 256`` equals get_obs()."""
        buf = self.obs_buf if buf is None else buf
        lay = self.layout
        return buf.as_strided((self.B, self.n, 3, self.N, self.N),
                              (lay.obs_env_stride, lay.obs_agent_stride, lay.obs_plane_stride, lay.obs_row_stride, 1))

    @property
    def grid(self):
        return self.grid_buf[:, :self.G].view(self.B, self.H, self.W)

    @property
    def agent_pos(self):
        """[B, n, 2] (row, col) int32."""
        a = self.agent_buf[:, :self.n]
        return torch.stack([a & 0xFF, (a >> 8) & 0xFF], dim=-1)

    @property
    def agent_orient(self):
        return ((self.agent_buf[:, :self.n] >> 16) & 3).to(torch.uint8)

    @property
    def ep_ret(self):
        return self.ep_ret_buf[:, :self.n]

    def set_state(self, b, grid=None, pos_rc=None, orient=None):
        """Test/injection hook (get/set_state_tensors)."""
        if grid is not None:
            g = torch.as_tensor(np.asarray(grid, dtype=np.uint8).reshape(-1), device=self.device)
            self.grid_buf[b, :self.G] = g
            self._recount(slice(b, b + 1))
        if pos_rc is not None or orient is not None:
            a = self.agent_buf[b, :self.n].cpu().numpy().astype(np.int64)
            if pos_rc is not None:
                pr = np.asarray(pos_rc, dtype=np.int64)
                a = (a & ~0xFFFF) | pr[:, 0] | (pr[:, 1] << 8)
            if orient is not None:
                a = (a & 0xFFFF) | (np.asarray(orient, dtype=np.int64) << 16)
            self.agent_buf[b, :self.n] = torch.as_tensor(a.astype(np.int32), device=self.device)

    def _recount(self, sel):
        """ssd_state.counts of the selected envs from their grids (needed after writing a grid directly)."""
        g = self.grid_buf[sel, :self.G]
        self.counts_buf[sel] = ((g == 2).sum(1) | ((g == 3).sum(1) << 16)).to(torch.int32)

    def load_state(self, grid=None, pos_rc=None, orient=None, t=None):
        """Batched variant of ``set_state``: NumPy arrays [B,H,W] / [B,n,2] / [B,n] / [B]."""
        if grid is not None:
            g = np.ascontiguousarray(grid, dtype=np.uint8).reshape(self.B, self.G)
            self.grid_buf[:, :self.G] = torch.as_tensor(g, device=self.device)
            self._recount(slice(None))
        if pos_rc is not None or orient is not None:
            a = self.agent_buf[:, :self.n].cpu().numpy().astype(np.int64)
            if pos_rc is not None:
                pr = np.asarray(pos_rc, dtype=np.int64)
                a = (a & ~0xFFFF) | pr[..., 0] | (pr[..., 1] << 8)
            if orient is not None:
                a = (a & 0xFFFF) | (np.asarray(orient, dtype=np.int64) << 16)
            self.agent_buf[:, :self.n] = torch.as_tensor(a.astype(np.int32), device=self.device)
        if t is not None:
            self.t_buf.copy_(torch.as_tensor(np.asarray(t, dtype=np.int32), device=self.device))

    # ------------------------------------------------------------------ draws
    def _draws(self, draws):
        if not draws:
            return None
        keep = {}
        for k in ("prio", "u_apple", "u_waste", "wkey", "spawn_key"):
            v = draws.get(k)
            if v is not None:
                v = np.ascontiguousarray(v, dtype=np.uint32).view(np.int32)
                keep[k] = torch.as_tensor(v, device=self.device).contiguous()
        if draws.get("rot") is not None:
            keep["rot"] = torch.as_tensor(np.ascontiguousarray(draws["rot"], dtype=np.uint8), device=self.device)
        expect = dict(prio=self.B * self.n, u_apple=self.B * self.G, u_waste=self.B * self.G, wkey=self.B * self.G,
                      spawn_key=self.B * self.n * self.G, rot=self.B * self.n)
        for k, t in keep.items():
            if t.numel() != expect[k]:
                raise ValueError(f"draws[{k!r}] has {t.numel()} elements, expected {expect[k]}")
        self._keep = keep
        return _capi.SsdDraws(*[keep[k].data_ptr() if k in keep else None
                                for k in ("prio", "u_apple", "u_waste", "wkey", "spawn_key", "rot")])

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ entry points
    def reset(self, mask=None, draws=None, obs=True, obs_out=None):
        d = self._draws(draws)
        out = None if not obs else (self.obs_buf if obs_out is None else obs_out)
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        _capi.check(self.lib.ssd_reset(self._h, C.byref(self._st), _ptr(mask), C.byref(d) if d else None,
                                       _ptr(out), self._stream()))

    def _check_actions(self, actions):
        if actions.dtype != torch.uint8 or not actions.is_cuda or not actions.is_contiguous() or actions.numel() != self.B * self.n:
            raise ValueError("actions must be a contiguous uint8 CUDA tensor of shape [B, n]")
        if self.check_actions and bool((actions >= self.n_actions).any()):
            # the reference raises KeyError from action_map (agent.py:176,237); the kernel alone would treat the code as a no-op
            raise KeyError(int(actions.max()))

    def _step_out(self, obs_out, want_obs, want_state):
        return _capi.SsdStepOut(self.reward.data_ptr(), self.clean.data_ptr(), self.apple_cnt.data_ptr(),
                                self.done.data_ptr(),
                                (self.obs_buf if obs_out is None else obs_out).data_ptr() if want_obs else None,
                                self.state_rgb.data_ptr() if (want_state and self.state_rgb is not None) else None)

    def step(self, actions, draws=None, obs_out=None, want_obs=True, want_state=False, auto_reset=False):
        """actions: u8 CUDA tensor [B, n].  Results land in self.reward / clean / apple_cnt / done / obs.

        ``auto_reset=True``: every env whose ``done`` flag this step raised is reset by a masked ``ssd_reset`` launch on the same
        stream (no host sync), and its slot of the observation buffer then holds the FIRST observation of the new episode;
        ``done`` / ``reward`` keep the values of the finished step.  Envs may therefore run on different episode clocks
        (``load_state(t=...)``), which the reference -- one env, ``get_done()`` constantly False -- never needs."""
        self._check_actions(actions)
        d = self._draws(draws)
        so = self._step_out(obs_out, want_obs, want_state)
        _capi.check(self.lib.ssd_step(self._h, C.byref(self._st), _ptr(actions), C.byref(d) if d else None,
                                      C.byref(so), self._stream()))
        if auto_reset:
            out = None if not want_obs else (self.obs_buf if obs_out is None else obs_out)
            _capi.check(self.lib.ssd_reset(self._h, C.byref(self._st), _ptr(self.done), None, _ptr(out), self._stream()))
            if want_state and self.state_rgb is not None:
                self.render(want_obs=False, want_state=True)

    def step_range(self, actions, env_begin, env_count, draws=None, obs_out=None, want_obs=True, want_state=False):
        """``step`` restricted to envs [env_begin, env_begin + env_count) (``ssd_step_range``).  ``actions`` is the
        whole-batch [B, n] tensor; disjoint ranges may run concurrently on different CUDA streams (the launch goes to the
        caller's current stream), which lets one group's policy / logic overlap another group's observation stores."""
        self._check_actions(actions)
        d = self._draws(draws)
        so = self._step_out(obs_out, want_obs, want_state)
        _capi.check(self.lib.ssd_step_range(self._h, C.byref(self._st), _ptr(actions), C.byref(d) if d else None,
                                            C.byref(so), int(env_begin), int(env_count), self._stream()))

    def render(self, obs_out=None, want_obs=True, want_state=False):
        o = (self.obs_buf if obs_out is None else obs_out) if want_obs else None
        s = self.state_rgb if want_state else None
        _capi.check(self.lib.ssd_render(self._h, C.byref(self._st), _ptr(o), _ptr(s), self._stream()))

    def make_host_io(self, with_obs=True, with_state=False):
        """Pinned host mirrors for ``step_host`` (the reference-facing call with host buffers)."""
        pin = lambda t: torch.zeros(t.shape, dtype=t.dtype).pin_memory()  # noqa: E731
        io = dict(actions=torch.zeros((self.B, self.n), dtype=torch.uint8).pin_memory(),
                  reward=pin(self.reward), clean=pin(self.clean), apple_cnt=pin(self.apple_cnt), done=pin(self.done),
                  obs=pin(self.obs_buf) if with_obs else None,
                  state=pin(self.state_rgb) if (with_state and self.state_rgb is not None) else None,
                  agent=pin(self.agent_buf))
        io["d_actions"] = torch.zeros((self.B, self.n), dtype=torch.uint8, device=self.device)
        return io

    def step_host(self, io):
        """H2D actions -> step -> D2H reward/clean/apple_cnt/done(/obs/state) -> stream sync, all inside the C call."""
        want_state = io.get("state") is not None
        d_out = self._step_out(None, io["obs"] is not None, want_state)
        h_out = _capi.SsdStepOut(io["reward"].data_ptr(), io["clean"].data_ptr(), io["apple_cnt"].data_ptr(),
                                 io["done"].data_ptr(), io["obs"].data_ptr() if io["obs"] is not None else None,
                                 io["state"].data_ptr() if want_state else None)
        _capi.check(self.lib.ssd_step_host(self._h, C.byref(self._st), _ptr(io["actions"]), _ptr(io["d_actions"]),
                                           C.byref(d_out), C.byref(h_out), self._stream()))

    def pull_host(self, io, obs=True, state=True, agent=True):
        """Copies the current observation / state image / agent records into the pinned mirrors (one sync)."""
        if obs and io.get("obs") is not None:
            io["obs"].copy_(self.obs_buf, non_blocking=True)
        if state and io.get("state") is not None:
            self.render(want_obs=False, want_state=True)
            io["state"].copy_(self.state_rgb, non_blocking=True)
        if agent:
            io["agent"].copy_(self.agent_buf, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()

    @property
    def launch_count(self):
        return int(self.lib.ssd_launch_count(self._h))

    def bytes_per_env_step(self):
        """ALGORITHMIC HBM bytes of one fused step+obs: 2G + n(3N^2 + 11) + 3."""
        return 2 * self.G + self.n * (3 * self.N * self.N + 11) + 3

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.ssd_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
