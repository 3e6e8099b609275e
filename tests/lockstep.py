"""Shared helpers for parity tests: deterministic draw/action/teleport schedules
and trace recording.  The same schedule drives the reference (via
oracle/refshim.py, dev container only), the C oracle and the CUDA path.

All randomness comes from ``np.random.RandomState`` (frozen legacy stream), so a
golden trace only has to store the seed and the expected outputs.
"""
from __future__ import annotations

import numpy as np

# named parity configurations (BASELINE.json configs + the only Harvest setting that runs unmodified)
CONFIGS = {
    # key: (env name, map=, num_agents, view_size, obs_color)
    "cleanup5":        ("cleanup", "default5", 5, 7, "simplified"),
    "cleanup5_full":   ("cleanup", "default5", 5, 7, "full"),
    "cleanup10":       ("cleanup", "default10", 10, 7, "simplified"),
    "cleanup10_full":  ("cleanup", "default10", 10, 7, "full"),
    "cleanup3":        ("cleanup", "default3", 3, 7, "simplified"),
    "harvest5":        ("harvest", "default5", 5, 15, "simplified"),     # SPAWN_PROB injected
    "harvest10_n5":    ("harvest", "default10", 5, 15, "simplified"),    # runs unmodified in the reference
    "harvest10_full":  ("harvest", "default10", 10, 7, "full"),
}


class Schedule:
    """Deterministic per-step inputs for one env of a given geometry."""

    def __init__(self, seed, n, H, W, n_actions, wall, apple_pts, waste_pts, base_grid,
                 teleport_every=25, move_bias=0.5):
        self.rs = np.random.RandomState(seed)
        self.n, self.H, self.W, self.G, self.n_actions = n, H, W, H * W, n_actions
        self.wall = np.asarray(wall).reshape(H, W).astype(bool)
        self.apple_pts, self.waste_pts = np.asarray(apple_pts), np.asarray(waste_pts)
        self.base_grid = np.asarray(base_grid).reshape(H, W)
        self.teleport_every, self.move_bias = teleport_every, move_bias
        self.free = np.argwhere(~self.wall)

    def reset_draws(self):
        rs = self.rs
        return dict(spawn_key=rs.randint(0, 2 ** 32, size=(self.n, self.H, self.W), dtype=np.uint64).astype(np.uint32),
                    rot=rs.randint(0, 4, size=self.n).astype(np.uint8))

    def step_inputs(self, t):
        rs = self.rs
        if rs.rand() < self.move_bias:          # bias towards movement so collisions are common
            actions = rs.randint(0, 5, size=self.n)
            k = rs.randint(0, self.n)
            actions[k] = self.n_actions - 1
        else:
            actions = rs.randint(0, self.n_actions, size=self.n)
        draws = dict(
            prio=rs.permutation(self.n).astype(np.uint32) if rs.rand() < 0.5
            else rs.randint(0, 2 ** 32, size=self.n, dtype=np.uint64).astype(np.uint32),
            u_apple=rs.randint(0, 2 ** 32, size=(self.H, self.W), dtype=np.uint64).astype(np.uint32),
            u_waste=rs.randint(0, 2 ** 32, size=(self.H, self.W), dtype=np.uint64).astype(np.uint32),
            wkey=rs.randint(0, 2 ** 32, size=(self.H, self.W), dtype=np.uint64).astype(np.uint32))
        teleport = None
        if self.teleport_every and t > 0 and t % self.teleport_every == 0:
            teleport = self._teleport()
        return actions.astype(np.uint8), draws, teleport

    def _teleport(self):
        """Cluster all agents around a random free cell (forces collisions, swaps, chains, cycles,
        duplicate-occupancy states) and perturb the grid (cleaned waste, extra apples)."""
        rs = self.rs
        centre = self.free[rs.randint(len(self.free))]
        d = np.abs(self.free - centre).sum(axis=1) + rs.rand(len(self.free)) * 0.5
        near = self.free[np.argsort(d)[: self.n]]
        pos = near[rs.permutation(self.n)].copy()
        if self.n >= 2 and rs.rand() < 0.2:       # duplicate occupancy (reachable via SURVEY D1)
            pos[rs.randint(1, self.n)] = pos[0]
        orient = rs.randint(0, 4, size=self.n).astype(np.uint8)
        grid = None
        if rs.rand() < 0.7:
            grid = self.base_grid.copy()
            flat = grid.reshape(-1)
            if len(self.waste_pts):
                frac = rs.choice([0.0, 0.02, 0.3, 0.7, 1.0])
                m = rs.rand(len(self.waste_pts)) < frac
                flat[self.waste_pts[m]] = 4        # cleaned waste becomes river
            if len(self.apple_pts):
                frac = rs.choice([0.0, 0.1, 0.5, 0.95])
                m = rs.rand(len(self.apple_pts)) < frac
                flat[self.apple_pts] = np.where(m, 2, 0)
        return dict(pos=pos.astype(np.int32), orient=orient, grid=grid)


TRACE_KEYS = ("grid", "pos", "orient", "reward", "clean", "apple_cnt", "done", "obs", "state")


def new_trace():
    return {k: [] for k in TRACE_KEYS}


def finish_trace(tr):
    return {k: np.stack(v) for k, v in tr.items()}


# ---------------------------------------------------------------------------
# backends: one env instance behind a tiny common surface
# ---------------------------------------------------------------------------
def spec_for(key, episode_limit=1000):
    from homophily_marl_b200 import mapspec
    name, map_name, n, view, color = CONFIGS[key]
    return mapspec.compile_map(name, map_name, n, view, episode_limit, obs_color=color)


def schedule_for(key, seed, **kw):
    s = spec_for(key)
    return Schedule(seed, s.n_agents, s.H, s.W, s.n_actions, s.wall, s.apple_pts, s.waste_pts, s.base_grid, **kw)


class OracleBackend:
    def __init__(self, key, random_spawn=False, episode_limit=1000):
        from oracle.oracle import OracleBatch
        self.spec = spec_for(key, episode_limit)
        self.o = OracleBatch.from_spec(self.spec, n_envs=1, random_spawn_point=random_spawn,
                                       spawn_rotation=None if random_spawn else 0)

    def reset(self, draws):
        d = dict(draws)
        d["spawn_key"] = d["spawn_key"].reshape(self.spec.n_agents, -1)
        self.o.reset_one(0, d)

    def set_state(self, pos, orient, grid):
        self.o.set_state(0, grid=grid, pos_rc=pos, orient=orient)

    def step(self, actions, draws):
        return self.o.step_one(actions, 0, {k: v.reshape(-1) for k, v in draws.items()})

    def snapshot(self):
        return dict(grid=self.o.grid[0].copy(), pos=self.o.pos_rc[0].astype(np.int32), orient=self.o.orient[0].copy(),
                    obs=self.o.obs_one(0), state=self.o.state_one(0))


def run_trace(backend, sched, T):
    """reset + T steps; records everything the parity contract covers."""
    tr = new_trace()
    backend.reset(sched.reset_draws())
    snap0 = backend.snapshot()
    for t in range(T):
        actions, draws, tele = sched.step_inputs(t)
        if tele is not None:
            backend.set_state(tele["pos"], tele["orient"], tele["grid"])
        reward, clean, cnt, done = backend.step(actions, draws)
        s = backend.snapshot()
        for k in ("grid", "pos", "orient", "obs", "state"):
            tr[k].append(s[k])
        tr["reward"].append(np.asarray(reward, dtype=np.int8))
        tr["clean"].append(np.asarray(clean, dtype=np.uint8))
        tr["apple_cnt"].append(np.uint16(cnt))
        tr["done"].append(np.uint8(done))
    out = finish_trace(tr)
    for k in ("grid", "pos", "orient", "obs", "state"):
        out["reset_" + k] = snap0[k]
    return out


def assert_traces_equal(a, b, what=""):
    for k in a:
        if k not in b:
            continue
        if not np.array_equal(a[k], b[k]):
            x, y = np.asarray(a[k]), np.asarray(b[k])
            if x.shape != y.shape:
                raise AssertionError(f"{what}: {k} shape {x.shape} vs {y.shape}")
            bad = np.argwhere(x != y)
            raise AssertionError(f"{what}: first mismatch in '{k}' at index {bad[0].tolist()} "
                                 f"({x[tuple(bad[0])]} vs {y[tuple(bad[0])]}); {len(bad)} differing elements")


class CudaBackend:
    """One env instance on the GPU behind the lockstep surface (calls go through the C ABI)."""

    def __init__(self, key, random_spawn=False, episode_limit=1000, device="cuda:0", **kw):
        import torch
        from homophily_marl_b200.batch_env import SSDBatchEnv
        name, map_name, n, view, color = CONFIGS[key]
        extra = dict(obs_color=color)
        if random_spawn:
            extra.update(random_spawn_point=True, random_spawn_rotation=None)
        self.torch = torch
        self.env = SSDBatchEnv(name, 1, n, map=map_name, view_size=view, episode_limit=episode_limit,
                               extra_args=extra, device=device, want_state=True, **kw)

    def reset(self, draws):
        e = self.env
        e.reset(draws=dict(spawn_key=draws["spawn_key"].reshape(1, e.n, e.G), rot=draws["rot"].reshape(1, e.n)))

    def set_state(self, pos, orient, grid):
        self.env.set_state(0, grid=grid, pos_rc=pos, orient=orient)

    def step(self, actions, draws):
        e, torch = self.env, self.torch
        a = torch.as_tensor(np.asarray(actions, dtype=np.uint8).reshape(1, e.n), device=e.device)
        e.step(a, draws={k: v.reshape(1, -1) for k, v in draws.items()}, want_state=True)
        return (e.reward[0].cpu().numpy(), e.clean[0].cpu().numpy(),
                int(e.apple_cnt[0].cpu().numpy().view(np.uint16)), bool(e.done[0].item()))

    def snapshot(self):
        e = self.env
        e.render(want_obs=True, want_state=True)
        return dict(grid=e.grid[0].cpu().numpy(), pos=e.agent_pos[0].cpu().numpy().astype(np.int32),
                    orient=e.agent_orient[0].cpu().numpy(), obs=e.obs_view()[0].cpu().numpy(),
                    state=e.state_rgb[0].cpu().numpy())


def cluster_states(rs, spec, B):
    """[B,n,2] positions clustered around random free cells (+ some duplicates) and [B,n] orientations."""
    wall = np.asarray(spec.wall).reshape(spec.H, spec.W).astype(bool)
    free = np.argwhere(~wall)
    n = spec.n_agents
    pos = np.zeros((B, n, 2), dtype=np.int32)
    for b in range(B):
        centre = free[rs.randint(len(free))]
        d = np.abs(free - centre).sum(axis=1) + rs.rand(len(free)) * 0.5
        near = free[np.argsort(d)[:n]]
        pos[b] = near[rs.permutation(n)]
        if n >= 2 and rs.rand() < 0.2:
            pos[b, rs.randint(1, n)] = pos[b, 0]
    return pos, rs.randint(0, 4, size=(B, n)).astype(np.uint8)
