"""Generates the committed golden traces from the UNMODIFIED reference.

Run in the dev container only (needs /root/reference):
    python tests/golden/make_golden.py
Each ``<config>.npz`` holds the schedule seed and, per step, the reference's
grid / positions / orientations / rewards / clean_num / apple count / done /
u8 observations / u8 global state.  The schedule itself (actions, injected
draws, teleports) is regenerated from the seed by tests/lockstep.py.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import refshim as rs  # noqa: E402
import lockstep as ls  # noqa: E402


class RefBackend:
    """The reference env + Injector behind the lockstep surface."""

    def __init__(self, key, random_spawn=False, episode_limit=1000):
        name, map_name, n, view, color = ls.CONFIGS[key]
        self.spec = ls.spec_for(key, episode_limit)
        extra = dict(obs_color=color)
        if random_spawn:
            extra.update(random_spawn_point=True, random_spawn_rotation=None)
        self.env = rs.make(name, n, map_name, view, episode_limit=episode_limit,
                           harvest_spawn_prob=self.spec.params.spawn_prob if name == "harvest" else None, **extra)
        self.inj = rs.Injector(self.env)

    def reset(self, draws):
        self.inj.spawn_key, self.inj.rot = draws["spawn_key"], draws["rot"]
        self.inj.u_apple = self.inj.u_waste = self.inj.wkey = np.zeros((self.spec.H, self.spec.W), np.uint32)
        self.inj.begin_reset()
        with self.inj:
            self.env.reset()

    def set_state(self, pos, orient, grid):
        rs.set_state(self.env, grid=grid, pos=pos, orient=orient)

    def step(self, actions, draws):
        for k, v in draws.items():
            setattr(self.inj, k, v)
        self.inj.begin_step()
        with self.inj:
            reward, term, info = self.env.step(actions)
        cnt = int(round(float(info["apple_den"][0]) * self.spec.G))
        assert abs(cnt / self.spec.G - float(info["apple_den"][0])) < 1e-12
        return reward.astype(np.int8), info["clean_num"].astype(np.uint8), cnt, bool(term)

    def snapshot(self):
        return dict(grid=rs.grid_codes(self.env), pos=rs.agent_pos(self.env), orient=rs.agent_orient(self.env),
                    obs=rs.obs_u8(self.env), state=rs.state_u8(self.env))


GOLDEN = [  # (config, seed, steps, random_spawn)
    ("cleanup5", 11, 160, False), ("cleanup5_full", 12, 120, True), ("cleanup10", 13, 160, True),
    ("cleanup10_full", 14, 100, False), ("cleanup3", 15, 160, True), ("harvest5", 16, 120, False),
    ("harvest10_n5", 17, 120, True), ("harvest10_full", 18, 120, True),
]


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for key, seed, T, rnd in GOLDEN:
        be = RefBackend(key, random_spawn=rnd)
        tr = ls.run_trace(be, ls.schedule_for(key, seed), T)
        tr["seed"], tr["steps"], tr["random_spawn"] = np.int64(seed), np.int64(T), np.bool_(rnd)
        tr["ascii_sha"] = np.array(rs_map_sha(be.env))
        np.savez_compressed(os.path.join(out_dir, key + ".npz"), **tr)
        print(key, "steps", T, "returns", tr["reward"].sum(axis=0).tolist(),
              "bytes", os.path.getsize(os.path.join(out_dir, key + ".npz")))


def rs_map_sha(env):
    import hashlib
    rows = ["".join(r) for r in env.base_map]
    return hashlib.sha256("\n".join(rows).encode()).hexdigest()[:16]


if __name__ == "__main__":
    main()
