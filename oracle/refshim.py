"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the *unmodified* reference SSD envs from ``/root/reference/src`` (this
container only; the GPU box has no /root/reference) and drives them with
injected, position-indexed random draws so that the reference, the C oracle
(``oracle/ssd_oracle.c``) and the CUDA path all consume identical randomness.

Used by ``tests/golden/make_golden.py`` (fixture generator) and by the
``needs_reference`` tests that pin the oracle against the live reference.

Reference call sites that are intercepted (no reference file is edited;
module attributes are looked up at call time):
  * ``np.random.shuffle``  map_env.py:541      mover priority
  * ``np.random.rand``     cleanup.py:172,183  harvest.py:119   spawn draws
  * ``random.shuffle``     cleanup.py:178      map_env.py:777   waste order / spawn points
  * ``np.random.randint``  map_env.py:789      spawn rotation
"""
import contextlib
import io
import os
import random
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("SSD_REFERENCE_ROOT", "/root/reference")

# cell codes shared by oracle / kernels / tests
CODES = {' ': 0, '@': 1, 'A': 2, 'H': 3, 'R': 4, 'S': 5}
ORIENT_NAMES = ['LEFT', 'RIGHT', 'UP', 'DOWN']      # index == position in map_env.ORIENTATIONS
TWO32 = float(2 ** 32)


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "src", "envs", "ssd"))


def _import_registry():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]
    src = os.path.join(REF_ROOT, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    from envs import REGISTRY  # noqa: E402  (reference package)
    return REGISTRY


def make(name, num_agents, map_name, view_size, episode_limit=1000, harvest_spawn_prob=None, **extra):
    """Construct a reference env exactly as ``episode_runner.py:15`` does."""
    registry = _import_registry()
    extra_args = dict(random_spawn_point=False, random_spawn_rotation=0,
                      disable_rotation_action=True, disable_fire_action=True,
                      obs_color='simplified')
    extra_args.update(extra)
    with contextlib.redirect_stdout(io.StringIO()):
        env = registry[name](num_agents=num_agents, render=False, episode_limit=episode_limit,
                             is_replay=False, view_size=view_size, map=map_name,
                             extra_args=extra_args)
    if name == 'harvest' and not hasattr(env, 'SPAWN_PROB'):
        # the reference leaves SPAWN_PROB unset unless map == "default10"
        import envs.ssd.harvest as hv
        env.SPAWN_PROB = list(harvest_spawn_prob if harvest_spawn_prob is not None else hv.SPAWN_PROB)
    elif name == 'harvest' and harvest_spawn_prob is not None:
        env.SPAWN_PROB = list(harvest_spawn_prob)
    return env


class Injector:
    """Feeds position-indexed draws to one reference env.

    Before ``env.step``: set ``prio`` (u32[n]), ``u_apple``/``u_waste``/``wkey`` (u32[H,W]).
    Before ``env.reset``: set ``spawn_key`` (u32[n,H,W]) and ``rot`` (int[n]).
    A u32 draw k is presented to the reference as the float64 k / 2**32, so
    ``rand < p``  <=>  ``k < ceil(p * 2**32)`` exactly.
    """

    def __init__(self, env):
        self.env = env
        self.W = env.base_map.shape[1]
        self.prio = self.u_apple = self.u_waste = self.wkey = None
        self.spawn_key = self.rot = None
        self.in_waste = False
        self._spawn_calls = 0
        self._rot_calls = 0

    # -- context manager ------------------------------------------------
    def __enter__(self):
        self._saved = (np.random.rand, np.random.shuffle, random.shuffle, np.random.randint)
        np.random.rand, np.random.shuffle = self._rand, self._npshuffle
        random.shuffle, np.random.randint = self._pyshuffle, self._randint
        return self

    def __exit__(self, *exc):
        np.random.rand, np.random.shuffle, random.shuffle, np.random.randint = self._saved
        return False

    def begin_step(self):
        self.in_waste = False

    def begin_reset(self):
        self.in_waste = False
        self._spawn_calls = 0
        self._rot_calls = 0

    # -- shims ------------------------------------------------------------
    def _rand(self, k):
        frame = sys._getframe(1)
        row, col = frame.f_locals['row'], frame.f_locals['col']
        waste = frame.f_code.co_name == 'spawn_apples_and_waste' and self.in_waste
        src = self.u_waste if waste else self.u_apple
        return np.array([float(src[row][col]) / TWO32])

    def _npshuffle(self, lst):
        lst.sort(key=lambda t: (int(self.prio[int(t[0].split('-')[1])]), int(t[0].split('-')[1])))

    def _pyshuffle(self, lst):
        W = self.W
        if lst is getattr(self.env, 'waste_points', None):
            self.in_waste = True
            lst.sort(key=lambda rc: (int(self.wkey[rc[0]][rc[1]]), rc[0] * W + rc[1]))
        elif lst is self.env.spawn_points:
            key = self.spawn_key[self._spawn_calls]
            self._spawn_calls += 1
            lst.sort(key=lambda rc: (int(key[rc[0]][rc[1]]), rc[0] * W + rc[1]))
        else:  # pragma: no cover
            raise AssertionError("unexpected random.shuffle call site")

    def _randint(self, n):
        assert n == 4
        r = int(self.rot[self._rot_calls])
        self._rot_calls += 1
        return r


# ---------------------------------------------------------------------------
# state extraction / injection (uses only public attributes of the reference)
# ---------------------------------------------------------------------------
def grid_codes(env):
    """world_map (<U1) -> u8 codes [H, W]."""
    wm = env.world_map
    out = np.zeros(wm.shape, dtype=np.uint8)
    for ch, code in CODES.items():
        out[wm == ch] = code
    return out


def agent_pos(env):
    return np.array([env.agents['agent-%d' % i].get_pos() for i in range(env.num_agents)], dtype=np.int32)


def agent_orient(env):
    return np.array([ORIENT_NAMES.index(env.agents['agent-%d' % i].get_orientation())
                     for i in range(env.num_agents)], dtype=np.uint8)


def set_state(env, grid=None, pos=None, orient=None):
    inv = {v: k for k, v in CODES.items()}
    if grid is not None:
        for r in range(grid.shape[0]):
            for c in range(grid.shape[1]):
                env.world_map[r, c] = inv[int(grid[r, c])]
    for i in range(env.num_agents):
        ag = env.agents['agent-%d' % i]
        if pos is not None:
            ag.set_pos(np.array(pos[i]))
        if orient is not None:
            ag.set_orientation(ORIENT_NAMES[int(orient[i])])


def obs_u8(env):
    """get_obs() (float64 k/256) -> exact u8 [n, 3, N, N]."""
    o = np.stack(env.get_obs()) * 256.0
    r = np.rint(o)
    assert np.array_equal(o, r)
    return r.astype(np.uint8)


def state_u8(env):
    s = env.get_state() * 256.0
    r = np.rint(s)
    assert np.array_equal(s, r)
    return r.astype(np.uint8)
