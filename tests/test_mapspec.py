"""CPU suite: the host-side map compiler against the oracle's independently derived tables."""
import ctypes as C

import numpy as np
import pytest

import lockstep as ls
from homophily_marl_b200 import mapspec
from oracle import oracle as O


def _oracle_map(spec):
    ob = O.OracleBatch.from_spec(spec, n_envs=1)
    raw = ob._map

    class M(C.Structure):
        _fields_ = [("ints", C.c_int * 13), ("base", C.c_uint8 * 2048), ("wall", C.c_uint8 * 2048),
                    ("n_apple", C.c_int), ("n_waste", C.c_int), ("n_spawn", C.c_int),
                    ("apple_pts", C.c_int * 2048), ("waste_pts", C.c_int * 2048), ("spawn_pts", C.c_int * 2048),
                    ("dbl", C.c_double * 8), ("thr_apple", C.c_uint32 * 2049), ("thr_waste", C.c_uint32 * 2049),
                    ("thr_harvest", C.c_uint32 * 4), ("color", (C.c_uint8 * 3) * 16)]
    assert C.sizeof(M) == len(raw)
    return M.from_buffer_copy(raw.tobytes())


@pytest.mark.parametrize("key", sorted(ls.CONFIGS))
def test_tables_match_oracle(key):
    spec = ls.spec_for(key)
    m = _oracle_map(spec)
    assert list(m.apple_pts[: m.n_apple]) == spec.apple_pts.tolist()
    assert list(m.waste_pts[: m.n_waste]) == spec.waste_pts.tolist()
    assert list(m.spawn_pts[: m.n_spawn]) == spec.spawn_pts.tolist()
    P = len(spec.waste_pts)
    assert list(m.thr_apple[: P + 1]) == spec.thr_apple.tolist()
    assert list(m.thr_waste[: P + 1]) == spec.thr_waste.tolist()
    assert list(m.thr_harvest) == spec.thr_harvest.tolist()
    assert np.array_equal(np.array(m.color, dtype=np.uint8), spec.lut)
    assert np.array_equal(np.frombuffer(bytes(m.wall[: spec.G]), dtype=np.uint8), spec.wall)


def test_survey_table_a4():
    want = {("cleanup", "default3"): (10, 10, 12, 8, 3, 4), ("cleanup", "default5"): (25, 18, 103, 55, 5, 55),
            ("cleanup", "default10"): (48, 18, 206, 110, 10, 109), ("cleanup", "whatever"): (25, 18, 103, 55, 5, 55)}
    for (name, mp), (H, W, na, nw, ns, first_zero) in want.items():
        s = mapspec.compile_map(name, mp, 1, 7, 100)
        assert (s.H, s.W, len(s.apple_pts), len(s.waste_pts), len(s.spawn_pts)) == (H, W, na, nw, ns)
        assert int(np.argmax(s.thr_apple == 0)) == first_zero and (s.thr_apple[first_zero:] == 0).all()
        assert (s.thr_waste[:first_zero] == 2 ** 31).all() and (s.thr_waste[first_zero:] == 0).all()
    s = mapspec.compile_map("cleanup", "default3", 3, 7, 100)
    got = [t / 2.0 ** 32 for t in s.thr_apple[:3]]
    assert np.allclose(got, [0.3, 0.20625, 0.1125], atol=1e-9)          # SURVEY A.4 worked example
    h = mapspec.compile_map("harvest", "default10", 5, 15, 100)
    assert (h.H, h.W, len(h.apple_pts), len(h.spawn_pts)) == (9, 38, 57, 10)
    assert h.params.spawn_prob == (0.0, 0.05, 0.08, 0.1)
    assert mapspec.compile_map("harvest", "default5", 5, 15, 100).params.spawn_prob == (0.0, 0.005, 0.02, 0.05)


def test_errors_and_agent_chars():
    with pytest.raises(AssertionError, match="not enough spawn points"):
        mapspec.compile_map("cleanup", "default3", 4, 7, 100)
    with pytest.raises(KeyError):
        mapspec.params_for("nope", "default")
    assert [mapspec.agent_char(i) for i in range(11)] == [1, 2, 3, 4, 5, 6, 7, 8, 9, 1, 1]
    for k, sha in mapspec.MAP_SHA.items():
        assert mapspec.map_sha(mapspec.ascii_map(k)) == sha


def test_decoded_maps_equal_the_reference_constants_cell_for_cell():
    """The run-length rows in mapspec.py decode to exactly the ASCII maps of src/envs/ssd/constants.py (dev container only)."""
    from oracle import refshim
    if not refshim.reference_available():
        pytest.skip("/root/reference not present")
    refshim._import_registry()
    import envs.ssd.constants as C
    from homophily_marl_b200 import mapspec
    for key, ref in (("cleanup_n3", C.CLEANUP_N3_MAP), ("cleanup_n5", C.CLEANUP_N5_MAP), ("cleanup_n10", C.CLEANUP_N10_MAP),
                     ("harvest_n10", C.HARVEST_N10_MAP)):
        assert mapspec.ascii_map(key) == list(ref), key
    assert "cleanup_n10  48x18" in mapspec.dump_maps()
