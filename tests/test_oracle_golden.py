"""The CPU oracle (oracle/heroam_oracle.c) against golden vectors produced by the
reference's own object code (tests/golden/make_golden.py).  Bit-exact: the oracle is
compiled `-O2 -ffp-contract=off`, the goldens come from the reference built the same way.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

from conftest import assert_records_equal, load_golden
from sphere_roaming_b200.types import PARTICLE_DTYPE, Config, make_config

_fp = C.POINTER(C.c_float)


def fptr(a):
    return a.ctypes.data_as(_fp)


def test_layouts_match_reference():
    g = load_golden("layout.npz")
    assert tuple(g["sizes"]) == (PARTICLE_DTYPE.itemsize, 16, C.sizeof(Config))
    want = [PARTICLE_DTYPE.fields[n][1] for n in PARTICLE_DTYPE.names]
    assert list(g["particle_offsets"]) == want
    assert list(g["config_offsets"]) == [getattr(Config, n).offset for n, _ in Config._fields_]


def test_synthetic_table_is_reproducible(table):
    g = load_golden("layout.npz")
    assert len(table) == 1_596_000 and table.dtype == np.float32
    assert np.array_equal(table[:8], g["table_head"]) and np.array_equal(table[-8:], g["table_tail"])
    assert float(table.astype(np.float64).sum()) == float(g["table_sum"])
    # T-ref: -1e-6 (sqrt(d)/6)^-4 on d = 4.0 + 0.001 i  
    d = 4.0 + 0.001 * np.array([0, 32000, 1595999])
    assert np.allclose(table[[0, 32000, 1595999]], -1e-6 * (np.sqrt(d) / 6.0) ** -4, rtol=1e-6)


def test_vector_kats(port):
    g = load_golden("vector_kats.npz")
    L = port.lib
    for f in (L.orc_kat_norm, L.orc_kat_norm_sq, L.orc_kat_qnorm):
        f.restype = C.c_float
    a, b, s, q1, q2 = g["a"], g["b"], g["s"], g["q1"], g["q2"]
    n = len(a)
    out3, out4 = np.zeros(3, np.float32), np.zeros(4, np.float32)
    for i in range(n):
        ai, bi = np.ascontiguousarray(a[i]), np.ascontiguousarray(b[i])
        qi, qj = np.ascontiguousarray(q1[i]), np.ascontiguousarray(q2[i])
        L.orc_kat_cross(fptr(ai), fptr(bi), fptr(out3)); assert np.array_equal(out3, g["cross"][i])
        L.orc_kat_add(fptr(ai), fptr(bi), fptr(out3)); assert np.array_equal(out3, g["add"][i])
        L.orc_kat_scale(C.c_float(float(s[i])), fptr(ai), fptr(out3)); assert np.array_equal(out3, g["scal"][i])
        L.orc_kat_tangent(fptr(ai), fptr(out3)); assert np.array_equal(out3, g["perp"][i])
        assert np.float32(L.orc_kat_norm(fptr(ai))) == g["norm"][i]
        assert np.float32(L.orc_kat_norm_sq(fptr(ai))) == g["norm_sq"][i]
        L.orc_kat_qmul(fptr(qi), fptr(qj), fptr(out4)); assert np.array_equal(out4, g["qprod"][i])
        L.orc_kat_qinv(fptr(qi), fptr(out4)); assert np.array_equal(out4, g["qinv"][i])
        L.orc_kat_pole(fptr(qi), fptr(out3)); assert np.array_equal(out3, g["qpos"][i])
        assert np.float32(L.orc_kat_qnorm(fptr(qi))) == g["qnorm"][i]
        # find_accelration / update_orientation (file-static in the reference)
        L.orc_kat_accel(C.c_float(float(g["mass"])), C.c_float(float(g["radius"])), fptr(ai), fptr(bi), fptr(out3))
        assert np.array_equal(out3, g["acc"][i])
        quat = qi.copy()
        wi = np.ascontiguousarray(g["w"][i])
        L.orc_kat_orientation(C.c_float(float(g["radius"])), C.c_float(1.0), fptr(quat), fptr(wi), fptr(out3))
        assert np.array_equal(quat, g["uq"][i]) and np.array_equal(out3, g["upos"][i])


def test_sampler_streams(port):
    g = load_golden("sampler_streams.npz")
    L = port.lib
    n = len(g["poisson"])
    got = np.zeros(n, np.int32)
    L.orc_kat_poisson(int(g["poisson_seed"]), C.c_double(float(g["poisson_lambda"])), n, got.ctypes.data_as(C.POINTER(C.c_int)))
    assert np.array_equal(got, g["poisson"])
    sp = np.zeros(len(g["speeds"]), np.float32)
    L.orc_kat_speed(int(g["speed_seed"]), C.c_float(28.0), len(sp), fptr(sp))
    assert np.array_equal(sp, g["speeds"])
    qs = np.zeros_like(g["quats"])
    L.orc_kat_quat(int(g["quat_seed"]), len(qs), fptr(qs))
    assert np.array_equal(qs, g["quats"])


def test_force_cases(port, table):
    g = load_golden("force_cases.npz")
    conf = make_config()
    names = sorted(k[: -len("__in")] for k in g.files if k.endswith("__in"))
    assert len(names) >= 8
    port.lib.orc_find_success.restype = C.c_int
    for name in names:
        got = port.calculate_force(conf, table, g[f"{name}__in"])
        assert_records_equal(got, g[f"{name}__out"], name)
        assert int(port.lib.orc_find_success(len(got), got.ctypes.data)) == int(g[f"{name}__success"]), name
    # the behaviours the cases were built for
    assert list(g["triple_flag__out"]["success"]) == [1, 1, 1, 0, 0]  # Q3: three flags in one sweep
    assert g["frozen_neighbour__out"]["success"][0] == 0  # Q8: a frozen neighbour cannot be met
    assert g["last_index_force_mag__out"]["force_mag"][-1] == 0.0  # Q9


def test_sampler_matches_reference_rand_stream(port, table, golden_traj):
    g = golden_traj
    conf = make_config(number=len(g["counts"]), max_time=float(g["max_time"]))
    counts, init = port.sample(conf, int(g["he_number"]), table, len(g["counts"]), source="rand")
    assert np.array_equal(counts, g["counts"])
    assert_records_equal(init, g["init"], "initialize_particles")


@pytest.mark.parametrize("k", [1, 10, 100, 1000, 10000])
def test_state_after_k_steps(port, table, golden_traj, k):
    g = golden_traj
    conf = make_config(number=len(g["counts"]), max_time=float(k))
    got, res = port.simulate(conf, int(g["he_number"]), table, g["counts"], g["init"])
    assert_records_equal(got, g[f"step_{k}"], f"after {k} steps")
    assert res["steps"].max() <= k


def test_final_states(port, table, golden_traj):
    g = golden_traj
    conf = make_config(number=len(g["counts"]), max_time=float(g["max_time"]))
    got, res = port.simulate(conf, int(g["he_number"]), table, g["counts"], g["init"])
    assert_records_equal(got, g["final"], "final")
    assert set(np.unique(res["status"])) <= {0, 1}
    assert (res["status"] == 0).sum() > 10 and (res["status"] == 1).sum() > 10


def test_dense_droplet(port, table, golden_traj):
    g = golden_traj
    conf = make_config(number=len(g["dense_counts"]), max_time=float(g["dense_max_time"]),
                       intensity=float(g["dense_intensity"]))
    he = int(g["dense_he_number"])
    counts, init = port.sample(conf, he, table, len(g["dense_counts"]), source="rand")
    assert np.array_equal(counts, g["dense_counts"])
    assert_records_equal(init, g["dense_init"], "dense init")
    got, _ = port.simulate(conf, he, table, counts, init)
    assert_records_equal(got, g["dense_final"], "dense final")


def test_defined_exits(port, table):
    """Places where the reference's loop would never return get a defined exit."""
    conf = make_config(number=1, max_time=50.0)
    R = port.radius_sim(10000)
    p = np.zeros(3, dtype=PARTICLE_DTYPE)
    p["orientation"][:, 0] = 1.0
    # three particles on top of each other: all three flagged in one sweep, 3 != 2 (Q4)
    p["position"] = [(0, 0, R), (0.5, 0, R), (0, 0.5, R)]
    got, res = port.simulate(conf, 10000, table, np.array([3], np.uint32), p)
    assert list(got["success"]) == [1, 1, 1] and res["status"][0] == 2 and res["steps"][0] == 1
    # n = 1: exactly one step, then the odd rule succeeds (Q2)
    one = np.zeros(1, dtype=PARTICLE_DTYPE)
    one["orientation"][0] = (1, 0, 0, 0)
    one["position"][0] = (0, 0, R)
    one["ang_velocity"][0] = (1e-5, 0, 0)
    got, res = port.simulate(conf, 10000, table, np.array([1], np.uint32), one)
    assert res["steps"][0] == 1 and res["status"][0] == 0 and got["time"][0] == 1.0
