"""Parity tests proper: the CUDA path, called through the C ABI (libheroam_gpu.so),
against the CPU oracle and the committed golden vectors of the reference.

EXACT mode  : bit-for-bit (float ==) with the reference compiled -O2 -ffp-contract=off.
FAST mode   : stated tolerance, from identical initial states, versus the strict oracle
              (the reference's own -ffast-math build shows the same spread,
              see test_oracle_vs_ref.py::test_release_build_agrees_within_stated_tolerance):
                * per-trajectory exit time within max(50 steps, 1e-4 t) for >= 99 %
                * identical success pattern for >= 99 % of trajectories
                * positions of trajectories with identical exit step within 5e-3 A (median < 2e-4 A)
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

from conftest import assert_records_equal
from sphere_roaming_b200 import engine
from sphere_roaming_b200.types import PARTICLE_DTYPE, ParticlesC, make_config, offsets_of

pytestmark = pytest.mark.gpu

HE = 10000


@pytest.fixture(scope="module")
def need_gpu():
    if engine.device_count() < 1:
        pytest.fail("GPU tests need a CUDA device: the engine has no CPU fallback")


def test_exact_mode_matches_reference_goldens(need_gpu, table, golden_traj):
    g = golden_traj
    he = int(g["he_number"])
    for k in (1, 10, 100, 1000, 10000):
        with engine.Engine(make_config(number=len(g["counts"]), max_time=float(k)), table, mode="exact") as eng:
            got, res = eng.simulate(he, g["counts"], g["init"])
        assert_records_equal(got, g[f"step_{k}"], f"state after {k} steps")
    with engine.Engine(make_config(number=len(g["counts"]), max_time=float(g["max_time"])), table, mode="exact") as eng:
        got, res = eng.simulate(he, g["counts"], g["init"])
    assert_records_equal(got, g["final"], "final state")


def test_exact_mode_dense_droplet_golden(need_gpu, table, golden_traj):
    g = golden_traj
    conf = make_config(number=len(g["dense_counts"]), max_time=float(g["dense_max_time"]),
                       intensity=float(g["dense_intensity"]))
    with engine.Engine(conf, table, mode="exact") as eng:
        got, _ = eng.simulate(int(g["dense_he_number"]), g["dense_counts"], g["dense_init"])
    assert_records_equal(got, g["dense_final"], "dense droplet final state")


@pytest.mark.parametrize("segment", [0, 257])
def test_exact_mode_matches_oracle_at_scale(need_gpu, port, table, segment):
    """4096 Philox-sampled trajectories (config C2's physics; the per-trajectory
    cross-check block), 1e5 steps, every record, outcome and work counter."""
    n, max_time = 4096, 1.0e5
    conf = make_config(number=n, max_time=max_time)
    counts, init = port.sample(conf, HE, table, n, source="philox")
    want, wres, wctr = port.simulate(conf, HE, table, counts, init, want_counters=True)
    with engine.Engine(conf, table, mode="exact", segment_steps=segment) as eng:
        got, res = eng.simulate(HE, counts, init)
        info = eng.run_info()
    assert_records_equal(got, want, "final records")
    for f in ("steps", "status", "n_flagged", "time"):
        assert np.array_equal(res[f], wres[f]), f
    assert info["particle_steps"] == wctr["particle_steps"]
    assert info["pair_steps"] == wctr["pair_steps"]
    assert info["inrange_steps"] == wctr["inrange_steps"]
    assert counts.max() > 8, "the batch should exercise the wide kernel too"


@pytest.mark.parametrize("intensity,max_time", [(1.5, 4000.0), (5.0, 1500.0)])
def test_wide_kernels_match_oracle(need_gpu, port, table, intensity, max_time):
    """Tens of excitations per trajectory (config C3's regime, on a smaller droplet so
    that pairs do meet): the warp-per-trajectory kernels (1, 2 and 4 particles per lane)
    against the oracle, and against the slow record-resident kernel on a short horizon."""
    n, he = 96, 40000
    conf = make_config(number=n, max_time=max_time, intensity=intensity, helium_number=he)
    counts, init = port.sample(conf, he, table, n, source="philox")
    assert counts.max() <= 128 and counts.max() > (64 if intensity > 2 else 32)
    want, wres = port.simulate(conf, he, table, counts, init)
    # 0: hybrid (9..12) + shared-memory (13..16) + warp kernels; 2: warp kernels for all > 8; 3: shared-memory for 9..16
    for wide in (0, 2, 3):
        with engine.Engine(conf, table, mode="exact", segment_steps=500, wide_kernel=wide) as eng:
            got, res = eng.simulate(he, counts, init)
        assert_records_equal(got, want, f"wide kernels (wide_kernel={wide})")
        assert np.array_equal(res["steps"], wres["steps"]) and np.array_equal(res["n_flagged"], wres["n_flagged"])
    short = make_config(number=n, max_time=200.0, intensity=intensity, helium_number=he)
    want_s, _ = port.simulate(short, he, table, counts, init)
    with engine.Engine(short, table, mode="exact", wide_kernel=1) as eng:
        got_s, _ = eng.simulate(he, counts, init)
    assert_records_equal(got_s, want_s, "record-resident kernel")


def test_sparse_large_droplet_partner_lists(need_gpu, port, table):
    """Config C3's regime proper: N_He = 1e7 (R = 478 A), ~49 excitations, almost every pair out of
    range.  The warp kernels then work from partner lists (pairs that provably cannot interact during
    the segment are skipped); results must still be bit-identical to the oracle, which evaluates all."""
    n, he, max_time = 24, 10**7, 6000.0
    conf = make_config(number=n, max_time=max_time, intensity=0.0125, helium_number=he)
    counts, init = port.sample(conf, he, table, n, source="philox")
    assert 30 < counts.mean() < 70
    want, wres, wctr = port.simulate(conf, he, table, counts, init, want_counters=True)
    with engine.Engine(conf, table, mode="exact", segment_steps=4096) as eng:
        got, res = eng.simulate(he, counts, init)
        info = eng.run_info()
    assert_records_equal(got, want, "sparse large droplet")
    assert np.array_equal(res["steps"], wres["steps"])
    assert info["pair_steps"] == wctr["pair_steps"] and info["inrange_steps"] == wctr["inrange_steps"]
    assert info["pair_skipped"] > 0.9 * info["pair_steps"], "the partner lists should skip almost every pair here"


def test_large_table_gathers_from_hbm(need_gpu, port):
    """The stress table of SURVEY.md 8d ("T-large"): same r^2 range on a 6e-5 A^2 grid, 26.6 M
    entries = 106 MB in memory, so the gathers go past L1 to L2/HBM.  Same bits as the oracle."""
    from sphere_roaming_b200.forcetable import synthetic_table

    length, grid = 26_600_000, 6.0e-5
    big = synthetic_table(length=length, grid=grid)
    n, max_time = 512, 2.0e4
    conf = make_config(number=n, max_time=max_time, force_grid=grid, force_grid_length=length, velocity=280.0)
    counts, init = port.sample(conf, HE, big, n, source="philox")
    want, wres, wctr = port.simulate(conf, HE, big, counts, init, want_counters=True)
    with engine.Engine(conf, big, mode="exact") as eng:
        got, res = eng.simulate(HE, counts, init)
        info = eng.run_info()
    assert_records_equal(got, want, "T-large")
    assert np.array_equal(res["steps"], wres["steps"]) and np.array_equal(res["status"], wres["status"])
    assert info["inrange_steps"] == wctr["inrange_steps"] > 0
    assert (res["status"] == engine.DONE_SUCCESS).sum() > 32, "the batch should contain meetings"


#: stated position tolerance [A] after 2e5 steps, trajectories with identical exit step: (median, 99 %)
POSITION_TOLERANCE = {
    "fast": (2e-4, 5e-3),          # measured 2.4e-5 / 6.0e-4: the roundings mirror the reference's step for step
    "fast_relaxed": (5e-3, 1e-1),  # measured 1.6e-3 / 2.3e-2: renormalisation every 8th step, independent rounding walk
    "fast_closed": (5e-3, 1e-1),   # free flights as single rotations: the same class (the reference's own rounding walk)
}


@pytest.mark.parametrize("mode", ["fast", "fast_relaxed", "fast_closed"])
def test_fast_mode_within_stated_tolerance(need_gpu, port, table, mode):
    n, max_time = 4096, 2.0e5
    conf = make_config(number=n, max_time=max_time)
    counts, init = port.sample(conf, HE, table, n, source="philox")
    want, wres = port.simulate(conf, HE, table, counts, init)
    with engine.Engine(conf, table, mode=mode) as eng:
        got, res = eng.simulate(HE, counts, init)
    off = offsets_of(counts)
    dt = np.abs(res["time"].astype(np.float64) - wres["time"])
    ok_time = dt <= np.maximum(50.0, 1e-4 * wres["time"])
    assert ok_time.mean() >= 0.99, f"exit time off for {(~ok_time).sum()} of {n}"
    same_flags = np.array([np.array_equal(got["success"][off[i]:off[i + 1]], want["success"][off[i]:off[i + 1]])
                           for i in range(n)])
    assert same_flags.mean() >= 0.99
    assert np.array_equal(res["status"] == 1, wres["status"] == 1) or ((res["status"] == 1) != (wres["status"] == 1)).mean() < 0.01
    same_step = np.repeat(res["steps"] == wres["steps"], counts)
    dr = np.abs(got["position"].astype(np.float64) - want["position"]).max(axis=1)[same_step & (got["time"] == want["time"])]
    tol_median, tol_99 = POSITION_TOLERANCE[mode]
    assert np.median(dr) < tol_median and np.quantile(dr, 0.99) < tol_99, (np.median(dr), np.quantile(dr, 0.99), dr.max())
    # radius is preserved by the re-projection
    R = port.radius_sim(HE)
    assert np.allclose(np.linalg.norm(got["position"].astype(np.float64), axis=1), R, rtol=3e-6)


@pytest.mark.parametrize("mode", ["fast", "fast_relaxed", "exact"])
def test_free_flight_and_splitting_change_nothing(need_gpu, table, mode):
    """The shortcuts (whole-trajectory free flight, loners split off, partner lists) against the
    same kernels with the shortcuts switched off: every record identical, in every mode -- the
    fast modes have no bitwise oracle, but they must at least not depend on the schedule."""
    n, max_time = 8192, 6.0e4
    conf = make_config(number=n, max_time=max_time)
    with engine.Engine(conf, table, mode=mode) as eng:
        counts, init = eng.sample(HE, n)
        with_flight, res_a = eng.simulate(HE, counts, init)
        info = eng.run_info()
    with engine.Engine(conf, table, mode=mode, no_flight=1, segment_steps=1000) as eng:
        without, res_b = eng.simulate(HE, counts, init)
        info_b = eng.run_info()
    assert info["free_steps"] > 0.1 * info["particle_steps"] and info_b["free_steps"] == 0
    assert np.array_equal(res_a, res_b)
    assert with_flight.tobytes() == without.tobytes()
    assert info["particle_steps"] == info_b["particle_steps"] and info["pair_steps"] == info_b["pair_steps"]
    # segments cut very differently
    for kw in (dict(segment_steps=3000), dict(segment_steps=100000)):
        with engine.Engine(conf, table, mode=mode, **kw) as eng:
            other, res_c = eng.simulate(HE, counts, init)
            info_c = eng.run_info()
        assert np.array_equal(res_a, res_c), kw
        assert with_flight.tobytes() == other.tobytes(), kw
        assert info["particle_steps"] == info_c["particle_steps"] and info["pair_steps"] == info_c["pair_steps"], kw


def test_gpu_sampler_matches_oracle_twin(need_gpu, port, table):
    """The Philox sampling kernel against its CPU twin: counts and geometry bit-exact,
    speeds within a few float ulps (device vs glibc log/sin/cos in double)."""
    n = 20000
    conf = make_config(number=n)
    counts, init = port.sample(conf, HE, table, n, source="philox", first_index=12345)
    with engine.Engine(conf, table) as eng:
        g_counts, g_init = eng.sample(HE, n, first_index=12345)
    assert np.array_equal(g_counts, counts)
    for f in ("index", "success", "time", "orientation", "position"):
        assert np.array_equal(g_init[f], init[f]), f
    for f in ("velocity_sq", "velocity", "ang_velocity", "force", "force_mag"):
        a, b = g_init[f].astype(np.float64), init[f].astype(np.float64)
        rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-30)
        assert rel.max() < 1e-6, (f, rel.max())
        assert (a != b).mean() < 0.02, f


def test_simulate_batch_reference_layout(need_gpu, port, table):
    """heroam_gpu_simulate_batch: the pointer-per-trajectory layout of the reference's
    Particles[] (main.c:55, simulation.c:310), in place."""
    n, max_time = 64, 5000.0
    conf = make_config(number=n, max_time=max_time)
    counts, init = port.sample(conf, HE, table, n, source="philox")
    want, wres = port.simulate(conf, HE, table, counts, init)
    off = offsets_of(counts)
    bufs = [np.ascontiguousarray(init[off[i]:off[i + 1]].copy()) for i in range(n)]
    arr = (ParticlesC * n)()
    for i in range(n):
        arr[i].no, arr[i].index, arr[i].particle = int(counts[i]), i, bufs[i].ctypes.data
    res = np.zeros(n, dtype=engine.RESULT_DTYPE)
    with engine.Engine(conf, table, mode="exact") as eng:
        eng.reserve(2 * n, 2 * len(init))  # heroam_gpu_reserve: buffers sized ahead, results unchanged
        rc = eng.lib.heroam_gpu_simulate_batch(eng.handle, HE, arr, n, res.ctypes.data)
        assert rc == 0, eng.lib.heroam_gpu_last_error()
        with pytest.raises(engine.HeroamError):
            eng.reserve(1, 2**32)
    assert_records_equal(np.concatenate(bufs), want, "simulate_batch")
    assert np.array_equal(res["steps"], wres["steps"])


def test_sample_batch_ownership(need_gpu, table):
    n = 32
    conf = make_config(number=n)
    with engine.Engine(conf, table) as eng:
        counts, flat = eng.sample(HE, n, first_index=5)
        arr = (ParticlesC * n)()
        assert eng.lib.heroam_gpu_sample_batch(eng.handle, HE, 5, n, arr) == 0
        off = offsets_of(counts)
        for i in range(n):
            assert arr[i].no == counts[i] and arr[i].index == 5 + i
            rec = np.frombuffer((C.c_char * (84 * arr[i].no)).from_address(arr[i].particle), dtype=PARTICLE_DTYPE)
            assert rec.tobytes() == flat[off[i]:off[i + 1]].tobytes()
            eng.lib.heroam_gpu_free_particles(C.byref(arr[i]))
            assert arr[i].particle is None


def test_edge_cases(need_gpu, port, table):
    R = port.radius_sim(HE)
    conf = make_config(number=1, max_time=50.0)
    with engine.Engine(conf, table, mode="exact") as eng:
        # empty batch
        got, res = eng.simulate(HE, np.zeros(0, np.uint32), np.zeros(0, PARTICLE_DTYPE))
        assert len(got) == 0 and len(res) == 0
        # ragged: n = 1 (one step, Q2), an all-frozen pair, a 3-way deadlock (Q4), a far pair
        p = np.zeros(8, dtype=PARTICLE_DTYPE)
        p["orientation"][:, 0] = 1.0
        p["position"] = [(0, 0, R), (0, 0, R), (1, 0, R), (0, 0, R), (0.5, 0, R), (0, 0.5, R), (0, 0, R), (0, 0, -R)]
        p["orientation"][7] = (0, 1, 0, 0)  # half turn about x: this particle really sits at the south pole
        p["ang_velocity"][0] = (1e-5, 0, 0)
        p["success"][1:3] = 1
        p["force"][1] = (5, 5, 5)
        counts = np.array([1, 2, 3, 2], np.uint32)
        want, wres = port.simulate(conf, HE, table, counts, p)
        got, res = eng.simulate(HE, counts, p)
        assert_records_equal(got, want, "edge batch")
        for f in ("steps", "status", "n_flagged", "time"):
            assert np.array_equal(res[f], wres[f]), f
        assert list(res["status"]) == [0, 0, 2, 1] and list(res["steps"]) == [1, 1, 1, 50]
        assert np.all(got["force"][1] == 0)  # frozen particles get the zero force of the first step
    # max_time <= 0: the loop body never runs, records come back untouched (time zeroed)
    with engine.Engine(make_config(number=1, max_time=0.0), table) as eng:
        got, res = eng.simulate(HE, counts, p)
        assert res["steps"].max() == 0 and np.array_equal(got["position"], p["position"]) and np.all(got["force"][1] == 5)
    # a clock that cannot reach max_time is refused (the reference would never return, Q10)
    with pytest.raises(engine.HeroamError, match="clock|never end"):
        with engine.Engine(make_config(number=1, max_time=1.0e9, dt=1.0), table) as eng:
            eng.simulate(HE, counts, p)
    # more than 128 moving particles: refused, not silently mis-integrated
    big = np.zeros(130, dtype=PARTICLE_DTYPE)
    rng = np.random.default_rng(1)
    v = rng.standard_normal((130, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
    big["position"] = (v * port.radius_sim(10**7)).astype(np.float32)
    big["orientation"][:, 0] = 1.0
    with pytest.raises(engine.HeroamError, match="moving particles"):
        with engine.Engine(make_config(number=1, max_time=10.0), table) as eng:
            eng.simulate(10**7, np.array([130], np.uint32), big)


def test_step_cap_and_resume_free_determinism(need_gpu, port, table):
    n = 512
    conf = make_config(number=n, max_time=1.0e5)
    counts, init = port.sample(conf, HE, table, n, source="philox")
    want, wres = port.simulate(conf, HE, table, counts, init, max_steps=3000)
    with engine.Engine(conf, table, mode="exact", max_steps=3000) as eng:
        got, res = eng.simulate(HE, counts, init)
        again, _ = eng.simulate(HE, counts, init)
    assert_records_equal(got, want, "step cap")
    assert got.tobytes() == again.tobytes()
    assert np.array_equal(res["status"], wres["status"]) and (res["status"] == 4).any()

@pytest.mark.parametrize("pool_slots", [0, 700])
def test_run_stats_streaming_pool_equals_batch_path(need_gpu, port, table, pool_slots):
    """heroam_gpu_run_stats (streaming pool with retire-and-refill, sampling fused in) must
    give exactly the statistics of sample + simulate on the same index block -- in exact
    mode bit for bit against the CPU oracle, whatever the pool size."""
    from sphere_roaming_b200.sharding import stats_from_records

    n, max_time, first = 5000, 3.0e4, 777
    conf = make_config(number=n, max_time=max_time)
    counts, init = port.sample(conf, HE, table, n, source="philox", first_index=first)
    final, res, ctr = port.simulate(conf, HE, table, counts, init, want_counters=True)
    want = stats_from_records(counts, final, res, max_time)
    with engine.Engine(conf, table, mode="exact", pool_slots=pool_slots, segment_steps=1000) as eng:
        got = eng.run_stats(HE, n, first_index=first)
    for key in ("hist_n", "hist_traj_time", "hist_meet_time", "done"):
        assert np.array_equal(got[key], want[key]), key
    assert got["trajectories"] == n and got["particles"] == int(counts.sum())
    assert got["unmet_particles"] == want["unmet_particles"]
    assert abs(got["sum_traj_time"] - want["sum_traj_time"]) < 1e-9 * want["sum_traj_time"]
    assert got["info"]["particle_steps"] == ctr["particle_steps"]
    assert got["info"]["pair_steps"] == ctr["pair_steps"] and got["info"]["inrange_steps"] == ctr["inrange_steps"]


def test_full_size_properties(need_gpu, port, table):
    """Size-independent properties on a large batch (a slice of config C2/C5: 2^18
    trajectories, fast mode): every particle stays on the sphere, clocks never exceed the
    horizon, flags come in meeting groups, sharding does not change the histograms."""
    n, max_time = 1 << 18, 2.0e4
    conf = make_config(number=n, max_time=max_time)
    with engine.Engine(conf, table, mode="fast") as eng:
        counts, init = eng.sample(HE, n)
        final, res = eng.simulate(HE, counts, init)
        whole = eng.run_stats(HE, n)
        a = eng.run_stats(HE, n // 2)
        b = eng.run_stats(HE, n - n // 2, first_index=n // 2)
    R = port.radius_sim(HE)
    assert np.allclose(np.linalg.norm(final["position"].astype(np.float64), axis=1), R, rtol=3e-6)
    assert final["time"].max() <= max_time and res["time"].max() <= max_time
    off = offsets_of(counts)
    flagged = np.add.reduceat(final["success"], off[:-1])
    assert np.array_equal(flagged, res["n_flagged"])
    ok = res["status"] == 0
    assert np.all(np.where(counts[ok] % 2 == 0, flagged[ok] == counts[ok], flagged[ok] == counts[ok] - 1))
    assert np.all(final["time"][final["success"] == 0] >= 0)
    # frozen particles carry zero force (Q8)
    assert np.all(final["force"][final["success"] != 0] == 0)
    for key in ("hist_n", "hist_traj_time", "hist_meet_time", "done"):
        assert np.array_equal(whole[key], a[key] + b[key]), key
    assert whole["trajectories"] == n
    # the relaxed mode keys its renormalisation to the absolute step, so it too is independent of
    # how the index space is cut and of what else shares the pool
    with engine.Engine(conf, table, mode="fast_relaxed") as eng:
        whole_r = eng.run_stats(HE, n)
        parts_r = [eng.run_stats(HE, n // 4, first_index=k * (n // 4)) for k in range(4)]
        batch_r, _ = eng.simulate(HE, counts[:4096], init[:off[4096]])
        batch_again, _ = eng.simulate(HE, counts[1000:4096], init[off[1000]:off[4096]])
    for key in ("hist_n", "hist_traj_time", "hist_meet_time", "done"):
        assert np.array_equal(whole_r[key], sum(p[key] for p in parts_r)), key
    assert batch_r[off[1000]:].tobytes() == batch_again.tobytes(), "a trajectory's result must not depend on its batch"
    assert np.array_equal(whole["hist_n"][:64], np.bincount(counts, minlength=64)[:64])
