"""Parity at the horizon the benchmark runs at (BASELINE.json configs, full max_time), against
fixtures produced by the REFERENCE'S OWN OBJECT CODE (oracle/_ref strict build; the generating
script is tests/golden/make_golden_horizon.py, run where /root/reference exists).

  C2  repo config.conf physics, 4096 trajectories from reference-exported initial states
      (initialize_particles, libc rand) run to max_time = 1e7 fs           -> horizon_c2.npz
  C3  N_He = 1e7, lambda = 49.2, 256 trajectories, max_time = 1e5 fs       -> horizon_c3.npz
  C4  the lambda = 0.49 and lambda = 31.5 corners of the laser sweep, 1e7  -> horizon_c4lo/hi.npz

EXACT mode : every final record bit-identical (float ==), through heroam_gpu_simulate_batch --
             the reference-layout call INTEGRATION.md tells a maintainer to use.
FAST modes : stated tolerance AT THIS HORIZON versus the strict reference, from identical
             initial states (tolerance table below; the trajectories outside it are listed in the
             assertion message).  The reference's own release build (-O3 -ffast-math) is held to
             the same table in tests/test_oracle_horizon.py.
Statistics : two-sample KS of exit times and per-particle meeting times, GPU Philox run at
             max_time = 1e7 against the reference sample of horizon_c2.npz.
"""
from __future__ import annotations

import numpy as np
import pytest
from scipy import stats

from conftest import assert_records_equal
from horizon_common import HORIZON_TOLERANCE, compare_with_tolerance, config_of, exit_times, load
from sphere_roaming_b200 import engine
from sphere_roaming_b200.types import ParticlesC, make_config, offsets_of

pytestmark = pytest.mark.gpu

def simulate_batch(eng, he, counts, init):
    """heroam_gpu_simulate_batch: the pointer-per-trajectory layout of the reference (main.c:55,
    simulation.c:310), in place.  Returns (final records, results)."""
    n = len(counts)
    off = offsets_of(counts)
    bufs = [np.ascontiguousarray(init[off[i]:off[i + 1]].copy()) for i in range(n)]
    arr = (ParticlesC * n)()
    for i in range(n):
        arr[i].no, arr[i].index, arr[i].particle = int(counts[i]), i, bufs[i].ctypes.data
    res = np.zeros(n, dtype=engine.RESULT_DTYPE)
    rc = eng.lib.heroam_gpu_simulate_batch(eng.handle, he, arr, n, res.ctypes.data)
    assert rc == 0, eng.lib.heroam_gpu_last_error()
    return np.concatenate(bufs), res



@pytest.fixture(scope="module")
def need_gpu():
    if engine.device_count() < 1:
        pytest.fail("GPU tests need a CUDA device: the engine has no CPU fallback")


# ------------------------------------------------------------------------------------------------
# C2 at max_time = 1e7
# ------------------------------------------------------------------------------------------------
def test_c2_exact_bitwise_at_1e7(need_gpu, table):
    g = load("c2")
    counts, init, want = g["counts"], g["init"], g["final"]
    with engine.Engine(config_of(g, len(counts)), table, mode="exact") as eng:
        got, res = simulate_batch(eng, int(g["he_number"]), counts, init)
        info = eng.run_info()
    assert_records_equal(got, want, "C2, 4096 reference-exported trajectories at max_time = 1e7")
    assert np.array_equal(res["time"], exit_times(counts, want))
    # particle-steps = sum of the particle clocks / dt
    assert info["particle_steps"] == int(want["time"].astype(np.float64).sum())
    assert (res["status"] == engine.DONE_MAXTIME).sum() > 20, "the fixture should contain trajectories that run to max_time"


@pytest.mark.parametrize("mode", ["fast", "fast_relaxed", "fast_closed"])
def test_c2_fast_within_stated_tolerance_at_1e7(need_gpu, table, mode):
    g = load("c2")
    counts, init, want = g["counts"], g["init"], g["final"]
    with engine.Engine(config_of(g, len(counts)), table, mode=mode) as eng:
        got, res = simulate_batch(eng, int(g["he_number"]), counts, init)
    matched = compare_with_tolerance(counts, got, want, HORIZON_TOLERANCE[mode], f"C2 {mode} at 1e7")
    # outcomes as populations: the number of trajectories that run to max_time agrees to a few
    max_time = float(g["max_time"])
    n_cap_got = int((exit_times(counts, got) >= max_time).sum())
    n_cap_want = int((exit_times(counts, want) >= max_time).sum())
    assert abs(n_cap_got - n_cap_want) <= max(3, int(0.1 * n_cap_want)), (n_cap_got, n_cap_want)
    assert matched.sum() > 0


def binned_ks(sample, hist, max_time):
    """Two-sample KS between a sample and a 4096-bin histogram over [0, max_time] (bins coarsened to
    256: at 4e3-2e4 sample points nothing is lost)."""
    edges = np.linspace(0.0, max_time, 4097)
    a = np.histogram(sample, bins=edges)[0].astype(np.float64).reshape(256, 16).sum(1)  # the last bin includes max_time
    b = hist.astype(np.float64).reshape(256, 16).sum(1)
    d = np.abs(np.cumsum(a) / a.sum() - np.cumsum(b) / b.sum()).max()
    n_eff = a.sum() * b.sum() / (a.sum() + b.sum())
    return d, stats.kstwobign.sf(d * np.sqrt(n_eff))


@pytest.mark.parametrize("mode", ["fast", "fast_relaxed", "fast_closed"])
def test_c2_statistics_at_1e7(need_gpu, table, mode):
    """Check (b) at the benchmark's own horizon: Philox-sampled GPU run (200 000 trajectories to
    max_time = 1e7) against the reference sample (libc-rand sampler + strict integrator, 4096
    trajectories): exit times, per-particle meeting times, outcome fractions."""
    g = load("c2")
    counts, want = g["counts"], g["final"]
    max_time = float(g["max_time"])
    n_gpu = 200_000
    with engine.Engine(make_config(number=n_gpu, max_time=max_time), table, mode=mode) as eng:
        st = eng.run_stats(int(g["he_number"]), n_gpu, first_index=5 * 10**6)
    assert st["trajectories"] == n_gpu
    t_exit = exit_times(counts, want)
    d1, p1 = binned_ks(t_exit, st["hist_traj_time"], max_time)
    met = want["success"] != 0
    d2, p2 = binned_ks(want["time"][met], st["hist_meet_time"], max_time)
    print(f"KS at 1e7 ({mode}): exit times D = {d1:.4f} p = {p1:.3f}; meeting times D = {d2:.4f} p = {p2:.3f}")
    assert p1 > 1e-3 and p2 > 1e-3, (d1, p1, d2, p2)
    # fraction of trajectories that reach max_time, and of particles left unpaired
    f_cap_ref = (t_exit >= max_time).mean()
    f_cap_gpu = st["done"][engine.DONE_MAXTIME] / n_gpu
    assert abs(f_cap_ref - f_cap_gpu) < 4.0 * np.sqrt(f_cap_ref * (1 - f_cap_ref) / len(counts)) + 1e-3, (f_cap_ref, f_cap_gpu)
    f_unmet_ref = 1.0 - met.mean()
    f_unmet_gpu = st["unmet_particles"] / st["particles"]
    assert abs(f_unmet_ref - f_unmet_gpu) < 4.0 * np.sqrt(f_unmet_ref * (1 - f_unmet_ref) / met.size) + 2e-3


# ------------------------------------------------------------------------------------------------
# C3 at spec and the C4 corners
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c3", "c4lo", "c4hi"])
def test_other_configs_exact_bitwise(need_gpu, table, name):
    g = load(name)
    counts, init, want = g["counts"], g["init"], g["final"]
    with engine.Engine(config_of(g, len(counts)), table, mode="exact") as eng:
        got, res = simulate_batch(eng, int(g["he_number"]), counts, init)
    assert_records_equal(got, want, f"{name} at its full horizon")
    assert np.array_equal(res["time"], exit_times(counts, want))


@pytest.mark.parametrize("name", ["c3", "c4lo", "c4hi"])
def test_other_configs_fast_within_tolerance(need_gpu, table, name):
    g = load(name)
    counts, init, want = g["counts"], g["init"], g["final"]
    with engine.Engine(config_of(g, len(counts)), table, mode="fast") as eng:
        got, res = simulate_batch(eng, int(g["he_number"]), counts, init)
    if name == "c3":
        # 1e5 steps on R = 478 A, few meetings (2.5 % of the particles): every particle with the same flag
        # and the same clock is compared
        flags_differ = (got["success"] != want["success"]).mean()
        same = (got["time"] == want["time"]) & (got["success"] == want["success"])
        dr = np.abs(got["position"].astype(np.float64) - want["position"]).max(axis=1)[same]
        line = (f"C3 fast: flags differ on {flags_differ:.2%} of the particles, {same.mean():.2%} share flag and clock; their position "
                f"error: median {np.median(dr):.3g} A, 99% {np.quantile(dr, 0.99):.3g} A (R = 478 A)")
        print(line)
        assert flags_differ < 5e-3 and same.mean() > 0.98 and np.median(dr) < 2e-3 and np.quantile(dr, 0.99) < 5e-2, line
    else:
        compare_with_tolerance(counts, got, want, HORIZON_TOLERANCE["fast_dense" if name == "c4hi" else "fast"], f"{name} fast")


# ------------------------------------------------------------------------------------------------
# the pin itself, in the GPU record: the repo's restatement == the reference's object code
# ------------------------------------------------------------------------------------------------
def test_port_oracle_equals_reference_object_code_on_this_box(port, table):
    """oracle/_ref (the reference's simulation.c + vector.c compiled unmodified, shipped prebuilt)
    loaded directly on the GPU box: the C restatement the GPU tests compare with must equal it bit
    for bit, sampler and integrator."""
    from oracle.binding import RefOracle, ref_available

    if not ref_available("strict"):
        pytest.fail("oracle/_ref/libheroam_ref_strict.so did not travel to this box")
    ref = RefOracle("strict")
    assert ref.sizes() == (84, 16, 152)
    conf = make_config(number=192, max_time=1.0e5)
    counts, init = ref.initialize(conf, 10000, table)
    p_counts, p_init = port.sample(conf, 10000, table, 192, source="rand")
    assert np.array_equal(counts, p_counts) and init.tobytes() == p_init.tobytes()
    want = ref.simulate(conf, 10000, table, counts, init)
    got, _ = port.simulate(conf, 10000, table, counts, init)
    assert_records_equal(got, want, "port vs reference object code")
