"""Check (b) of the north star on the GPU: sampled-statistics distributions of the CUDA path
(Philox sampling + fast integration) against the reference's own sampler (libc rand) integrated
by the CPU oracle, on the same synthetic force table: excitation counts (chi-square against the
zero-truncated Poisson law) and meeting-time histograms (two-sample KS)."""
from __future__ import annotations

import numpy as np
import pytest
from scipy import stats

from sphere_roaming_b200 import engine
from sphere_roaming_b200.types import make_config

pytestmark = pytest.mark.gpu
HE = 10000
P_MIN = 1e-3


@pytest.mark.parametrize("mode", ["fast", "fast_relaxed", "fast_closed"])
def test_histograms_agree_with_reference_distribution(port, table, mode):
    if engine.device_count() < 1:
        pytest.fail("needs a CUDA device")
    max_time, n_gpu, n_cpu = 1.0e5, 200_000, 3000
    conf = make_config(number=n_cpu, max_time=max_time)
    counts, init = port.sample(conf, HE, table, n_cpu, source="rand")     # the reference's sampler
    final, res = port.simulate(conf, HE, table, counts, init)              # the reference's integrator
    with engine.Engine(make_config(number=n_gpu, max_time=max_time), table, mode=mode) as eng:
        st = eng.run_stats(HE, n_gpu, first_index=10**6)
    assert st["trajectories"] == n_gpu
    # excitation count: exact zero-truncated Poisson pmf
    lam = port.lam(conf, HE)
    k = np.arange(1, 16)
    pmf = stats.poisson.pmf(k, lam) / (1.0 - np.exp(-lam))
    pmf[-1] += 1.0 - pmf.sum()
    obs = st["hist_n"][1:16].astype(np.float64)
    obs[-1] += st["hist_n"][16:].sum()
    assert stats.chisquare(obs, pmf * n_gpu).pvalue > P_MIN
    # per-particle meeting times: GPU histogram (4096 bins over max_time) vs CPU sample
    edges = np.linspace(0.0, max_time, 4097)
    met = final["success"] != 0
    cpu_hist = np.histogram(final["time"][met], bins=edges)[0]
    gpu_hist = st["hist_meet_time"].astype(np.float64)
    # coarsen to 64 bins and compare as a contingency table (KS on binned data loses nothing at this width)
    c = cpu_hist.reshape(64, 64).sum(1).astype(np.float64)
    g = gpu_hist.reshape(64, 64).sum(1)
    cdf_c, cdf_g = np.cumsum(c) / c.sum(), np.cumsum(g) / g.sum()
    d = np.abs(cdf_c - cdf_g).max()
    n_eff = c.sum() * g.sum() / (c.sum() + g.sum())
    p = stats.kstwobign.sf(d * np.sqrt(n_eff))
    assert p > P_MIN, (d, p)
    # fraction of particles paired by max_time
    frac_cpu = met.mean()
    frac_gpu = 1.0 - st["unmet_particles"] / st["particles"]
    assert abs(frac_cpu - frac_gpu) < 4.0 * np.sqrt(frac_cpu * (1 - frac_cpu) / met.size) + 0.005
    # trajectory exit times
    ce = np.histogram(res["time"], bins=edges)[0].reshape(64, 64).sum(1).astype(np.float64)
    ge = st["hist_traj_time"].astype(np.float64).reshape(64, 64).sum(1)
    d2 = np.abs(np.cumsum(ce) / ce.sum() - np.cumsum(ge) / ge.sum()).max()
    assert stats.kstwobign.sf(d2 * np.sqrt(ce.sum() * ge.sum() / (ce.sum() + ge.sum()))) > P_MIN
