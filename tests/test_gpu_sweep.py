"""Laser sweep in one launch (BASELINE config 4): every point of
heroam_gpu_run_sweep must equal that point run on its own through heroam_gpu_run_stats."""
from __future__ import annotations

import numpy as np
import pytest

from sphere_roaming_b200 import engine
from sphere_roaming_b200.sharding import shard_sweep
from sphere_roaming_b200.types import make_config

HE = 10000
POINTS = [(1.0, 100.0, 3000, 0), (0.25, 50.0, 2000, 100), (4.0, 200.0, 600, 7), (2.0, 100.0, 2500, 12345678901)]


def assert_same_stats(a, b, what):
    for k in ("trajectories", "particles", "unmet_particles"):
        assert a[k] == b[k], f"{what}: {k}"
    for k in ("hist_n", "hist_traj_time", "hist_meet_time", "done"):
        assert np.array_equal(a[k], b[k]), f"{what}: {k}"
    assert abs(a["sum_traj_time"] - b["sum_traj_time"]) <= 1e-9 * max(1.0, b["sum_traj_time"]), what


@pytest.mark.gpu
@pytest.mark.parametrize("mode,pool_slots", [("exact", 0), ("exact", 1024), ("fast", 0)])
def test_sweep_equals_separate_runs(table, mode, pool_slots):
    max_time = 2.0e4
    with engine.Engine(make_config(max_time=max_time), table, mode=mode, pool_slots=pool_slots) as eng:
        swept = eng.run_sweep(HE, POINTS)
    assert len(swept) == len(POINTS)
    for (intensity, width, count, first), got in zip(POINTS, swept):
        conf = make_config(max_time=max_time, intensity=intensity, pulse_width=width)
        with engine.Engine(conf, table, mode=mode) as eng:
            want = eng.run_stats(HE, count, first_index=first)
        assert_same_stats(got, want, f"point I={intensity} tau={width}")
    # lambda = 31.5 at the strongest point: the wide kernels took part
    assert int(np.nonzero(swept[2]["hist_n"])[0].max()) > 32
    assert swept[0]["info"]["trajectories"] == sum(p[2] for p in POINTS)


@pytest.mark.gpu
def test_sweep_is_independent_of_sharding(table):
    """Two shards of every point, run as two sweeps, add up to the unsharded sweep."""
    max_time = 2.0e4
    with engine.Engine(make_config(max_time=max_time), table, mode="fast") as eng:
        whole = eng.run_sweep(HE, POINTS)
        parts = [eng.run_sweep(HE, shard_sweep(POINTS, r, 2)) for r in range(2)]
    for p, w in enumerate(whole):
        for k in ("hist_n", "hist_traj_time", "hist_meet_time", "done"):
            assert np.array_equal(parts[0][p][k] + parts[1][p][k], w[k]), (p, k)
        assert parts[0][p]["trajectories"] + parts[1][p]["trajectories"] == w["trajectories"]


@pytest.mark.gpu
def test_sweep_argument_checks(table):
    with engine.Engine(make_config(), table) as eng:
        with pytest.raises(engine.HeroamError, match="points"):
            eng.run_sweep(HE, [(1.0, 100.0, 1, 0)] * 33)
        with pytest.raises(engine.HeroamError, match="excitation number"):
            eng.run_sweep(HE, [(0.0, 100.0, 10, 0)])
        assert eng.run_sweep(HE, [(1.0, 100.0, 0, 0)])[0]["trajectories"] == 0


def test_shard_sweep_partitions_every_point():
    for world in (1, 2, 3, 8):
        shards = [shard_sweep(POINTS, r, world) for r in range(world)]
        for p, (intensity, width, count, first) in enumerate(POINTS):
            at = first
            for s in shards:
                assert s[p][:2] == (intensity, width)
                assert s[p][3] == at
                at += s[p][2]
            assert at == first + count


# The reference's own sweep, over droplet sizes (main.c:158-164), as one run of the partitioned pool
SIZE_POINTS = [(6000, 2500, 0), (10000, 3000, 50), (16000, 1200, 7), (30000, 900, 123456789012)]


@pytest.mark.gpu
@pytest.mark.parametrize("mode,pool_slots", [("exact", 0), ("exact", 1024), ("fast", 0), ("fast", 2048)])
def test_size_sweep_equals_separate_runs(table, mode, pool_slots):
    """Every droplet size of heroam_gpu_run_size_sweep -- its own sphere radius, constants and scaled table copy,
    sharing the passes of one pool run with the other sizes -- must equal that size run on its own."""
    max_time = 3.0e4
    with engine.Engine(make_config(max_time=max_time), table, mode=mode, pool_slots=pool_slots) as eng:
        swept = eng.run_size_sweep(SIZE_POINTS)
        info = swept[0]["info"]
        singles = [eng.run_stats(he, count, first_index=first) for he, count, first in SIZE_POINTS]
    assert len(swept) == len(SIZE_POINTS)
    for (he, count, first), got, want in zip(SIZE_POINTS, swept, singles):
        assert_same_stats(got, want, f"N_He = {he}")
    assert info["trajectories"] == sum(p[1] for p in SIZE_POINTS)
    assert info["particle_steps"] == sum(s["info"]["particle_steps"] for s in singles)
    # different sizes do differ (the sweep is not one size four times over)
    assert not np.array_equal(swept[0]["hist_meet_time"], swept[3]["hist_meet_time"])


@pytest.mark.gpu
def test_size_sweep_argument_checks(table):
    with engine.Engine(make_config(), table) as eng:
        with pytest.raises(engine.HeroamError, match="sizes"):
            eng.run_size_sweep([(10000, 1, 0)] * 33)
        one = eng.run_size_sweep([(10000, 40, 3)])
        assert one[0]["trajectories"] == 40
        assert eng.run_size_sweep([(10000, 0, 0), (20000, 0, 0)])[1]["trajectories"] == 0
