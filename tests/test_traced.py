"""The saveflag = 1 path: every iteration's state, as the reference hands it to
save_timestep_to_hdf5 (simulation.c:475-476), and the trajectory_%06d.h5 files built from it.

Golden: tests/golden/traced_steps.npz, recorded from the reference's own object code
(oracle/_ref, strict build) by tests/golden/make_golden_traced.py.
"""
from __future__ import annotations

import os
import subprocess

import numpy as np
import pytest

import h5lite
from conftest import ROOT, assert_records_equal, load_golden
from sphere_roaming_b200 import engine, host
from sphere_roaming_b200.types import PARTICLE_DTYPE, make_config, offsets_of


@pytest.fixture(scope="module")
def golden():
    return load_golden("traced_steps.npz")


def golden_conf(g, number=1):
    return make_config(number=number, max_time=float(g["max_time"]), intensity=float(g["intensity"]),
                       helium_number=int(g["he_number"]), velocity=float(g["velocity"]))


def golden_batch(g):
    n = int(g["n"])
    counts = np.array([len(g[f"init_{k}"]) for k in range(n)], dtype=np.uint32)
    init = np.concatenate([g[f"init_{k}"] for k in range(n)])
    return n, counts, init


def test_golden_covers_meetings_and_short_trajectories(golden):
    g = golden
    n = int(g["n"])
    lengths = [len(g[f"snaps_{k}"]) for k in range(n)]
    assert 1 in lengths and max(lengths) == 200 and any(1 < x < 200 for x in lengths)
    frozen_during = sum(int(g[f"final_{k}"]["success"].sum() - g[f"init_{k}"]["success"].sum()) for k in range(n))
    assert frozen_during >= 16
    for k in range(n):
        s = g[f"snaps_{k}"]
        assert_records_equal(s[0], g[f"init_{k}"], "snapshot 0 is the imported state")
        # a frozen particle's record never changes again (Q8)
        for j in range(s.shape[1]):
            hit = np.nonzero(s["success"][:, j])[0]
            if len(hit) and hit[0] + 1 < len(s):
                assert all(s[i, j].tobytes() == s[hit[0] + 1, j].tobytes() for i in range(hit[0] + 1, len(s)))


def test_oracle_restatement_reproduces_every_snapshot(golden, port, table):
    """State after k iterations (the oracle's step cap) == snapshot k of the reference."""
    g = golden
    he = int(g["he_number"])
    for k in range(int(g["n"])):
        snaps, init = g[f"snaps_{k}"], g[f"init_{k}"]
        conf = golden_conf(g)
        counts = np.array([len(init)], dtype=np.uint32)
        for step in range(1, len(snaps)):
            got, _ = port.simulate(conf, he, table, counts, init, max_steps=step)
            assert_records_equal(got, snaps[step], f"trajectory {k}, state after {step} iterations")
        final, res = port.simulate(conf, he, table, counts, init)
        assert_records_equal(final, g[f"final_{k}"], f"trajectory {k}, final")
        assert res["steps"][0] == len(snaps)


def test_golden_snapshots_through_the_hdf5_file(golden, tmp_path):
    g = golden
    for k in range(int(g["n"])):
        snaps = g[f"snaps_{k}"]
        path = str(tmp_path / f"trajectory_{k:06d}.h5")
        with host.TrajectoryFileWriter(path) as w:
            for step in range(len(snaps)):
                w.save_timestep(step, snaps[step])
        with host.TrajectoryFileReader(path) as r:
            assert r.num_steps() == len(snaps)
            for step in range(len(snaps)):
                assert r.read_timestep(step).tobytes() == snaps[step].tobytes()


def test_h5read_exports_normalised_positions(golden, tmp_path):
    """`h5read <file> <time_step>` (reference reader.c:65-228): trajectory_<p>.bin holds the
    positions divided by |position of particle 0 at step 0|, as float32 triples."""
    exe = os.path.join(ROOT, "host", "h5read")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "host")])
    snaps = golden["snaps_0"]
    path = str(tmp_path / "trajectory_000000.h5")
    with host.TrajectoryFileWriter(path) as w:
        for step in range(len(snaps)):
            w.save_timestep(step, snaps[step])
    r = subprocess.run([exe, path, "50"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert f"Time steps in file: {len(snaps)}" in r.stdout
    p0 = snaps[0]["position"][0]
    radius = np.sqrt(np.float64(np.float32(np.float32(p0[0] * p0[0] + p0[1] * p0[1]) + p0[2] * p0[2])))
    for p in range(snaps.shape[1]):
        got = np.fromfile(tmp_path / f"trajectory_{p}.bin", dtype="<f4").reshape(-1, 3)
        want = (snaps["position"][:, p, :].astype(np.float64) / radius).astype(np.float32)
        assert np.array_equal(got, want)
    # the frames are drawn by one gnuplot script (gnuplot itself is not in the image)
    text = (tmp_path / "images" / "frames.gp").read_text()
    assert f"do for [t = 0:{len(snaps) - 1}:50]" in text and "images/step_%010d.png" in text
    assert f"for [p = 0:{snaps.shape[1] - 1}] path(p)" in text and "$parallels << EOD" in text
    # the usage message of the reference
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage:" in r.stderr


# ------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------
def collect(eng, he, counts, init):
    seen = {}

    def sink(traj, step, records):
        seen.setdefault(traj, []).append((step, records.copy()))

    final, res = eng.simulate_traced(he, counts, init, sink)
    return seen, final, res


@pytest.mark.gpu
@pytest.mark.parametrize("segment", [0, 7])
def test_traced_exact_matches_reference_golden(golden, table, segment):
    g = golden
    n, counts, init = golden_batch(g)
    with engine.Engine(golden_conf(g, n), table, mode="exact", segment_steps=segment) as eng:
        seen, final, res = collect(eng, int(g["he_number"]), counts, init)
    off = offsets_of(counts)
    for k in range(n):
        snaps = g[f"snaps_{k}"]
        assert [s for s, _ in seen[k]] == list(range(len(snaps))), f"trajectory {k}: steps delivered in order"
        for step, rec in seen[k]:
            assert_records_equal(rec, snaps[step], f"trajectory {k}, iteration {step}")
        assert res["steps"][k] == len(snaps)
        assert_records_equal(final[off[k]:off[k + 1]], g[f"final_{k}"], f"trajectory {k}, final")


@pytest.mark.gpu
def test_traced_matches_plain_simulate_and_live_reference(table, ref):
    """A larger batch against the reference run live with saveflag = 1, and the traced path's
    final records against the production kernels'."""
    he, n = 10000, 96
    conf = make_config(number=n, max_time=1500.0, velocity=9000.0)
    counts, init = ref.initialize(conf, he, table)
    off = offsets_of(counts)
    with engine.Engine(conf, table, mode="exact") as eng:
        seen, final, res = collect(eng, he, counts, init)
        plain, pres = eng.simulate(he, counts, init)
    assert_records_equal(final, plain, "traced vs production kernels")
    assert np.array_equal(res, pres)
    for i in range(n):
        snaps, rfinal = ref.simulate_traced(conf, he, table, init[off[i]:off[i + 1]], 1502)
        assert len(seen[i]) == len(snaps) == res["steps"][i]
        got = np.stack([r for _, r in seen[i]])
        for name in PARTICLE_DTYPE.names:
            assert np.array_equal(got[name], snaps[name]), f"trajectory {i}: field {name}"
        assert_records_equal(final[off[i]:off[i + 1]], rfinal, f"trajectory {i}, final")
    assert (res["status"] == engine.DONE_SUCCESS).sum() > 10, "the batch should contain finished trajectories"


@pytest.mark.gpu
def test_traced_fast_mode_within_tolerance(golden, table):
    g = golden
    n, counts, init = golden_batch(g)
    with engine.Engine(golden_conf(g, n), table, mode="fast") as eng:
        seen, final, res = collect(eng, int(g["he_number"]), counts, init)
    for k in range(n):
        snaps = g[f"snaps_{k}"]
        assert abs(len(seen[k]) - len(snaps)) <= 2
        m = min(len(seen[k]), len(snaps)) - 1
        err = np.abs(seen[k][m][1]["position"].astype(np.float64) - snaps[m]["position"]).max()
        assert err < 5e-3, f"trajectory {k}: {err} A after {m} iterations"


@pytest.mark.gpu
def test_cli_saveflag_writes_reference_layout(tmp_path, table, port):
    """HeRoam <config> with saveflag = 1: ./results/trajectory_%06d.h5, one /Step_<k>/particles
    per iteration, holding what the oracle computes for the same Philox stream."""
    exe = os.path.join(ROOT, "host", "HeRoam")
    conf_text = open(os.path.join(ROOT, "host", "config.example.conf")).read()
    conf_text = (conf_text.replace("\nnumber = 1000", "\nnumber = 70").replace("max_time = 10000000", "max_time = 300")
                 .replace("saveflag = 0", "saveflag = 1"))
    assert "saveflag = 1" in conf_text
    (tmp_path / "run.conf").write_text(conf_text)
    r = subprocess.run([exe, "run.conf"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    conf = make_config(number=70, max_time=300.0)
    counts, init = port.sample(conf, 10000, table, 70, source="philox")
    off = offsets_of(counts)
    final, res = port.simulate(conf, 10000, table, counts, init)
    for i in (0, 1, 33, 63, 64, 69):  # both sides of the 64-trajectory file groups
        path = tmp_path / "results" / f"trajectory_{i:06d}.h5"
        assert path.exists(), path
        one = np.array([counts[i]], dtype=np.uint32)
        with host.TrajectoryFileReader(str(path)) as rd:
            assert rd.num_steps() == res["steps"][i]
            for step in (0, 1, 17, int(res["steps"][i]) - 1):
                if step >= res["steps"][i]:
                    continue
                want = init[off[i]:off[i + 1]] if step == 0 else \
                    port.simulate(conf, 10000, table, one, init[off[i]:off[i + 1]], max_steps=step)[0]
                assert_records_equal(rd.read_timestep(step), want, f"trajectory {i}, /Step_{step}")
        f = h5lite.File(str(path))
        assert len(f.group_entries(f.root_header)) == res["steps"][i]
