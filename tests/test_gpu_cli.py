"""The HeRoam <config.conf> driver and the reference-named Python interface on a GPU."""
from __future__ import annotations

import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, assert_records_equal
from sphere_roaming_b200 import engine, host, simulation
from sphere_roaming_b200.types import make_config

pytestmark = pytest.mark.gpu


def parse_particlefile(path):
    rows = np.loadtxt(path, skiprows=1, ndmin=2)
    return rows


def test_heroam_cli_writes_reference_files(tmp_path, port, table):
    if engine.device_count() < 1:
        pytest.fail("needs a CUDA device")
    exe = os.path.join(ROOT, "host", "HeRoam")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "host")])
    conf_text = open(os.path.join(ROOT, "host", "config.example.conf")).read()
    conf_text = conf_text.replace("\nnumber = 1000", "\nnumber = 40").replace("max_time = 10000000", "max_time = 20000")
    (tmp_path / "run.conf").write_text(conf_text)
    r = subprocess.run([exe, "run.conf"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Configuration File Loaded" in r.stdout
    for name in ("forcelist.txt", "particlefile_10000", "particlefile_after_10000"):
        assert (tmp_path / "results" / name).exists(), name
    # the files carry exactly what the oracle computes for the same Philox stream (exact mode)
    conf = make_config(number=40, max_time=2.0e4)
    counts, init = port.sample(conf, 10000, table, 40, source="philox")
    final, _ = port.simulate(conf, 10000, table, counts, init)
    host.write_particlefile(str(tmp_path / "want_after"), counts, final)
    assert (tmp_path / "results" / "particlefile_after_10000").read_text() == (tmp_path / "want_after").read_text()
    before = parse_particlefile(tmp_path / "results" / "particlefile_10000")
    assert len(before) == counts.sum() and np.all(before[:, 3] == 0)
    # the streaming writer (runs above 2^21 trajectories; forced here): the same two files, byte for byte
    for name in ("particlefile_10000", "particlefile_after_10000"):
        os.rename(tmp_path / "results" / name, tmp_path / ("held_" + name))
    r = subprocess.run([exe, "run.conf"], cwd=tmp_path, capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, HEROAM_STREAM_ABOVE="8"))
    assert r.returncode == 0, r.stderr
    for name in ("particlefile_10000", "particlefile_after_10000"):
        assert (tmp_path / "results" / name).read_bytes() == (tmp_path / ("held_" + name)).read_bytes(), name
    # histogram-only mode
    r = subprocess.run([exe, "run.conf"], cwd=tmp_path, capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, HEROAM_STATS="1", HEROAM_MODE="fast"))
    assert r.returncode == 0 and "particle-steps/s" in r.stdout
    assert (tmp_path / "results" / "histograms_10000.tsv").exists()


def test_reference_named_interface(port, table):
    conf = make_config(number=48, max_time=1.0e4)
    batch = simulation.initialize_particles(10000, conf, table)
    counts, init = port.sample(conf, 10000, table, 48, source="philox")
    assert np.array_equal(batch.counts, counts)
    want, wres = port.simulate(conf, 10000, table, counts, batch.particle)
    simulation.simulate_particles(conf, 10000, table, batch)
    assert_records_equal(batch.particle, want, "simulate_particles")
    assert np.array_equal(batch.results["steps"], wres["steps"])
    simulation.free_particles(batch)
    assert len(batch.particle) == 0
