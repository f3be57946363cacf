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


def test_heroam_cli_size_sweep_as_one_run(tmp_path):
    """start/step/stop with HEROAM_STATS=1: all droplet sizes in one pool run, one histogram file per size, equal
    to the files of the sizes run one by one."""
    if engine.device_count() < 1:
        pytest.fail("needs a CUDA device")
    exe = os.path.join(ROOT, "host", "HeRoam")
    conf_text = open(os.path.join(ROOT, "host", "config.example.conf")).read()
    conf_text = conf_text.replace("\nnumber = 1000", "\nnumber = 600").replace("max_time = 10000000", "max_time = 20000")
    sweep = conf_text.replace("\nstart = 0", "\nstart = 6000").replace("\nstep = 0", "\nstep = 4000").replace("\nstop = 0", "\nstop = 14001")
    assert sweep != conf_text
    (tmp_path / "sweep.conf").write_text(sweep)
    env = dict(os.environ, HEROAM_STATS="1", HEROAM_MODE="exact")
    r = subprocess.run([exe, "sweep.conf"], cwd=tmp_path, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr
    assert "3 droplet sizes as one run" in r.stdout
    for n in (6000, 10000, 14000):
        one = conf_text.replace("helium_number = 10000", f"helium_number = {n}")
        (tmp_path / "one.conf").write_text(one)
        os.rename(tmp_path / "results" / f"histograms_{n}.tsv", tmp_path / f"swept_{n}.tsv")
        r = subprocess.run([exe, "one.conf"], cwd=tmp_path, capture_output=True, text=True, timeout=300, env=env)
        assert r.returncode == 0, r.stderr
        assert (tmp_path / "results" / f"histograms_{n}.tsv").read_bytes() == (tmp_path / f"swept_{n}.tsv").read_bytes(), n


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
