"""The C host side (host/*.c): config parser, force-file loader, result files -- against the
reference's own read_config / save_particles (oracle/_ref) and committed goldens."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_golden
from sphere_roaming_b200 import host
from sphere_roaming_b200.forcetable import synthetic_table, write_force_file
from sphere_roaming_b200.types import Config, make_config, offsets_of

EXAMPLE = os.path.join(ROOT, "host", "config.example.conf")


def test_config_parser_reads_the_reference_keys():
    conf = host.read_config(EXAMPLE)
    want = dict(print=1, saveflag=0, hv=np.float32(23.8), intensity=1.0, cross_section=15.0, pulse_width=100.0,
                mass=4.0, velocity=28.0, icd_dist=6.0, number=1000, start=0, step=0, stop=0, helium_number=10000,
                dt=1.0, max_time=1.0e7, seed=100401452, force_grid=np.float32(0.001), force_grid_start=4.0,
                force_grid_length=1596000, force_file="synthetic")
    got = conf.as_dict()
    for k, v in want.items():
        assert got[k] == v, (k, got[k], v)


def test_config_parser_matches_reference_read_config(ref, tmp_path):
    """Same bytes in the struct as the reference's own parser on a file with comments,
    blank lines, odd spacing, an unknown key and a line without '='."""
    text = ("# comment\n\n  hv=23.8   # eV\nintensity =   2.5\ncross_section= 15\npulse_width = 50 # fs\n"
            "number = 12\nsaveflag = 0\nprint = 0\nstart = 100\nstep = 50\nstop = 300\nhelium_number = 5000\n"
            "seed = 42\nmass = 4.0\nvelocity = 28.0\ndt = 0.5\nmax_time = 2500\nicd_dist = 6.0\n"
            "force_grid = 0.001\nforce_grid_start = 4.0\nforce_grid_length = 1596000\n"
            "force_file = ./data/fit_force.txt\nbogus_key = 1\nthis line has no equals sign\n")
    path = tmp_path / "c.conf"
    path.write_text(text)
    mine = host.read_config(str(path))
    theirs = Config()  # zero-initialised, so padding and unset bytes compare equal
    assert ref.lib.ref_read_config(str(path).encode(), C.byref(theirs)) == 1
    assert bytes(mine) == bytes(theirs)
    assert mine.dt == 0.5 and mine.start == 100 and mine.force_file == b"./data/fit_force.txt"
    with pytest.raises(FileNotFoundError):
        host.read_config(str(tmp_path / "missing.conf"))


def test_force_table_generators_agree(table):
    assert np.array_equal(host.synthetic_table_c(len(table)), table)
    t2, n = host.load_force_file("synthetic", 1000)
    assert n == 1000 and np.array_equal(t2, table[:1000])


def test_force_file_text_round_trip(tmp_path, table):
    """The reference's 4-column text format (main.c:136); extra rows are ignored instead of
    overflowing (Q12), a short file zero-fills."""
    path = str(tmp_path / "force.txt")
    write_force_file(path, table[:500])
    got, n = host.load_force_file(path, 400)
    assert n == 400 and np.array_equal(got, table[:400])
    got, n = host.load_force_file(path, 600)
    assert n == 500 and np.array_equal(got[:500], table[:500]) and np.all(got[500:] == 0)
    with pytest.raises(FileNotFoundError):
        host.load_force_file(str(tmp_path / "nope.txt"), 10)
    conf = make_config()
    out = str(tmp_path / "forcelist.txt")
    host.write_forcelist(out, conf, table, 3)
    lines = open(out).read().splitlines()
    assert lines[0] == "Distance_Squared\tForce" and lines[1] == "4.000000\t%.10e" % table[0] and len(lines) == 4


def test_particlefile_format_matches_reference(tmp_path):
    g = load_golden("trajectories.npz")
    off = offsets_of(g["counts"])
    path = str(tmp_path / "particlefile_after_10000")
    host.write_particlefile(path, g["counts"][:6], g["final"][: off[6]])
    want = open(os.path.join(GOLDEN, "particlefile_after_golden.tsv")).read()
    assert open(path).read() == want


def test_cli_usage_and_errors(tmp_path):
    exe = os.path.join(ROOT, "host", "HeRoam")
    if not os.path.exists(exe):
        pytest.skip("host/HeRoam not built")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage:" in r.stderr
    r = subprocess.run([exe, str(tmp_path / "missing.conf")], capture_output=True, text=True)
    assert r.returncode == 1 and "Error opening config file" in r.stderr


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver times next to the GPU arm): one JSON line
    with the contract's keys, produced by the reference's own object code on the host cores."""
    import json
    import subprocess
    import sys

    from oracle.binding import ref_available

    if not ref_available("release") and not ref_available("strict"):
        pytest.skip("oracle/_ref not built")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-sample", "8", "--max-time", "2e5"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "particle-steps/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in line["config"]
