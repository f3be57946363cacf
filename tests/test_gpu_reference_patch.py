"""INTEGRATION.md section 2, compiled and run: the reference's own driver flow with the one change a
maintainer would make -- the OpenMP loop over simulate_particles (reference src/main.c:76-78) replaced
by heroam_gpu_simulate_batch -- against the reference's unmodified main.c.

Both executables are linked by oracle/Makefile from the reference's sources where they lie
(simulation.c, vector.c; main.c for the first) and travel prebuilt in oracle/_ref/:

  _ref/HeRoam_ref      main.c + simulation.c + vector.c, unmodified (HDF5 entry points as no-ops)
  _ref/HeRoam_patched  oracle/ref_patched_driver.c (includes the reference's simulation.h, then
                       heroam_gpu.h: the layout static assertions of heroam_types.h) + the reference's
                       read_config / initialize_particles / save_particles / free_particles

Same config file, same text force file, same libc-rand initial states: ./results/particlefile_<N>
and ./results/particlefile_after_<N> must be equal byte for byte (exact mode)."""
from __future__ import annotations

import os
import subprocess

import pytest

from sphere_roaming_b200 import engine
from sphere_roaming_b200.forcetable import synthetic_table, write_force_file

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "HeRoam_ref")
PATCHED_CLI = os.path.join(ROOT, "oracle", "_ref", "HeRoam_patched")

CONFIG = """print = 0
hv = 23.8
intensity = 1.0
cross_section = 15.0
pulse_width = 100.0
number = {number}
saveflag = 0
start = {start}
step = {step}
stop = {stop}
helium_number = 10000
seed = 100401452
mass = 4.0
velocity = 28.0
dt = 1.0
max_time = {max_time}
icd_dist = 6.0
force_grid = 0.1
force_grid_start = 4.0
force_grid_length = 15960
force_file = force.txt
"""


def prepare(tmp_path, **kw):
    d = str(tmp_path)
    write_force_file(os.path.join(d, "force.txt"), synthetic_table(length=15960, grid=0.1), start=4.0, grid=0.1)
    with open(os.path.join(d, "config.conf"), "w") as f:
        f.write(CONFIG.format(**kw))
    return d


def run(exe, cwd, out):
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    subprocess.run([exe, "config.conf"], cwd=cwd, check=True, timeout=600, env=env)
    os.rename(os.path.join(cwd, "results"), os.path.join(cwd, out))


def need_binaries():
    for exe in (REF_CLI, PATCHED_CLI):
        if not os.path.exists(exe):
            pytest.fail(f"{exe} is missing: `make -C oracle ref` where /root/reference exists (it travels prebuilt)")


@pytest.mark.gpu
@pytest.mark.parametrize("sweep", [False, True])
def test_patched_reference_driver_writes_the_reference_files(tmp_path, sweep):
    if engine.device_count() < 1:
        pytest.fail("needs a CUDA device: the engine has no CPU fallback")
    need_binaries()
    kw = dict(number=96, max_time=200000.0, start=0, step=0, stop=0)
    sizes = [10000]
    if sweep:  # the droplet-size sweep of main.c:158-164
        kw.update(number=32, max_time=50000.0, start=6000, step=3000, stop=12001)
        sizes = [6000, 9000, 12000]
    d = prepare(tmp_path, **kw)
    run(REF_CLI, d, "results_ref")
    run(PATCHED_CLI, d, "results_gpu")
    for n in sizes:
        for stem in ("particlefile", "particlefile_after"):
            a = open(os.path.join(d, "results_ref", f"{stem}_{n}"), "rb").read()
            b = open(os.path.join(d, "results_gpu", f"{stem}_{n}"), "rb").read()
            assert len(a) > 1000, f"{stem}_{n} is empty"
            assert a == b, f"{stem}_{n} differs between the reference and the patched driver"


def test_reference_cli_builds_and_patched_driver_has_no_cpu_path(tmp_path):
    """CPU side: the reference's own CLI runs from its unmodified sources; the patched driver, which
    was compiled with the reference's simulation.h ahead of heroam_gpu.h (layout assertions), refuses
    to compute without a GPU."""
    if not os.path.exists("/root/reference/src/main.c") and not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref executables were never built (needs /root/reference once)")
    need_binaries()
    d = prepare(tmp_path, number=8, max_time=2000.0, start=0, step=0, stop=0)
    run(REF_CLI, d, "results_ref")
    text = open(os.path.join(d, "results_ref", "particlefile_after_10000")).read().splitlines()
    assert text[0].split("\t") == ["Index", "Number", "Success", "Time", "X", "Y", "Z", "VX", "VY", "VZ", "FX", "FY", "FZ"]
    assert len(text) > 8
    if engine.device_count() < 1:
        p = subprocess.run([PATCHED_CLI, "config.conf"], cwd=d, capture_output=True, text=True, timeout=120)
        assert p.returncode != 0 and ("CUDA" in p.stderr or "device" in p.stderr)
