"""Shared fixtures.  `-m "not gpu"` runs here on CPU; `-m gpu` needs a B200."""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def table():
    from sphere_roaming_b200.forcetable import synthetic_table

    return synthetic_table()


@pytest.fixture(scope="session")
def port():
    """The repo's C restatement of the reference (built on demand with gcc)."""
    from oracle.binding import PortOracle

    return PortOracle()


@pytest.fixture(scope="session")
def ref():
    """The reference's own object code (oracle/_ref); absent only if never built."""
    from oracle.binding import RefOracle, build_ref, ref_available

    if not ref_available("strict") and os.path.exists("/root/reference/src/simulation.c"):
        build_ref()
    if not ref_available("strict"):
        pytest.skip("oracle/_ref not built (needs /root/reference once)")
    return RefOracle("strict")


def load_golden(name: str):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def golden_traj():
    return load_golden("trajectories.npz")


def assert_records_equal(got: np.ndarray, want: np.ndarray, what: str = ""):
    """Field-by-field equality of Particle records (float ==, so -0.0 == +0.0)."""
    from sphere_roaming_b200.types import PARTICLE_DTYPE

    assert len(got) == len(want), f"{what}: {len(got)} records vs {len(want)}"
    for name in PARTICLE_DTYPE.names:
        a, b = got[name], want[name]
        if not np.array_equal(a, b):
            bad = np.nonzero((a != b).reshape(len(a), -1).any(axis=1))[0]
            raise AssertionError(
                f"{what}: field '{name}' differs in {len(bad)} of {len(a)} records, first at {bad[0]}: "
                f"{a[bad[0]]} vs {b[bad[0]]}"
            )


@pytest.fixture(scope="session")
def gpu_available():
    from sphere_roaming_b200 import engine

    return engine.device_count() > 0
