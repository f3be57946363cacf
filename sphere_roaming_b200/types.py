"""Boundary data types, mirrored for the Python host side.

The layouts are those of ``include/heroam_types.h`` which in turn restate the
reference's ``struct config`` (src/simulation.h:10-32), ``Particle``
(src/simulation.h:37-48, 84 bytes) and ``Particles`` (src/simulation.h:53-57).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

#: numpy view of one ``Particle`` record (reference src/simulation.h:37-48).
PARTICLE_DTYPE = np.dtype(
    [
        ("index", "<u4"),
        ("success", "<u4"),
        ("time", "<f4"),
        ("force_mag", "<f4"),
        ("velocity_sq", "<f4"),
        ("orientation", "<f4", (4,)),
        ("position", "<f4", (3,)),
        ("velocity", "<f4", (3,)),
        ("ang_velocity", "<f4", (3,)),
        ("force", "<f4", (3,)),
    ]
)
assert PARTICLE_DTYPE.itemsize == 84


class Config(C.Structure):
    """``struct config`` of the reference (src/simulation.h:10-32), 152 bytes."""

    _fields_ = [
        ("print", C.c_int),
        ("hv", C.c_float),
        ("intensity", C.c_float),
        ("cross_section", C.c_float),
        ("pulse_width", C.c_float),
        ("number", C.c_int),
        ("saveflag", C.c_int),
        ("start", C.c_int),
        ("step", C.c_int),
        ("stop", C.c_int),
        ("helium_number", C.c_int),
        ("seed", C.c_long),
        ("mass", C.c_float),
        ("velocity", C.c_float),
        ("dt", C.c_float),
        ("max_time", C.c_float),
        ("icd_dist", C.c_float),
        ("force_grid", C.c_float),
        ("force_grid_start", C.c_float),
        ("force_grid_length", C.c_long),
        ("force_file", C.c_char * 50),
    ]

    def as_dict(self) -> dict:
        out = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            out[name] = v.decode() if isinstance(v, bytes) else v
        return out


assert C.sizeof(Config) == 152


class ParticlesC(C.Structure):
    """``Particles`` of the reference (src/simulation.h:53-57), 16 bytes."""

    _fields_ = [("no", C.c_uint), ("index", C.c_uint), ("particle", C.c_void_p)]


assert C.sizeof(ParticlesC) == 16

#: Physics of the repository's own ``config.conf`` (reference config.conf:7-33)
#: with saveflag off; ``number`` and ``max_time`` are the per-run knobs.
REPO_CONFIG = dict(
    print=0,
    saveflag=0,
    hv=23.8,
    intensity=1.0,
    cross_section=15.0,
    pulse_width=100.0,
    mass=4.0,
    velocity=28.0,
    icd_dist=6.0,
    number=1,
    start=0,
    step=0,
    stop=0,
    helium_number=10000,
    dt=1.0,
    max_time=1.0e7,
    seed=100401452,
    force_grid=0.001,
    force_grid_start=4.0,
    force_grid_length=1596000,
    force_file=b"./data/fit_force.txt",
)


def make_config(**overrides) -> Config:
    """A zero-initialised ``Config`` filled from REPO_CONFIG plus overrides."""
    conf = Config()
    vals = dict(REPO_CONFIG)
    vals.update(overrides)
    for k, v in vals.items():
        if k == "force_file" and isinstance(v, str):
            v = v.encode()
        setattr(conf, k, v)
    return conf


def offsets_of(counts: np.ndarray) -> np.ndarray:
    """Exclusive prefix sum: first record of every trajectory in a flat batch."""
    off = np.zeros(len(counts) + 1, dtype=np.int64)
    np.cumsum(counts, out=off[1:])
    return off
