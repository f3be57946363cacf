"""ctypes bindings for the CPU oracles -- TEST INFRASTRUCTURE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  Two libraries:

* ``RefOracle``  : oracle/_ref/libheroam_ref_<flavour>.so -- the reference's own
  simulation.c + vector.c compiled unmodified (see oracle/Makefile).
* ``PortOracle`` : oracle/libheroam_oracle.so -- the repo's C restatement
  (oracle/heroam_oracle.c).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from sphere_roaming_b200.types import PARTICLE_DTYPE, Config, offsets_of

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_fp = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)


def _ptr(a: np.ndarray, typ):
    return a.ctypes.data_as(typ)


def build_port(force: bool = False) -> str:
    """Compile the C restatement (gcc); returns the .so path."""
    so = os.path.join(HERE, "libheroam_oracle.so")
    src = os.path.join(HERE, "heroam_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    return so


def build_ref() -> None:
    """Compile the reference by path when /root/reference exists (dev container)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def ref_available(flavour: str = "strict") -> bool:
    return os.path.exists(os.path.join(REF_DIR, f"libheroam_ref_{flavour}.so"))


class OrcResult(C.Structure):
    _fields_ = [("steps", C.c_uint32), ("status", C.c_uint32), ("n_flagged", C.c_uint32), ("time", C.c_float)]


RESULT_DTYPE = np.dtype([("steps", "<u4"), ("status", "<u4"), ("n_flagged", "<u4"), ("time", "<f4")])


class OrcCounters(C.Structure):
    _fields_ = [("particle_steps", C.c_uint64), ("pair_steps", C.c_uint64), ("inrange_steps", C.c_uint64)]


class RefOracle:
    """The reference's own object code behind oracle/ref_harness.c."""

    def __init__(self, flavour: str = "strict"):
        path = os.path.join(REF_DIR, f"libheroam_ref_{flavour}.so")
        self.lib = C.CDLL(path)
        self.flavour = flavour
        L = self.lib
        L.ref_initialize.restype = C.c_long
        L.ref_initialize.argtypes = [C.c_ulonglong, C.POINTER(Config), _fp, _u32p, C.c_void_p, C.c_long]
        L.ref_simulate.restype = None
        L.ref_simulate.argtypes = [C.c_ulonglong, C.POINTER(Config), _fp, C.c_int, _u32p, C.c_void_p, C.c_int, C.c_int]
        if hasattr(L, "ref_simulate_guarded"):
            L.ref_simulate_guarded.restype = None
            L.ref_simulate_guarded.argtypes = [C.c_ulonglong, C.POINTER(Config), _fp, C.c_int, _u32p, C.c_void_p, C.c_int,
                                               C.c_int, C.c_void_p]
        L.ref_calculate_force.restype = None
        L.ref_calculate_force.argtypes = [C.c_uint, C.c_void_p, C.POINTER(Config), _fp]
        if hasattr(L, "ref_simulate_traced"):
            L.ref_simulate_traced.restype = C.c_ulong
            L.ref_simulate_traced.argtypes = [C.c_ulonglong, C.POINTER(Config), _fp, C.c_uint, C.c_void_p, C.c_ulong, C.c_void_p]
        L.ref_find_success.restype = C.c_int
        L.ref_find_success.argtypes = [C.c_uint, C.c_void_p]
        L.ref_find_accelration.restype = None
        L.ref_find_accelration.argtypes = [C.c_float, C.c_float, _fp, _fp, _fp]
        L.ref_update_angvel.restype = None
        L.ref_update_angvel.argtypes = [C.c_float, _fp, _fp, _fp]
        L.ref_update_orientation.restype = None
        L.ref_update_orientation.argtypes = [C.c_float, C.c_float, _fp, _fp, _fp]
        L.ref_generate_velocity.restype = C.c_float
        L.ref_generate_velocity.argtypes = [C.POINTER(Config)]
        L.ref_generate_no_of_excitation.restype = C.c_int
        L.ref_generate_no_of_excitation.argtypes = [C.c_double]
        L.ref_srand.argtypes = [C.c_uint]
        L.ref_read_config.restype = C.c_int
        L.ref_read_config.argtypes = [C.c_char_p, C.POINTER(Config)]
        # vector.c functions are plain exported symbols of the reference itself
        for name, n_in in (("cross_product", 2), ("add_vectors", 2)):
            f = getattr(L, name)
            f.restype = None
            f.argtypes = [_fp] * (n_in + 1)
        L.norm.restype = C.c_float
        L.norm.argtypes = [_fp]
        L.norm_sq.restype = C.c_float
        L.norm_sq.argtypes = [_fp]
        L.scalar_multiply.restype = None
        L.scalar_multiply.argtypes = [C.c_float, _fp, _fp]
        L.create_perpend_vector.restype = None
        L.create_perpend_vector.argtypes = [_fp, _fp]
        for name in ("quat_to_pos", "quat_inverse"):
            f = getattr(L, name)
            f.restype = None
            f.argtypes = [_fp, _fp]
        L.quat_product.restype = None
        L.quat_product.argtypes = [_fp, _fp, _fp]
        L.quat_norm.restype = C.c_float
        L.quat_norm.argtypes = [_fp]
        L.quat_random.restype = None
        L.quat_random.argtypes = [_fp]

    # -- layout ---------------------------------------------------------------
    def sizes(self):
        return (self.lib.ref_sizeof_particle(), self.lib.ref_sizeof_particles(), self.lib.ref_sizeof_config())

    def particle_offsets(self):
        buf = (C.c_int * 16)()
        n = self.lib.ref_particle_offsets(buf)
        return list(buf[:n])

    def config_offsets(self):
        buf = (C.c_int * 32)()
        n = self.lib.ref_config_offsets(buf)
        return list(buf[:n])

    def max_threads(self) -> int:
        return int(self.lib.ref_max_threads())

    # -- stages ---------------------------------------------------------------
    def initialize(self, conf: Config, he_number: int, table: np.ndarray):
        """initialize_particles -> (counts[u4], flat Particle records)."""
        number = conf.number
        counts = np.zeros(number, dtype=np.uint32)
        cap = max(64, int(number * 8))
        while True:
            flat = np.zeros(cap, dtype=PARTICLE_DTYPE)
            got = self.lib.ref_initialize(he_number, C.byref(conf), _ptr(table, _fp), _ptr(counts, _u32p), flat.ctypes.data, cap)
            if got >= 0:
                return counts, flat[:got].copy()
            cap = -got

    def simulate(self, conf: Config, he_number: int, table: np.ndarray, counts: np.ndarray, flat: np.ndarray,
                 nthreads: int = 0, dynamic: bool = True) -> np.ndarray:
        """simulate_particles over a batch; returns the advanced copy."""
        out = np.ascontiguousarray(flat.copy())
        counts = np.ascontiguousarray(counts, dtype=np.uint32)
        self.lib.ref_simulate(he_number, C.byref(conf), _ptr(table, _fp), len(counts), _ptr(counts, _u32p),
                              out.ctypes.data, nthreads, 1 if dynamic else 0)
        return out

    def simulate_guarded(self, conf: Config, he_number: int, table: np.ndarray, counts: np.ndarray, flat: np.ndarray,
                         nthreads: int = 0, dynamic: bool = True):
        """simulate_particles over a batch with the harness's guard against the reference's one
        non-terminating case (every particle frozen, `==` criterion unmet: SURVEY.md Q4).  Returns
        (advanced copy, deadlocked[u1] per trajectory)."""
        out = np.ascontiguousarray(flat.copy())
        counts = np.ascontiguousarray(counts, dtype=np.uint32)
        dead = np.zeros(len(counts), dtype=np.uint8)
        self.lib.ref_simulate_guarded(he_number, C.byref(conf), _ptr(table, _fp), len(counts), _ptr(counts, _u32p),
                                      out.ctypes.data, nthreads, 1 if dynamic else 0, dead.ctypes.data)
        return out, dead

    def simulate_traced(self, conf: Config, he_number: int, table: np.ndarray, particles: np.ndarray, cap_steps: int):
        """One trajectory with saveflag=1: (snapshots[steps, no], final records).  The snapshots are
        what the reference hands to save_timestep_to_hdf5 at the top of every loop iteration."""
        final = np.ascontiguousarray(particles.copy())
        snaps = np.zeros((cap_steps, len(final)), dtype=PARTICLE_DTYPE)
        n = int(self.lib.ref_simulate_traced(he_number, C.byref(conf), _ptr(table, _fp), len(final), final.ctypes.data,
                                             cap_steps, snaps.ctypes.data))
        if n > cap_steps:
            raise ValueError(f"{n} steps, buffer holds {cap_steps}")
        return snaps[:n].copy(), final

    def calculate_force(self, conf: Config, table: np.ndarray, particles: np.ndarray) -> np.ndarray:
        out = np.ascontiguousarray(particles.copy())
        self.lib.ref_calculate_force(len(out), out.ctypes.data, C.byref(conf), _ptr(table, _fp))
        return out

    def find_success(self, particles: np.ndarray) -> int:
        p = np.ascontiguousarray(particles)
        return int(self.lib.ref_find_success(len(p), p.ctypes.data))


class PortOracle:
    """The repo's plain-C restatement (oracle/heroam_oracle.c)."""

    def __init__(self):
        self.lib = C.CDLL(build_port())
        L = self.lib
        L.orc_calculate_force.restype = None
        L.orc_calculate_force.argtypes = [C.c_uint32, C.c_void_p, C.POINTER(Config), _fp]
        L.orc_simulate.restype = None
        L.orc_simulate.argtypes = [C.POINTER(Config), C.c_ulonglong, _fp, C.c_int, _u32p, C.c_void_p, C.c_uint64,
                                   C.c_int, C.c_void_p, C.POINTER(OrcCounters)]
        L.orc_sample.restype = C.c_long
        L.orc_sample.argtypes = [C.POINTER(Config), C.c_ulonglong, _fp, C.c_int, C.c_uint64, C.c_int, _u32p,
                                 C.c_void_p, C.c_long]
        L.orc_philox4x32_10.restype = None
        L.orc_philox4x32_10.argtypes = [_u32p, _u32p, _u32p]
        L.orc_lambda.restype = C.c_double
        L.orc_lambda.argtypes = [C.POINTER(Config), C.c_ulonglong]
        L.orc_radius_sim.restype = C.c_float
        L.orc_radius_sim.argtypes = [C.c_ulonglong]
        L.orc_radius_init.restype = C.c_double
        L.orc_radius_init.argtypes = [C.c_ulonglong]

    def max_threads(self) -> int:
        return int(self.lib.orc_max_threads())

    def calculate_force(self, conf: Config, table: np.ndarray, particles: np.ndarray) -> np.ndarray:
        out = np.ascontiguousarray(particles.copy())
        self.lib.orc_calculate_force(len(out), out.ctypes.data, C.byref(conf), _ptr(table, _fp))
        return out

    def simulate(self, conf: Config, he_number: int, table: np.ndarray, counts: np.ndarray, flat: np.ndarray,
                 max_steps: int = 0, nthreads: int = 0, want_counters: bool = False):
        """Returns (advanced records, per-trajectory results[, counters])."""
        out = np.ascontiguousarray(flat.copy())
        counts = np.ascontiguousarray(counts, dtype=np.uint32)
        res = np.zeros(len(counts), dtype=RESULT_DTYPE)
        ctr = OrcCounters()
        self.lib.orc_simulate(C.byref(conf), he_number, _ptr(table, _fp), len(counts), _ptr(counts, _u32p),
                              out.ctypes.data, max_steps, nthreads, res.ctypes.data,
                              C.byref(ctr) if want_counters else None)
        if want_counters:
            return out, res, dict(particle_steps=ctr.particle_steps, pair_steps=ctr.pair_steps,
                                  inrange_steps=ctr.inrange_steps)
        return out, res

    def sample(self, conf: Config, he_number: int, table: np.ndarray, number: int, source: str = "philox",
               first_index: int = 0):
        """initialize_particles restated; source 'rand' (libc, = reference) or 'philox'."""
        counts = np.zeros(number, dtype=np.uint32)
        cap = max(64, int(number * 8))
        src = 0 if source == "rand" else 1
        while True:
            flat = np.zeros(cap, dtype=PARTICLE_DTYPE)
            got = self.lib.orc_sample(C.byref(conf), he_number, _ptr(table, _fp), src, first_index, number,
                                      _ptr(counts, _u32p), flat.ctypes.data, cap)
            if got >= 0:
                return counts, flat[:got].copy()
            cap = -got

    def philox(self, ctr, key):
        c = np.asarray(ctr, dtype=np.uint32)
        k = np.asarray(key, dtype=np.uint32)
        o = np.zeros(4, dtype=np.uint32)
        self.lib.orc_philox4x32_10(_ptr(c, _u32p), _ptr(k, _u32p), _ptr(o, _u32p))
        return o

    def lam(self, conf: Config, he_number: int) -> float:
        return float(self.lib.orc_lambda(C.byref(conf), he_number))

    def radius_sim(self, he_number: int) -> float:
        return float(self.lib.orc_radius_sim(he_number))


__all__ = ["RefOracle", "PortOracle", "build_port", "build_ref", "ref_available", "offsets_of", "RESULT_DTYPE"]
