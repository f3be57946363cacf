"""ctypes binding of libheroam_host.so: the C host side (config parser, force table,
result files) -- see host/heroam_host.h."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .types import PARTICLE_DTYPE, Config, ParticlesC, offsets_of

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libheroam_host.so")
_lib = None


def load_library() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `make -C host`")
        L = C.CDLL(LIB_PATH)
        L.heroam_read_config.restype = C.c_int
        L.heroam_read_config.argtypes = [C.c_char_p, C.POINTER(Config)]
        L.heroam_load_force_file.restype = C.c_long
        L.heroam_load_force_file.argtypes = [C.c_char_p, C.POINTER(C.c_float), C.c_long]
        L.heroam_synthetic_table.restype = None
        L.heroam_synthetic_table.argtypes = [C.POINTER(C.c_float), C.c_long, C.c_double, C.c_double, C.c_double, C.c_int]
        L.heroam_write_forcelist.restype = C.c_int
        L.heroam_write_forcelist.argtypes = [C.c_char_p, C.POINTER(Config), C.POINTER(C.c_float), C.c_long]
        L.heroam_write_particlefile.restype = C.c_int
        L.heroam_write_particlefile.argtypes = [C.c_char_p, C.POINTER(ParticlesC), C.c_int]
        _lib = L
    return _lib


def read_config(path: str) -> Config:
    """``read_config`` of the reference (simulation.c:48-156); raises if the file cannot be opened."""
    conf = Config()
    if load_library().heroam_read_config(path.encode(), C.byref(conf)) != 1:
        raise FileNotFoundError(path)
    return conf


def load_force_file(path: str, length: int):
    """Column 2 of the reference's 4-column text file -> (float32 table, rows read)."""
    table = np.zeros(length, dtype=np.float32)
    n = load_library().heroam_load_force_file(path.encode(), table.ctypes.data_as(C.POINTER(C.c_float)), length)
    if n < 0:
        raise FileNotFoundError(path)
    return table, int(n)


def synthetic_table_c(length: int, start: float = 4.0, grid: float = 0.001, amplitude: float = 1e-6, half_power: int = 2):
    table = np.zeros(length, dtype=np.float32)
    load_library().heroam_synthetic_table(table.ctypes.data_as(C.POINTER(C.c_float)), length, start, grid, amplitude, half_power)
    return table


def write_forcelist(path: str, conf: Config, table: np.ndarray, count: int) -> None:
    table = np.ascontiguousarray(table, dtype=np.float32)
    if not load_library().heroam_write_forcelist(path.encode(), C.byref(conf), table.ctypes.data_as(C.POINTER(C.c_float)), count):
        raise OSError(path)


def write_particlefile(path: str, counts: np.ndarray, flat: np.ndarray, first_index: int = 0) -> None:
    """``particlefile_<N>`` / ``particlefile_after_<N>`` (main.c:60-89) from a flat batch."""
    flat = np.ascontiguousarray(flat)
    assert flat.dtype == PARTICLE_DTYPE
    off = offsets_of(counts)
    arr = (ParticlesC * len(counts))()
    for i in range(len(counts)):
        arr[i].no, arr[i].index = int(counts[i]), first_index + i
        arr[i].particle = flat.ctypes.data + int(off[i]) * PARTICLE_DTYPE.itemsize
    if not load_library().heroam_write_particlefile(path.encode(), arr, len(counts)):
        raise OSError(path)
