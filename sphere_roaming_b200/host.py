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
        L.heroam_h5_create.restype = C.c_void_p
        L.heroam_h5_create.argtypes = [C.c_char_p]
        L.heroam_h5_save_timestep.restype = C.c_int
        L.heroam_h5_save_timestep.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_uint]
        L.heroam_h5_close.restype = C.c_int
        L.heroam_h5_close.argtypes = [C.c_void_p]
        L.heroam_h5_open.restype = C.c_void_p
        L.heroam_h5_open.argtypes = [C.c_char_p]
        L.heroam_h5_num_steps.restype = C.c_long
        L.heroam_h5_num_steps.argtypes = [C.c_void_p]
        L.heroam_h5_read_timestep.restype = C.c_int
        L.heroam_h5_read_timestep.argtypes = [C.c_void_p, C.c_uint, C.POINTER(ParticlesC)]
        L.heroam_h5_reader_close.restype = None
        L.heroam_h5_reader_close.argtypes = [C.c_void_p]
        L.heroam_h5_last_error.restype = C.c_char_p
        L.free_records = C.CDLL(None).free
        L.free_records.argtypes = [C.c_void_p]
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


class TrajectoryFileWriter:
    """``trajectory_%06d.h5`` in the reference's layout (h5file.c:69-112): one group
    ``/Step_<k>`` per call, holding the dataset ``particles``."""

    def __init__(self, path: str):
        self._lib = load_library()
        self._h = self._lib.heroam_h5_create(path.encode())
        if not self._h:
            raise OSError(self._lib.heroam_h5_last_error().decode())

    def save_timestep(self, step: int, records: np.ndarray) -> None:
        records = np.ascontiguousarray(records)
        assert records.dtype == PARTICLE_DTYPE
        if self._lib.heroam_h5_save_timestep(self._h, step, records.ctypes.data, len(records)) != 0:
            raise OSError(self._lib.heroam_h5_last_error().decode())

    def close(self) -> None:
        if self._h:
            h, self._h = self._h, None
            if self._lib.heroam_h5_close(h) != 0:
                raise OSError(self._lib.heroam_h5_last_error().decode())

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class TrajectoryFileReader:
    """``open_file`` / ``read_timestep_from_hdf5`` of the reference (h5file.c:78-143)."""

    def __init__(self, path: str):
        self._lib = load_library()
        self._h = self._lib.heroam_h5_open(path.encode())
        if not self._h:
            raise OSError(self._lib.heroam_h5_last_error().decode())

    def num_steps(self) -> int:
        return int(self._lib.heroam_h5_num_steps(self._h))

    def read_timestep(self, step: int) -> np.ndarray:
        out = ParticlesC()
        if self._lib.heroam_h5_read_timestep(self._h, step, C.byref(out)) != 0:
            raise KeyError(self._lib.heroam_h5_last_error().decode())
        try:
            n = int(out.no)
            buf = C.string_at(out.particle, n * PARTICLE_DTYPE.itemsize)
            return np.frombuffer(buf, dtype=PARTICLE_DTYPE).copy()
        finally:
            self._lib.free_records(out.particle)

    def close(self) -> None:
        if self._h:
            self._lib.heroam_h5_reader_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
