"""ctypes host binding of libheroam_gpu.so (include/heroam_gpu.h).

This is plumbing around the C ABI: every call below is one C-ABI call.  There is
no CPU fallback -- if the CUDA library is missing or no B200 is visible the
constructor raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .types import PARTICLE_DTYPE, Config, ParticlesC

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HEROAM_LIB") or os.path.join(_HERE, "libheroam_gpu.so")  # the env override: A/B builds only

MODE_EXACT = 0
MODE_FAST = 1
MODE_FAST_PLAIN = 2
MODE_FAST_RELAXED = 3
MODE_FAST_CLOSED = 4
_MODES = {"exact": MODE_EXACT, "fast": MODE_FAST, "fast_plain": MODE_FAST_PLAIN, "fast_relaxed": MODE_FAST_RELAXED,
          "fast_closed": MODE_FAST_CLOSED, 0: MODE_EXACT, 1: MODE_FAST, 2: MODE_FAST_PLAIN, 3: MODE_FAST_RELAXED,
          4: MODE_FAST_CLOSED}

DONE_SUCCESS, DONE_MAXTIME, DONE_DEADLOCK, DONE_STEPCAP, DONE_UNSUPPORTED = 0, 1, 2, 4, 6

HIST_N = 256
HIST_T = 4096

RESULT_DTYPE = np.dtype([("steps", "<u4"), ("status", "<u4"), ("n_flagged", "<u4"), ("time", "<f4")])
#: numpy view of the reference's ``Particles`` header (simulation.h:53-57): no, index, Particle*
PARTICLES_DTYPE = np.dtype([("no", "<u4"), ("index", "<u4"), ("particle", "<u8")])
assert PARTICLES_DTYPE.itemsize == 16

#: every symbol include/heroam_gpu.h declares
ABI_SYMBOLS = (
    "heroam_gpu_abi_version",
    "heroam_gpu_device_count",
    "heroam_gpu_last_error",
    "heroam_gpu_create",
    "heroam_gpu_destroy",
    "heroam_gpu_set_mode",
    "heroam_gpu_simulate_batch",
    "heroam_gpu_reserve",
    "heroam_gpu_simulate_flat",
    "heroam_gpu_simulate_traced",
    "heroam_gpu_upload",
    "heroam_gpu_run",
    "heroam_gpu_download",
    "heroam_gpu_sample_batch",
    "heroam_gpu_sample_flat",
    "heroam_gpu_free_particles",
    "heroam_gpu_run_stats",
    "heroam_gpu_run_sweep",
    "heroam_gpu_run_size_sweep",
    "heroam_gpu_last_run_info",
    "heroam_gpu_fp32_peak",
    "heroam_gpu_fp32_peak_rrr",
    "heroam_gpu_fp32_peak_x2",
)


class HeroamError(RuntimeError):
    pass


class Options(C.Structure):
    _fields_ = [
        ("mode", C.c_int),
        ("segment_steps", C.c_uint),
        ("max_steps", C.c_ulonglong),
        ("wide_kernel", C.c_int),
        ("pool_slots", C.c_uint),
        ("count_work", C.c_int),
        ("no_flight", C.c_int),
    ]


class RunInfo(C.Structure):
    _fields_ = [
        ("trajectories", C.c_ulonglong),
        ("particles", C.c_ulonglong),
        ("particle_steps", C.c_ulonglong),
        ("pair_steps", C.c_ulonglong),
        ("inrange_steps", C.c_ulonglong),
        ("passes", C.c_ulonglong),
        ("kernel_launches", C.c_ulonglong),
        ("device_ms", C.c_double),
        ("integrate_ms", C.c_double),
        ("h2d_ms", C.c_double),
        ("d2h_ms", C.c_double),
        ("h2d_bytes", C.c_ulonglong),
        ("d2h_bytes", C.c_ulonglong),
        ("pair_skipped", C.c_ulonglong),
        ("free_steps", C.c_ulonglong),
        ("resolve_ms", C.c_double),
    ]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


class Stats(C.Structure):
    _fields_ = [
        ("trajectories", C.c_ulonglong),
        ("particles", C.c_ulonglong),
        ("hist_n", C.c_ulonglong * HIST_N),
        ("hist_traj_time", C.c_ulonglong * HIST_T),
        ("hist_meet_time", C.c_ulonglong * HIST_T),
        ("done", C.c_ulonglong * 8),
        ("unmet_particles", C.c_ulonglong),
        ("sum_traj_time", C.c_double),
        ("info", RunInfo),
    ]


#: heroam_step_sink of include/heroam_gpu.h
STEP_SINK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32)

class SweepPoint(C.Structure):
    _fields_ = [("intensity", C.c_float), ("pulse_width", C.c_float), ("first_index", C.c_ulonglong),
                ("count", C.c_ulonglong)]


class SizePoint(C.Structure):
    _fields_ = [("he_number", C.c_ulonglong), ("first_index", C.c_ulonglong), ("count", C.c_ulonglong)]


_lib = None


def load_library() -> C.CDLL:
    """Load the CUDA library; raise (never fall back) when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HeroamError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)"
        )
    L = C.CDLL(LIB_PATH)
    vp, u32p = C.c_void_p, C.POINTER(C.c_uint32)
    L.heroam_gpu_abi_version.restype = C.c_int
    L.heroam_gpu_device_count.restype = C.c_int
    L.heroam_gpu_last_error.restype = C.c_char_p
    L.heroam_gpu_create.restype = C.c_int
    L.heroam_gpu_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_float), C.c_int, C.POINTER(Options), C.POINTER(vp)]
    L.heroam_gpu_destroy.restype = None
    L.heroam_gpu_destroy.argtypes = [vp]
    L.heroam_gpu_set_mode.restype = C.c_int
    L.heroam_gpu_set_mode.argtypes = [vp, C.c_int]
    L.heroam_gpu_simulate_batch.restype = C.c_int
    L.heroam_gpu_simulate_batch.argtypes = [vp, C.c_ulonglong, C.POINTER(ParticlesC), C.c_int, vp]
    L.heroam_gpu_reserve.restype = C.c_int
    L.heroam_gpu_reserve.argtypes = [vp, C.c_ulonglong, C.c_ulonglong]
    L.heroam_gpu_simulate_flat.restype = C.c_int
    L.heroam_gpu_simulate_flat.argtypes = [vp, C.c_ulonglong, u32p, vp, C.c_int, vp]
    L.heroam_gpu_simulate_traced.restype = C.c_int
    L.heroam_gpu_simulate_traced.argtypes = [vp, C.c_ulonglong, u32p, vp, C.c_int, vp, STEP_SINK, vp]
    L.heroam_gpu_upload.restype = C.c_int
    L.heroam_gpu_upload.argtypes = [vp, C.c_ulonglong, u32p, vp, C.c_int]
    L.heroam_gpu_run.restype = C.c_int
    L.heroam_gpu_run.argtypes = [vp]
    L.heroam_gpu_download.restype = C.c_int
    L.heroam_gpu_download.argtypes = [vp, vp, vp]
    L.heroam_gpu_sample_batch.restype = C.c_int
    L.heroam_gpu_sample_batch.argtypes = [vp, C.c_ulonglong, C.c_ulonglong, C.c_int, C.POINTER(ParticlesC)]
    L.heroam_gpu_sample_flat.restype = C.c_int
    L.heroam_gpu_sample_flat.argtypes = [vp, C.c_ulonglong, C.c_ulonglong, C.c_int, u32p, vp, C.c_long, C.POINTER(C.c_long)]
    L.heroam_gpu_free_particles.restype = None
    L.heroam_gpu_free_particles.argtypes = [C.POINTER(ParticlesC)]
    L.heroam_gpu_run_stats.restype = C.c_int
    L.heroam_gpu_run_stats.argtypes = [vp, C.c_ulonglong, C.c_ulonglong, C.c_ulonglong, C.POINTER(Stats)]
    L.heroam_gpu_run_sweep.restype = C.c_int
    L.heroam_gpu_run_sweep.argtypes = [vp, C.c_ulonglong, C.c_int, C.POINTER(SweepPoint), C.POINTER(Stats)]
    L.heroam_gpu_run_size_sweep.restype = C.c_int
    L.heroam_gpu_run_size_sweep.argtypes = [vp, C.c_int, C.POINTER(SizePoint), C.POINTER(Stats)]
    L.heroam_gpu_last_run_info.restype = C.c_int
    L.heroam_gpu_last_run_info.argtypes = [vp, C.POINTER(RunInfo)]
    L.heroam_gpu_fp32_peak.restype = C.c_int
    L.heroam_gpu_fp32_peak.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.heroam_gpu_fp32_peak_rrr.restype = C.c_int
    L.heroam_gpu_fp32_peak_rrr.argtypes = [vp, C.POINTER(C.c_double)]
    L.heroam_gpu_fp32_peak_x2.restype = C.c_int
    L.heroam_gpu_fp32_peak_x2.argtypes = [vp, C.POINTER(C.c_double)]
    _lib = L
    return L


def device_count() -> int:
    return int(load_library().heroam_gpu_device_count())


class Engine:
    """One GPU context (``heroam_gpu_create``): force table resident on the device."""

    def __init__(self, conf: Config, table: np.ndarray, device: int = 0, mode="exact", segment_steps: int = 0,
                 max_steps: int = 0, wide_kernel: int = 0, pool_slots: int = 0, count_work: int = 0, no_flight: int = 0):
        self.lib = load_library()
        self.conf = conf
        table = np.ascontiguousarray(table, dtype=np.float32)
        if len(table) != conf.force_grid_length:
            raise ValueError(f"table has {len(table)} entries, config says {conf.force_grid_length}")
        opt = Options()
        opt.mode = _MODES[mode]
        opt.segment_steps = segment_steps
        opt.max_steps = max_steps
        opt.wide_kernel = wide_kernel
        opt.pool_slots = pool_slots
        opt.count_work = count_work
        opt.no_flight = no_flight
        h = C.c_void_p()
        self._check(self.lib.heroam_gpu_create(C.byref(conf), table.ctypes.data_as(C.POINTER(C.c_float)), device,
                                               C.byref(opt), C.byref(h)))
        self.handle = h
        self.device = device

    # -- plumbing -----------------------------------------------------------------
    def _check(self, rc: int) -> None:
        if rc != 0:
            raise HeroamError(f"heroam_gpu error {rc}: {self.lib.heroam_gpu_last_error().decode()}")

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.lib.heroam_gpu_destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_mode(self, mode) -> None:
        self._check(self.lib.heroam_gpu_set_mode(self.handle, _MODES[mode]))

    # -- the drop-in path ------------------------------------------------------------
    def simulate(self, he_number: int, counts: np.ndarray, flat: np.ndarray):
        """``heroam_gpu_simulate_flat``: host records in, final host records out."""
        counts = np.ascontiguousarray(counts, dtype=np.uint32)
        out = np.ascontiguousarray(flat.copy())
        assert out.dtype == PARTICLE_DTYPE and int(counts.sum()) == len(out)
        res = np.zeros(len(counts), dtype=RESULT_DTYPE)
        self._check(self.lib.heroam_gpu_simulate_flat(self.handle, he_number, counts.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                      out.ctypes.data, len(counts), res.ctypes.data))
        return out, res

    def simulate_traced(self, he_number: int, counts: np.ndarray, flat: np.ndarray, sink):
        """``heroam_gpu_simulate_traced`` (the reference's saveflag = 1): ``sink(trajectory, step,
        records)`` receives a copy of the records the reference would write to
        ``/Step_<step>/particles`` of that trajectory's file.  Returns (final records, results)."""
        counts = np.ascontiguousarray(counts, dtype=np.uint32)
        out = np.ascontiguousarray(flat.copy())
        assert out.dtype == PARTICLE_DTYPE and int(counts.sum()) == len(out)
        res = np.zeros(len(counts), dtype=RESULT_DTYPE)
        failure = []

        def _sink(_user, traj, step, records, no):
            try:
                buf = C.string_at(records, no * PARTICLE_DTYPE.itemsize)
                sink(int(traj), int(step), np.frombuffer(buf, dtype=PARTICLE_DTYPE))
                return 0
            except Exception as exc:  # propagate after the C call has unwound
                failure.append(exc)
                return 1

        cb = STEP_SINK(_sink)
        rc = self.lib.heroam_gpu_simulate_traced(self.handle, he_number, counts.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                 out.ctypes.data, len(counts), res.ctypes.data, cb, None)
        if failure:
            raise failure[0]
        self._check(rc)
        return out, res

    def simulate_inplace(self, he_number: int, counts: np.ndarray, flat: np.ndarray, results: np.ndarray) -> None:
        """``heroam_gpu_simulate_flat`` on caller-owned (e.g. page-locked) buffers, in place."""
        assert counts.dtype == np.uint32 and flat.dtype == PARTICLE_DTYPE and results.dtype == RESULT_DTYPE
        assert counts.flags.c_contiguous and flat.flags.c_contiguous and results.flags.c_contiguous
        self._check(self.lib.heroam_gpu_simulate_flat(self.handle, he_number, counts.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                      flat.ctypes.data, len(counts), results.ctypes.data))

    def reserve(self, n_trajectories: int, n_records: int) -> None:
        """``heroam_gpu_reserve``: size device buffers and staging ahead of the first large batch."""
        self._check(self.lib.heroam_gpu_reserve(self.handle, n_trajectories, n_records))

    def simulate_batch_inplace(self, he_number: int, particles: np.ndarray, results: np.ndarray) -> None:
        """``heroam_gpu_simulate_batch`` on an array of the reference's ``Particles`` headers
        (``PARTICLES_DTYPE``: no, index, pointer to that trajectory's records), in place."""
        assert particles.dtype == PARTICLES_DTYPE and particles.flags.c_contiguous
        assert results.dtype == RESULT_DTYPE and len(results) == len(particles)
        self._check(self.lib.heroam_gpu_simulate_batch(self.handle, he_number,
                                                       C.cast(particles.ctypes.data, C.POINTER(ParticlesC)),
                                                       len(particles), results.ctypes.data))

    def sample_into(self, he_number: int, count: int, first_index: int, counts: np.ndarray, flat: np.ndarray) -> int:
        """``heroam_gpu_sample_flat`` into caller-owned buffers; returns the number of records
        (raises if the buffer is too small)."""
        total = C.c_long(0)
        self._check(self.lib.heroam_gpu_sample_flat(self.handle, he_number, first_index, count,
                                                    counts.ctypes.data_as(C.POINTER(C.c_uint32)), flat.ctypes.data,
                                                    len(flat), C.byref(total)))
        if total.value > len(flat):
            raise HeroamError(f"sample_into: buffer of {len(flat)} records, {total.value} needed")
        self._n = (count, total.value)
        return int(total.value)

    def upload(self, he_number: int, counts: np.ndarray, flat: np.ndarray) -> None:
        counts = np.ascontiguousarray(counts, dtype=np.uint32)
        flat = np.ascontiguousarray(flat)
        self._n = (len(counts), len(flat))
        self._check(self.lib.heroam_gpu_upload(self.handle, he_number, counts.ctypes.data_as(C.POINTER(C.c_uint32)),
                                               flat.ctypes.data, len(counts)))

    def run(self) -> None:
        self._check(self.lib.heroam_gpu_run(self.handle))

    def download(self):
        n_traj, n_part = self._n
        out = np.zeros(n_part, dtype=PARTICLE_DTYPE)
        res = np.zeros(n_traj, dtype=RESULT_DTYPE)
        self._check(self.lib.heroam_gpu_download(self.handle, out.ctypes.data, res.ctypes.data))
        return out, res

    # -- sampling stage --------------------------------------------------------------
    def sample(self, he_number: int, count: int, first_index: int = 0):
        """``heroam_gpu_sample_flat`` -> (counts, flat records)."""
        counts = np.zeros(count, dtype=np.uint32)
        cap = max(64, count * 6)
        while True:
            flat = np.zeros(cap, dtype=PARTICLE_DTYPE)
            total = C.c_long(0)
            self._check(self.lib.heroam_gpu_sample_flat(self.handle, he_number, first_index, count,
                                                        counts.ctypes.data_as(C.POINTER(C.c_uint32)), flat.ctypes.data,
                                                        cap, C.byref(total)))
            if total.value <= cap:
                self._n = (count, total.value)
                return counts, flat[: total.value].copy()
            cap = total.value

    # -- fused statistics -------------------------------------------------------------
    def run_sweep(self, he_number: int, points) -> list:
        """``heroam_gpu_run_sweep``: points = [(intensity, pulse_width, count, first_index), ...];
        one statistics dict per point."""
        arr = (SweepPoint * len(points))()
        for a, (intensity, pulse_width, count, first_index) in zip(arr, points):
            a.intensity, a.pulse_width, a.count, a.first_index = intensity, pulse_width, count, first_index
        out = (Stats * len(points))()
        self._check(self.lib.heroam_gpu_run_sweep(self.handle, he_number, len(points), arr, out))
        return [self._stats_dict(st) for st in out]

    def run_size_sweep(self, points) -> list:
        """``heroam_gpu_run_size_sweep``: points = [(he_number, count, first_index), ...] -- the reference's
        droplet-size sweep (main.c:158-164) as one pool run; one statistics dict per size."""
        arr = (SizePoint * len(points))()
        for a, (he_number, count, first_index) in zip(arr, points):
            a.he_number, a.count, a.first_index = he_number, count, first_index
        out = (Stats * len(points))()
        self._check(self.lib.heroam_gpu_run_size_sweep(self.handle, len(points), arr, out))
        return [self._stats_dict(st) for st in out]

    def run_stats(self, he_number: int, count: int, first_index: int = 0) -> dict:
        st = Stats()
        self._check(self.lib.heroam_gpu_run_stats(self.handle, he_number, first_index, count, C.byref(st)))
        return self._stats_dict(st)

    @staticmethod
    def _stats_dict(st) -> dict:
        return dict(
            trajectories=st.trajectories,
            particles=st.particles,
            hist_n=np.array(st.hist_n[:], dtype=np.uint64),
            hist_traj_time=np.array(st.hist_traj_time[:], dtype=np.uint64),
            hist_meet_time=np.array(st.hist_meet_time[:], dtype=np.uint64),
            done=np.array(st.done[:], dtype=np.uint64),
            unmet_particles=st.unmet_particles,
            sum_traj_time=st.sum_traj_time,
            info=st.info.as_dict(),
        )

    def run_info(self) -> dict:
        info = RunInfo()
        self._check(self.lib.heroam_gpu_last_run_info(self.handle, C.byref(info)))
        return info.as_dict()

    def fp32_peak(self):
        tf, mhz = C.c_double(0), C.c_double(0)
        self._check(self.lib.heroam_gpu_fp32_peak(self.handle, C.byref(tf), C.byref(mhz)))
        return tf.value, mhz.value

    def fp32_peak_rrr(self) -> float:
        tf = C.c_double(0)
        self._check(self.lib.heroam_gpu_fp32_peak_rrr(self.handle, C.byref(tf)))
        return tf.value

    def fp32_peak_x2(self) -> float:
        tf = C.c_double(0)
        self._check(self.lib.heroam_gpu_fp32_peak_x2(self.handle, C.byref(tf)))
        return tf.value
