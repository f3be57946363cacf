"""Multi-GPU sharding of the trajectory index space.

Trajectories are independent (reference main.c:76-78) and the Philox stream is keyed by
the GLOBAL trajectory index, so every rank simply owns a contiguous block of indices and
results do not depend on the number of ranks.  There is no data-path collective; the only
exchange is one sum of the histogram/counter block at the end of a run.
"""
from __future__ import annotations

import numpy as np

from .engine import HIST_N, HIST_T

#: integer entries of a statistics dict that are summed across ranks
SUM_KEYS = ("trajectories", "particles", "unmet_particles")
HIST_KEYS = ("hist_n", "hist_traj_time", "hist_meet_time", "done")
#: work counters of the run info that add up across ranks (passes and the timings stay the rank's own)
INFO_SUM_KEYS = ("trajectories", "particles", "particle_steps", "pair_steps", "inrange_steps", "pair_skipped",
                 "free_steps", "kernel_launches")


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """(first index, count) of rank's block: [rank*total//world, (rank+1)*total//world)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    first = rank * total // world
    last = (rank + 1) * total // world
    return first, last - first


def shard_sweep(points, rank: int, world: int):
    """A rank's share of a laser sweep: every point (intensity, pulse_width, count, first_index)
    keeps its place, its trajectories are split by index like a single run's."""
    out = []
    for intensity, pulse_width, count, first_index in points:
        first, n = shard_range(count, rank, world)
        out.append((intensity, pulse_width, n, first_index + first))
    return out


def shard_size_sweep(points, rank: int, world: int):
    """A rank's share of a droplet-size sweep (reference main.c:158-164; Engine.run_size_sweep):
    every point (he_number, count, first_index) keeps its place and its size, its trajectories
    are split by index like a single run's (host/heroam_main.c stats_for_sweep does the same per GPU)."""
    out = []
    for he_number, count, first_index in points:
        first, n = shard_range(count, rank, world)
        out.append((he_number, n, first_index + first))
    return out


def stats_from_records(counts: np.ndarray, flat: np.ndarray, results: np.ndarray, max_time: float) -> dict:
    """Host-side equivalent of the device histogram kernel (k_stats), for batches whose
    final records were brought back to the host (the drop-in simulate path)."""

    def tbin(t):
        b = (t.astype(np.float32) / np.float32(max_time) * np.float32(HIST_T)).astype(np.int64)
        return np.clip(b, 0, HIST_T - 1)

    met = flat["success"] != 0
    return dict(
        trajectories=int(len(counts)),
        particles=int(len(flat)),
        hist_n=np.bincount(np.minimum(counts, HIST_N - 1), minlength=HIST_N).astype(np.uint64),
        hist_traj_time=np.bincount(tbin(results["time"]), minlength=HIST_T).astype(np.uint64),
        hist_meet_time=np.bincount(tbin(flat["time"][met]), minlength=HIST_T).astype(np.uint64),
        done=np.bincount(np.minimum(results["status"], 7), minlength=8).astype(np.uint64),
        unmet_particles=int((~met).sum()),
        sum_traj_time=float(results["time"].astype(np.float64).sum()),
    )


def pack_stats(stats: dict) -> np.ndarray:
    """Flatten the summable part of a statistics dict into one int64 vector (~8.5k words,
    68 KB: the whole inter-GPU exchange of a run)."""
    parts = [np.array([stats[k] for k in SUM_KEYS], dtype=np.int64)]
    parts += [np.asarray(stats[k]).astype(np.int64) for k in HIST_KEYS]
    info = stats.get("info") or {}
    parts.append(np.array([info.get(k, 0) for k in INFO_SUM_KEYS], dtype=np.int64))
    return np.concatenate(parts)


def unpack_stats(vec: np.ndarray, sum_traj_time: float, template: dict) -> dict:
    out = dict(template)
    at = 0
    for k in SUM_KEYS:
        out[k] = int(vec[at]); at += 1
    for k in HIST_KEYS:
        n = len(template[k])
        out[k] = vec[at:at + n].astype(np.uint64); at += n
    info = dict(template.get("info") or {})
    for k in INFO_SUM_KEYS:
        info[k] = int(vec[at]); at += 1
    out["info"] = info
    out["sum_traj_time"] = float(sum_traj_time)
    return out


def reduce_stats(stats: dict, device=None, dst: int = 0) -> dict:
    """Sum per-rank statistics onto rank `dst` with ONE torch.distributed reduce (NCCL
    over NVLink on GPUs, gloo in the CPU tests).  Non-destination ranks get their own
    partial back.  Without an initialised process group this is the identity."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return stats
    vec = torch.from_numpy(pack_stats(stats))
    tim = torch.tensor([stats["sum_traj_time"]], dtype=torch.float64)
    if device is not None:
        vec, tim = vec.to(device), tim.to(device)
    # one message: the float64 sum travels bit-cast inside the int64 vector? no -- a sum of
    # doubles must be added as doubles; it rides in a second 8-byte reduce.
    dist.reduce(vec, dst=dst, op=dist.ReduceOp.SUM)
    dist.reduce(tim, dst=dst, op=dist.ReduceOp.SUM)
    if dist.get_rank() != dst:
        return stats
    return unpack_stats(vec.cpu().numpy(), float(tim.cpu()[0]), stats)
