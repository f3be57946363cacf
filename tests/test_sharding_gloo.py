"""The N>1 path on CPU: world_size-2 gloo.  Each rank owns a block of the global
trajectory index space, runs it (here with the CPU oracle standing in for the device),
and the per-rank statistics are summed by the one reduce the real multi-GPU run uses.
The result must equal the single-process run: sharding must not change anything."""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from sphere_roaming_b200.sharding import pack_stats, shard_range, stats_from_records, unpack_stats


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 1000, 10**8 + 3):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_range(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == total
            for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_shard_size_sweep_partitions_every_point():
    from sphere_roaming_b200.sharding import shard_size_sweep

    points = [(6000, 2500, 0), (10000, 3000, 50), (16000, 7, 7), (30000, 0, 123456789012)]
    for world in (1, 2, 3, 8):
        shards = [shard_size_sweep(points, r, world) for r in range(world)]
        for p, (he, count, first) in enumerate(points):
            at = first
            for s in shards:
                assert s[p][0] == he and s[p][2] == at  # contiguous, in rank order, size unchanged
                at += s[p][1]
            assert at == first + count


def _local_stats(first, count, total_conf_number, max_time):
    from oracle.binding import PortOracle
    from sphere_roaming_b200.forcetable import synthetic_table
    from sphere_roaming_b200.types import make_config

    table = synthetic_table()
    port = PortOracle()
    conf = make_config(number=total_conf_number, max_time=max_time)
    counts, init = port.sample(conf, 10000, table, count, source="philox", first_index=first)
    final, res = port.simulate(conf, 10000, table, counts, init, nthreads=2)
    return stats_from_records(counts, final, res, max_time)


def _worker(rank, world, port_no, total, max_time, out_path):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    from sphere_roaming_b200.sharding import reduce_stats

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard_range(total, rank, world)
    stats = _local_stats(first, count, total, max_time)
    merged = reduce_stats(stats, dst=0)
    dist.barrier()
    if rank == 0:
        np.savez(out_path, packed=pack_stats(merged), tsum=merged["sum_traj_time"])
    dist.destroy_process_group()


def test_two_ranks_equal_one(tmp_path):
    import torch.multiprocessing as mp

    total, max_time = 96, 2.0e4
    out = str(tmp_path / "merged.npz")
    port_no = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port_no, total, max_time, out), nprocs=2, join=True)
    merged = np.load(out)
    single = _local_stats(0, total, total, max_time)
    assert np.array_equal(merged["packed"], pack_stats(single))
    assert abs(float(merged["tsum"]) - single["sum_traj_time"]) < 1e-6 * max(1.0, single["sum_traj_time"])
    # round trip of the packing
    again = unpack_stats(pack_stats(single), single["sum_traj_time"], single)
    assert np.array_equal(again["hist_n"], single["hist_n"]) and again["trajectories"] == total
