"""Config C4 on N GPUs of one node: every point's trajectories sharded by index over the ranks
(sharding.shard_sweep), one pool run per rank (heroam_gpu_run_sweep), one reduce per point.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/run_c4_dist.py
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from sphere_roaming_b200 import engine
from sphere_roaming_b200.forcetable import synthetic_table
from sphere_roaming_b200.sharding import reduce_stats, shard_sweep
from sphere_roaming_b200.types import make_config

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N = int(os.environ.get("N", 100000)); MT = float(os.environ.get("MT", 1e7)); HE = 10000
points = [(i, w, N, 0) for i in (0.25, 0.5, 1.0, 2.0, 4.0) for w in (50.0, 100.0, 200.0)]
with engine.Engine(make_config(max_time=MT), synthetic_table(), device=local, mode="fast") as eng:
    eng.run_stats(HE, 4096)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    mine = eng.run_sweep(HE, shard_sweep(points, rank, world))
    merged = [reduce_stats(st, device=dev) for st in mine]  # one 68 KB reduce per point
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
if rank == 0:
    ps = merged[0]["info"]["particle_steps"]
    print("C4 on %d GPU(s): %d points x %d trajectories in %.2f s -> %.3g particle-steps/s, %.0f trajectories/s" % (
        world, len(points), N, dt, ps / dt, len(points) * N / dt))
    for p, st in zip(points, merged):
        print("  I=%.2f tau=%3.0f  trajectories %d  mean n %.2f  success %d  max_time %d" % (
            p[0], p[1], st["trajectories"], st["particles"] / max(1, st["trajectories"]), st["done"][0], st["done"][1]))
if world > 1:
    dist.destroy_process_group()
