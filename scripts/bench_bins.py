"""Per-bin throughput of the integrate kernels (exploratory, GPU only).

Builds batches in which every trajectory has exactly n particles (sampled with the
engine's own sampler at a suitable intensity, replicated to fill the GPU), runs them
for a fixed number of steps and prints ns per particle-step and the FP32 roofline
fraction (SURVEY.md 8d flop model) for each n and mode.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from sphere_roaming_b200 import engine
from sphere_roaming_b200.forcetable import synthetic_table
from sphere_roaming_b200.types import make_config, offsets_of

TLARGE = int(os.environ.get("TLARGE", 0))  # 1: the 26.6 M-entry table (106 MB), gathers leave L1/L2
GRID, LENGTH = (6.0e-5, 26_600_000) if TLARGE else (0.001, 1_596_000)
table = synthetic_table(length=LENGTH, grid=GRID)
HE = int(os.environ.get("HE", 10000))
STEPS = int(os.environ.get("STEPS", 4096))
M = int(os.environ.get("M", 148 * 1024))
NS = [int(x) for x in os.environ.get("NS", "1,2,3,4,5,6,7,8,9,12,16,24,32,48,64").split(",")]
MODES = os.environ.get("MODES", "exact,fast").split(",")
SEG = int(os.environ.get("SEG", 0))


def batch_with_n(n, m):
    lam_per_intensity = 3.9337242843335494 * (HE / 10000.0)
    conf = make_config(number=4096, intensity=max(n, 1) / lam_per_intensity, max_time=1e7, force_grid=GRID, force_grid_length=LENGTH)
    with engine.Engine(conf, table) as eng:
        counts, flat = eng.sample(HE, 4096)
    off = offsets_of(counts)
    idx = np.nonzero(counts == n)[0]
    # keep trajectories without an initial contact so that all n particles move
    keep = [i for i in idx if not flat["success"][off[i]:off[i + 1]].any()]
    assert keep, f"no trajectory with n={n}"
    reps = -(-m // len(keep))
    recs = np.concatenate([flat[off[i]:off[i + 1]] for i in keep])
    recs = np.tile(recs, reps)[: m * n]
    return np.full(m, n, dtype=np.uint32), np.ascontiguousarray(recs)


peak = None
for n in NS:
    m = M if n <= 8 else (148 * 256 if n <= 16 else max(148 * 16, M // 32))
    counts, flat = batch_with_n(n, m)
    conf = make_config(number=m, max_time=1e7, force_grid=GRID, force_grid_length=LENGTH)
    for mode in MODES:
        with engine.Engine(conf, table, mode=mode, max_steps=STEPS, segment_steps=SEG) as eng:
            if peak is None:
                peak = eng.fp32_peak()[0]
                print(f"fp32 peak {peak:.1f} TFLOP/s")
            eng.upload(HE, counts, flat); eng.run()  # warm-up
            eng.upload(HE, counts, flat); eng.run()
            info = eng.run_info()
        ps, pr, ir = info["particle_steps"], info["pair_steps"], info["inrange_steps"]
        ms = info["integrate_ms"]
        flops = 100.0 * ps + 10.0 * pr + 14.0 * ir
        print(f"n={n:3d} {mode:5s} traj={m:7d} steps={ps / m / n:7.0f} integrate={ms:8.2f} ms  "
              f"{ms * 1e6 / ps:7.4f} ns/particle-step  {ps / ms / 1e6:8.2f} G ps/s  "
              f"in-range/pair={ir / max(pr, 1):.3f}  {flops / ms / 1e9:6.2f} TFLOP/s = {flops / ms / 1e9 / peak:5.3f} of peak  passes={info['passes']}",
              flush=True)
