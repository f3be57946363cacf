#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json: particle-steps/s and
trajectories/s on 1/2/4/8 B200 vs the reference's OpenMP CPU build).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (config.workload): BASELINE.json configs[1] -- the repository config.conf physics
(N_He = 1e4, lambda = 3.93 excitations), T-ref synthetic force table of the reference's
shape (1 596 000 entries), max_time = 1e7 fs -- streamed in batches: ONE STEP = one batch
of `--batch` trajectories per GPU taken from the global Philox index space, run through
sample -> integrate -> histogram to completion.  Successive steps use successive index
blocks, so K steps are K different batches of the 1e6-trajectory run.

  value  : whole-job particle-steps/s, state resident on the device (heroam_gpu_run_stats:
           only histograms leave the GPU), CUDA-event timed inside the library on the
           launching stream, max over ranks.
  e2e    : the same metric through the reference-shaped C-ABI calls with HOST buffers:
           heroam_gpu_sample_flat (initial records -> pinned host memory, the reference's
           particlefile_<N> export) then heroam_gpu_simulate_flat (host records -> device ->
           final host records, the reference's particlefile_after_<N>), wall clock of the
           calls between device synchronisations.
  --impl reference : the reference's own C code (oracle/_ref, its release flags, OpenMP on
           every host core) on a bounded sample of the same workload, rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from sphere_roaming_b200.forcetable import synthetic_table  # noqa: E402
from sphere_roaming_b200.types import PARTICLE_DTYPE, make_config  # noqa: E402

HE = 10000
METRIC = "particle_steps_per_sec"
UNIT = "particle-steps/s"


def workload_config(args):
    return {
        "workload": "C2: repo config.conf physics (N_He=1e4, I=1 GW/cm2, tau=100 fs, lambda=3.93), "
                    "T-ref synthetic force table 1596000 x f32 on r^2, dt=1 fs, max_time=%g fs, "
                    "seed 100401452; one step = one run of `batch_per_gpu` trajectories per GPU (1e6 = the C2 run), "
                    "the K steps streamed back to back through the engine's retire-and-refill pool" % args.max_time,
        "batch_per_gpu": args.batch,
        "max_time": args.max_time,
        "mode": args.mode,
        "l2": "flushed between steps (256 MiB write); state lives in registers, table (6.4 MB) in L1/L2",
    }


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# the reference arm / cpu baseline
# ------------------------------------------------------------------------------------------
def pick_ref_oracle():
    """The reference's own object code, fastest flavour that runs on this host."""
    from oracle.binding import RefOracle, ref_available

    for flavour in ("release", "strict"):
        if not ref_available(flavour):
            continue
        probe = subprocess.run([sys.executable, "-c",
                                "import sys; sys.path.insert(0, %r); from oracle.binding import RefOracle; "
                                "RefOracle(%r).sizes()" % (ROOT, flavour)], capture_output=True)
        if probe.returncode == 0:
            return RefOracle(flavour), "reference", flavour
    return None, "port", "port"


def cpu_run(table, max_time: float, n_traj: int, seed_offset: int = 0, dynamic: bool = True):
    """n_traj trajectories of the workload on all host cores with the reference's code:
    initialize_particles (libc rand) + the OpenMP loop over simulate_particles,
    schedule(dynamic,1) (faster than the as-written static schedule on a heavy-tailed
    workload).  Returns (particle_steps, seconds, trajectories, cores, kind)."""
    ref, kind, flavour = pick_ref_oracle()
    conf = make_config(number=n_traj, max_time=max_time, seed=100401452 + seed_offset)
    # every host core, whatever OMP_NUM_THREADS says (torchrun sets it to 1)
    all_cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if ref is not None:
        cores = max(ref.max_threads(), all_cores)
        counts, init = ref.initialize(conf, HE, table)
        t0 = time.perf_counter()
        final = ref.simulate(conf, HE, table, counts, init, nthreads=cores, dynamic=dynamic)
        dt = time.perf_counter() - t0
    else:
        from oracle.binding import PortOracle

        port = PortOracle()
        cores = max(port.max_threads(), all_cores)
        counts, init = port.sample(conf, HE, table, n_traj, source="rand")
        t0 = time.perf_counter()
        final, _ = port.simulate(conf, HE, table, counts, init, nthreads=cores)
        dt = time.perf_counter() - t0
    # particle-steps = sum of particle clocks / dt, dt = 1 fs
    psteps = float(final["time"].astype(np.float64).sum())
    return psteps, dt, n_traj, cores, kind, flavour


def reference_arm(args, table):
    cores = os.cpu_count() or 1
    sample = max(8 * cores, 32) if args.ref_sample <= 0 else args.ref_sample
    for w in range(args.warmup):
        cpu_run(table, args.max_time, min(sample, cores), seed_offset=1000 + w)
    tot_ps, tot_s, tot_traj = 0.0, 0.0, 0
    kind = flavour = None
    for k in range(args.steps):
        ps, s, n, cores, kind, flavour = cpu_run(table, args.max_time, sample, seed_offset=k)
        tot_ps += ps; tot_s += s; tot_traj += n
    value = tot_ps / tot_s
    # the reference's loop as written has no schedule clause (main.c:76: static); report it too, on one sample
    ps_static, s_static, _, _, _, _ = cpu_run(table, args.max_time, sample, seed_offset=0, dynamic=False)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args),
        "trajectories_per_sec": tot_traj / tot_s,
        "as_written_static_schedule": {"value": ps_static / s_static, "unit": UNIT,
                                       "note": "same code with the reference's own `#pragma omp parallel for` (no schedule clause), "
                                               "one sample; the headline uses schedule(dynamic,1), the faster of the two"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{sample} trajectories per step x {args.steps} steps, oracle/_ref '{flavour}' build "
                                   f"(reference simulation.c+vector.c unmodified, -O3 -ffast-math -funroll-loops -fopenmp), "
                                   f"schedule(dynamic,1), {cores} OpenMP threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# the GPU arm
# ------------------------------------------------------------------------------------------
def flop_model(info):
    """100 flops per particle-step, 10 per pair-step, +14 per in-range pair-step."""
    return 100.0 * info["particle_steps"] + 10.0 * info["pair_steps"] + 14.0 * info["inrange_steps"]


def executed_flop_model(info):
    """The same model charged only for what the kernels evaluate: pairs of a free-flight segment
    (provably force-free) are skipped, and a free-flight particle-step is the orientation update
    alone (41 of the 100 flops: no kicks, no position, no stored velocity)."""
    full = info["particle_steps"] - info["free_steps"]
    return (100.0 * full + 41.0 * info["free_steps"] + 10.0 * (info["pair_steps"] - info["pair_skipped"])
            + 14.0 * info["inrange_steps"])


def gpu_arm(args, table):
    import torch
    import torch.distributed as dist

    from sphere_roaming_b200 import engine
    from sphere_roaming_b200.sharding import reduce_stats

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    B = args.batch
    conf = make_config(number=B, max_time=args.max_time)
    eng = engine.Engine(conf, table, device=local, mode=args.mode, segment_steps=args.segment)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peak_tflops, clock_mhz = eng.fp32_peak()
    peak_rrr_tflops = eng.fp32_peak_rrr()

    K, W = args.steps, args.warmup

    def block(start_step: int, n_steps: int):
        """(first index, count) of this rank's share of n_steps steps starting at global
        step start_step; weak scaling: every step consumes world * B consecutive indices."""
        return (start_step * world + rank * n_steps) * B, n_steps * B

    # ---- resident arm: sample -> integrate -> histogram on the device -----------------------
    # The K steps are handed to the engine as one streaming run (retire-and-refill pool):
    # the drain of the heavy tail of trajectory lengths is paid once per run, as in production.
    if W > 0:
        f, n = block(0, W)
        eng.run_stats(HE, n, first_index=f)
    sampler = ClockSampler(local)
    flush.fill_(1)
    barrier()
    sampler.start()
    t_wall0 = time.perf_counter()
    f, n = block(W, K)
    merged = eng.run_stats(HE, n, first_index=f)
    work = dict(merged["info"])
    merged = reduce_stats(merged, device=dev)  # the run's one collective
    barrier()
    wall_s = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    dev_s = max_over_ranks(work["device_ms"] * 1e-3)
    wall_s = max_over_ranks(wall_s)
    tot_ps = sum_over_ranks(float(work["particle_steps"]))
    tot_traj = sum_over_ranks(float(work["trajectories"]))
    tot_flops = sum_over_ranks(executed_flop_model(work))
    tot_flops_ref = sum_over_ranks(flop_model(work))
    tot_free = sum_over_ranks(float(work["free_steps"]))
    integ_s = max_over_ranks(work["integrate_ms"] * 1e-3)
    tot_launches = int(sum_over_ranks(float(work["kernel_launches"])))

    # ---- end-to-end arm: host buffers through the reference-shaped calls ------------------------
    # Same K steps, through heroam_gpu_sample_flat (initial records device -> host: the reference's
    # particlefile_<N> export) and heroam_gpu_simulate_flat (host -> device, integrate, final records
    # device -> host: particlefile_after_<N>), on page-locked host buffers allocated beforehand.
    e2e_steps = max(1, min(K, args.e2e_steps if args.e2e_steps > 0 else K))
    n_e2e = e2e_steps * B
    cap = int(n_e2e * 4.2) + 4096  # E[n] = 4.01 at lambda = 3.93
    pin_flat = torch.empty(cap * PARTICLE_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True)
    pin_counts = torch.empty(n_e2e * 4, dtype=torch.uint8, pin_memory=True)
    pin_res = torch.empty(n_e2e * 16, dtype=torch.uint8, pin_memory=True)
    h_flat = pin_flat.numpy().view(PARTICLE_DTYPE)
    h_counts = pin_counts.numpy().view(np.uint32)
    h_res = pin_res.numpy().view(engine.RESULT_DTYPE)
    e2e_ps = e2e_s = 0.0
    h2d = d2h = 0
    for rep in range(args.warmup_e2e + 1):
        timed = rep == args.warmup_e2e
        n = n_e2e if timed else min(n_e2e, B)
        f, _ = block(10000 + rep * e2e_steps, 1)
        flush.fill_(2)
        barrier()
        t0 = time.perf_counter()
        total = eng.sample_into(HE, n, f, h_counts[:n], h_flat)                 # device -> host
        eng.simulate_inplace(HE, h_counts[:n], h_flat[:total], h_res[:n])        # host -> device -> host
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if timed:
            info = eng.run_info()
            e2e_ps, e2e_s = float(info["particle_steps"]), dt
            h2d = info["h2d_bytes"]
            d2h = info["d2h_bytes"] + total * PARTICLE_DTYPE.itemsize + n * 4
    e2e_s = max_over_ranks(e2e_s)
    e2e_total_ps = sum_over_ranks(e2e_ps)

    # ---- in-range share of the evaluated pairs (the fast kernels do not count it while timed) ----
    with engine.Engine(conf, table, device=local, mode=args.mode, segment_steps=args.segment, count_work=1) as eng_c:
        ci = eng_c.run_stats(HE, min(B, 32768), first_index=block(30000, 1)[0])["info"]
    inrange_ratio = ci["inrange_steps"] / max(1, ci["pair_steps"] - ci["pair_skipped"])
    if work["inrange_steps"] < 0.5 * inrange_ratio * (work["pair_steps"] - work["pair_skipped"]):
        work["inrange_steps"] = int(inrange_ratio * (work["pair_steps"] - work["pair_skipped"]))
        tot_flops = sum_over_ranks(executed_flop_model(work))
        tot_flops_ref = sum_over_ranks(flop_model(work))

    # ---- exact mode, one batch, for the record ----------------------------------------------
    exact = None
    if args.mode != "exact" and args.exact_step:
        eng.set_mode("exact")
        eng.run_stats(HE, min(B, 8192), first_index=block(20000, 1)[0])
        st = eng.run_stats(HE, min(B, 262144), first_index=block(W, 1)[0])
        exact = {"value": sum_over_ranks(float(st["info"]["particle_steps"])) / max_over_ranks(st["info"]["device_ms"] * 1e-3),
                 "unit": UNIT, "note": "HEROAM_MODE_EXACT (bit-exact with the strict reference), one run of min(batch, 262144) trajectories (drain dominated)"}
        eng.set_mode(args.mode)

    # ---- relaxed fast mode (renormalisation every 8th step), same K steps, for the record ----------
    relaxed = None
    if args.mode == "fast" and args.relaxed_step:
        eng.set_mode("fast_relaxed")
        f, n = block(W, K)
        flush.fill_(3)
        barrier()
        st = eng.run_stats(HE, n, first_index=f)
        relaxed = {"value": sum_over_ranks(float(st["info"]["particle_steps"])) / max_over_ranks(st["info"]["device_ms"] * 1e-3),
                   "unit": UNIT, "trajectories_per_sec": sum_over_ranks(float(st["info"]["trajectories"])) /
                   max_over_ranks(st["info"]["device_ms"] * 1e-3),
                   "note": "HEROAM_MODE_FAST_RELAXED (opt-in): same K steps with the orientation renormalised after every 8th "
                           "step; positions within ~2e-3 A of the strict reference after 2e5 steps instead of ~2e-5 A"}
        eng.set_mode(args.mode)

    if rank == 0:
        traffic = None
        prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        achieved = tot_flops / integ_s / 1e12
        line = {
            "metric": METRIC, "value": tot_ps / dev_s, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args),
            "trajectories_per_sec": tot_traj / dev_s,
            "wall_ms_per_step": 1e3 * wall_s / args.steps,
            "clocks": clocks,
            "e2e": {"value": e2e_total_ps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d // e2e_steps,
                    "d2h_bytes_per_step": d2h // e2e_steps, "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                    "api": "heroam_gpu_sample_flat + heroam_gpu_simulate_flat, host records in and out"},
            "gpu_launches": tot_launches,
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": peak_tflops * world, "unit": "TFLOP/s",
                         "frac": achieved / (peak_tflops * world), "traffic": traffic,
                         "peak_three_register_ffma": peak_rrr_tflops * world,
                         "frac_of_three_register_peak": achieved / (peak_rrr_tflops * world),
                         "frac_of_reference_work": tot_flops_ref / integ_s / 1e12 / (peak_tflops * world),
                         "free_flight_share_of_particle_steps": tot_free / tot_ps,
                         "inrange_share_of_evaluated_pairs": inrange_ratio,
                         "model": "flops actually evaluated: 100 per full particle-step, 41 per free-flight particle-step, 10 per "
                                  "evaluated pair-step, +14 per in-range pair-step, over the CUDA-event time of the integrate "
                                  "kernels (frac_of_reference_work charges the SURVEY.md 8d model 100/10/14 for every step and "
                                  "pair the reference would evaluate); peak = register-resident FFMA "
                                  "microbenchmark measured in this run (MEASURED_PEAKS.json has no FP32 figure), "
                                  "multiplier in a uniform register; peak_three_register_ffma = the same chains with "
                                  "multiplier and addend in ordinary registers, the form the kernels' FFMAs have",
                         "integrate_share_of_step": integ_s / dev_s},
            "exact_mode": exact,
            "relaxed_mode": relaxed,
            "outcomes": None if merged is None else {"success": int(merged["done"][0]), "max_time": int(merged["done"][1]),
                                                     "deadlock": int(merged["done"][2]), "trajectories": int(merged["trajectories"])},
        }
        if world == 1 and not args.no_cpu:
            ps, s, n, cores, kind, flavour = cpu_run(table, args.max_time, max(16 * (os.cpu_count() or 1), 64))
            line["cpu_baseline"] = {"value": ps / s, "unit": UNIT, "cores": cores, "kind": kind,
                                    "trajectories_per_sec": n / s,
                                    "sample": f"{n} trajectories of the same workload (libc-rand initial states), "
                                              f"oracle/_ref '{flavour}' build, schedule(dynamic,1), {cores} threads, {s:.1f} s"}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1000000, help="trajectories per GPU per step (C2: 1e6)")
    ap.add_argument("--max-time", type=float, default=1.0e7)
    ap.add_argument("--mode", default="fast", choices=["fast", "exact", "fast_relaxed"])
    ap.add_argument("--segment", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end arm (0 = all K)")
    ap.add_argument("--warmup-e2e", type=int, default=1)
    ap.add_argument("--exact-step", type=int, default=1)
    ap.add_argument("--relaxed-step", type=int, default=1)
    ap.add_argument("--ref-sample", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    table = synthetic_table()
    if args.impl == "reference":
        if int(os.environ.get("RANK", 0)) == 0:
            reference_arm(args, table)
        return
    gpu_arm(args, table)


if __name__ == "__main__":
    main()
