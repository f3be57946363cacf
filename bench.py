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
           heroam_gpu_sample_flat (initial records -> host memory, the reference's
           particlefile_<N> export) then heroam_gpu_simulate_batch -- the drop-in for the loop
           at main.c:76-78 that INTEGRATION.md tells a maintainer to call, on the reference's
           own pointer-per-trajectory Particles[] layout (host records -> device -> final host
           records, the reference's particlefile_after_<N>) -- wall clock of the calls between
           device synchronisations.
  extra keys (same JSON line): c2_single_run (ONE run of 1e6 trajectories, drain included),
           c3, c4_sweep and c5_slice (BASELINE.json configs[2..4]; the last two are FIXED jobs
           split over the ranks, i.e. strong scaling, summarised again under strong_scaling),
           t_large (the 106 MB table: what the gather costs once it leaves L1), each with
           particle-steps/s, trajectories/s and its own roofline fraction.
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
_REF_CHOICE = None


def pick_ref_oracle():
    """The reference's own object code, the fastest flavour that runs on this host: `native` is the
    reference's own -march=native (meson.build:19-21) but was compiled on another machine and may not
    run here; `release` is the same flags with -march=x86-64-v3.  Every candidate is first run in a
    subprocess (an illegal instruction must not take the benchmark down), then timed on a small sample."""
    global _REF_CHOICE
    if _REF_CHOICE is not None:
        return _REF_CHOICE
    from oracle.binding import RefOracle, ref_available

    probe_src = ("import sys, time; sys.path.insert(0, %r); import numpy as np; "
                 "from oracle.binding import RefOracle; from sphere_roaming_b200.forcetable import synthetic_table; "
                 "from sphere_roaming_b200.types import make_config; "
                 "r = RefOracle(%r); t = synthetic_table(); c = make_config(number=64, max_time=2.0e5); "
                 "k, i = r.initialize(c, 10000, t); t0 = time.perf_counter(); f, _ = r.simulate_guarded(c, 10000, t, k, i, nthreads=0, dynamic=True); "
                 "print(float(f['time'].astype('f8').sum()) / (time.perf_counter() - t0))")
    tried = {}
    for flavour in ("native", "release", "strict"):
        if not ref_available(flavour):
            continue
        env = dict(os.environ)
        env.pop("OMP_NUM_THREADS", None)
        probe = subprocess.run([sys.executable, "-c", probe_src % (ROOT, flavour)], capture_output=True, text=True, env=env)
        if probe.returncode == 0:
            try:
                tried[flavour] = float(probe.stdout.strip().splitlines()[-1])
            except (ValueError, IndexError):
                pass
        else:
            tried[flavour + " (does not run here)"] = 0.0
    usable = {k: v for k, v in tried.items() if v > 0.0 and k != "strict"} or {k: v for k, v in tried.items() if v > 0.0}
    if usable:
        best = max(usable, key=usable.get)
        _REF_CHOICE = (RefOracle(best), "reference", best, tried)
    else:
        _REF_CHOICE = (None, "port", "port", tried)
    return _REF_CHOICE


def cpu_run(table, max_time: float, n_traj: int, seed_offset: int = 0, dynamic: bool = True):
    """n_traj trajectories of the workload on all host cores with the reference's code:
    initialize_particles (libc rand) + the OpenMP loop over simulate_particles,
    schedule(dynamic,1) (faster than the as-written static schedule on a heavy-tailed
    workload).  Returns (particle_steps, seconds, trajectories, cores, kind)."""
    ref, kind, flavour, _ = pick_ref_oracle()
    conf = make_config(number=n_traj, max_time=max_time, seed=100401452 + seed_offset)
    # every host core, whatever OMP_NUM_THREADS says (torchrun sets it to 1)
    all_cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if ref is not None:
        cores = max(ref.max_threads(), all_cores)
        counts, init = ref.initialize(conf, HE, table)
        t0 = time.perf_counter()
        # guarded: the reference never returns from a trajectory whose particles are all frozen with its `==`
        # success criterion unmet (~7 in 1e5 trajectories); oracle/ref_harness.c steps out of that
        # loop.  Same object code, same results, no measurable cost (one no-op call per iteration).
        final, _dead = ref.simulate_guarded(conf, HE, table, counts, init, nthreads=cores, dynamic=dynamic)
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
                                   f"schedule(dynamic,1), {cores} OpenMP threads",
                         "flavours_probed": pick_ref_oracle()[3]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# the GPU arm
# ------------------------------------------------------------------------------------------
#: flops of one free-flight particle-step as the fast kernels execute it: the increment w' (x) q
#: (4 x (mul + 2 fma + add) = 24), the renormalisation (|q|^2: mul + 3 fma = 7, rsqrt 1, 4 mul) and the
#: clock (1); the relaxed mode renormalises after every 8th step only
FREE_STEP_FLOPS = {"fast": 37.0, "fast_relaxed": 24.0 + 1.0 + 12.0 / 8.0, "exact": 37.0, "fast_closed": 0.0}


def flop_model(info):
    """100 flops per particle-step, 10 per pair-step, +14 per in-range pair-step."""
    return 100.0 * info["particle_steps"] + 10.0 * info["pair_steps"] + 14.0 * info["inrange_steps"]


def executed_flop_model(info, mode="fast"):
    """The same model charged only for what the kernels evaluate: pairs of a free-flight segment
    (provably force-free) are skipped, and a free-flight particle-step is the orientation update
    alone (FREE_STEP_FLOPS of the 100: no kicks, no position, no stored velocity)."""
    full = info["particle_steps"] - info["free_steps"]
    return (100.0 * full + FREE_STEP_FLOPS.get(mode, 37.0) * info["free_steps"]
            + 10.0 * (info["pair_steps"] - info["pair_skipped"]) + 14.0 * info["inrange_steps"])


C4_POINTS = [(i, w) for i in (0.25, 0.5, 1.0, 2.0, 4.0) for w in (50.0, 100.0, 200.0)]


def gpu_arm(args, table):
    import torch
    import torch.distributed as dist

    from sphere_roaming_b200 import engine
    from sphere_roaming_b200.sharding import reduce_stats, shard_range, shard_sweep

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

    def inrange_share(engine_kwargs, he, n_small, first):
        """In-range share of the evaluated pairs of a workload (the fast kernels do not count it while
        timed): a small run of the same workload with the counting variant of the kernels."""
        with engine.Engine(device=local, mode=args.mode, count_work=1, **engine_kwargs) as eng_c:
            ci = eng_c.run_stats(he, n_small, first_index=first)["info"]
        return ci["inrange_steps"] / max(1, ci["pair_steps"] - ci["pair_skipped"])

    def summarise(work, ratio, mode, n_traj_total=None):
        """Whole-job figures of one measured run from the per-rank run info: device time = max over
        ranks, work = sum over ranks; roofline fraction over the DEVICE time (first kernel to last)."""
        w = dict(work)
        if w["inrange_steps"] < 0.5 * ratio * (w["pair_steps"] - w["pair_skipped"]):
            w["inrange_steps"] = int(ratio * (w["pair_steps"] - w["pair_skipped"]))
        dev_s = max_over_ranks(w["device_ms"] * 1e-3)
        ps = sum_over_ranks(float(w["particle_steps"]))
        traj = sum_over_ranks(float(w["trajectories"]))
        flops = sum_over_ranks(executed_flop_model(w, mode))
        return {"value": ps / dev_s, "unit": UNIT, "trajectories_per_sec": traj / dev_s, "seconds": dev_s,
                "trajectories": int(traj), "particle_steps": ps,
                "roofline_frac": flops / dev_s / 1e12 / (peak_tflops * world),
                "free_flight_share_of_particle_steps": sum_over_ranks(float(w["free_steps"])) / max(ps, 1.0),
                "passes": int(w["passes"])}

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
    wall_s = max_over_ranks(wall_s)
    tot_launches = int(sum_over_ranks(float(work["kernel_launches"])))
    c2_kwargs = dict(conf=conf, table=table, segment_steps=args.segment)
    inrange_ratio = inrange_share(c2_kwargs, HE, min(B, 32768), block(30000, 1)[0])
    head = summarise(work, inrange_ratio, args.mode)
    dev_s = head["seconds"]
    integ_s = max_over_ranks(work["integrate_ms"] * 1e-3)
    resolve_s = max_over_ranks(work["resolve_ms"] * 1e-3)
    w_ref = dict(work)
    if w_ref["inrange_steps"] < 0.5 * inrange_ratio * (w_ref["pair_steps"] - w_ref["pair_skipped"]):
        w_ref["inrange_steps"] = int(inrange_ratio * (w_ref["pair_steps"] - w_ref["pair_skipped"]))
    tot_flops_ref = sum_over_ranks(flop_model(w_ref))
    achieved = head["roofline_frac"] * peak_tflops * world

    # ---- end-to-end arm: host buffers through the reference-shaped calls ------------------------
    # Same K steps, through heroam_gpu_sample_flat (initial records device -> host: the reference's
    # particlefile_<N> export) and heroam_gpu_simulate_batch, the drop-in for the loop at main.c:76-78,
    # on the reference's own pointer-per-trajectory Particles[] (host -> device, integrate, final records
    # device -> host: particlefile_after_<N>).
    e2e_steps = max(1, min(K, args.e2e_steps if args.e2e_steps > 0 else K))
    n_e2e = e2e_steps * B
    cap = int(n_e2e * 4.2) + 4096  # E[n] = 4.01 at lambda = 3.93
    pin_flat = torch.empty(cap * PARTICLE_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True)
    h_flat = pin_flat.numpy().view(PARTICLE_DTYPE)
    h_counts = np.empty(n_e2e, dtype=np.uint32)
    h_res = np.empty(n_e2e, dtype=engine.RESULT_DTYPE)
    h_particles = np.empty(n_e2e, dtype=engine.PARTICLES_DTYPE)  # the reference's Particles[number] (main.c:55)
    e2e_ps = e2e_s = 0.0
    h2d = d2h = 0
    eng.reserve(n_e2e, cap)  # buffers sized once, as a long-running caller would: not part of a step
    for rep in range(args.warmup_e2e + 1):
        timed = rep == args.warmup_e2e
        n = n_e2e if timed else min(n_e2e, B)
        f, _ = block(10000 + rep * e2e_steps, 1)
        flush.fill_(2)
        barrier()
        t0 = time.perf_counter()
        total = eng.sample_into(HE, n, f, h_counts[:n], h_flat)                 # device -> host
        hp = h_particles[:n]
        hp["no"] = h_counts[:n]
        hp["index"] = np.arange(n, dtype=np.uint32)
        off = np.zeros(n, dtype=np.uint64)
        off[1:] = np.cumsum(h_counts[:n - 1], dtype=np.uint64)
        hp["particle"] = np.uint64(h_flat.ctypes.data) + off * np.uint64(PARTICLE_DTYPE.itemsize)
        eng.simulate_batch_inplace(HE, hp, h_res[:n])                           # host -> device -> host
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if timed:
            info = eng.run_info()
            e2e_ps, e2e_s = float(info["particle_steps"]), dt
            h2d = info["h2d_bytes"]
            d2h = info["d2h_bytes"] + total * PARTICLE_DTYPE.itemsize + n * 4
    e2e_s = max_over_ranks(e2e_s)
    e2e_total_ps = sum_over_ranks(e2e_ps)
    del pin_flat, h_flat, h_particles

    # ---- exact mode, one batch, for the record ----------------------------------------------
    exact = None
    if args.mode != "exact" and args.exact_step:
        eng.set_mode("exact")
        eng.run_stats(HE, min(B, 8192), first_index=block(20000, 1)[0])
        st = eng.run_stats(HE, min(B, 262144), first_index=block(W, 1)[0])
        exact = {"value": sum_over_ranks(float(st["info"]["particle_steps"])) / max_over_ranks(st["info"]["device_ms"] * 1e-3),
                 "unit": UNIT, "note": "HEROAM_MODE_EXACT (bit-exact with the strict reference), one run of min(batch, 262144) trajectories (drain dominated)"}
        eng.set_mode(args.mode)

    # ---- relaxed fast mode (renormalisation every 8th step), for the record ------------------------
    relaxed = None
    if args.mode == "fast" and args.relaxed_step:
        eng.set_mode("fast_relaxed")
        k_rel = K if args.relaxed_step < 0 else max(1, min(K, args.relaxed_step))
        f, n = block(W, k_rel)
        flush.fill_(3)
        barrier()
        st = eng.run_stats(HE, n, first_index=f)
        relaxed = summarise(st["info"], inrange_ratio, "fast_relaxed")
        relaxed["note"] = (f"HEROAM_MODE_FAST_RELAXED (opt-in): the same index block as the headline ({k_rel} steps as one run) with the "
                           "orientation renormalised after every 8th step; its own position tolerance (DESIGN.md section 4)")
        eng.set_mode(args.mode)

    # ---- closed-form free flight (opt-in), for the record -------------------------------------------
    closed = None
    if args.mode == "fast" and args.closed_step:
        eng.set_mode("fast_closed")
        k_cl = K if args.closed_step < 0 else max(1, min(K, args.closed_step))
        f, n = block(W, k_cl)
        flush.fill_(3)
        barrier()
        st = eng.run_stats(HE, n, first_index=f)
        closed = summarise(st["info"], inrange_ratio, "fast_closed")
        closed["note"] = (f"HEROAM_MODE_FAST_CLOSED (opt-in): the same index block as the headline ({k_cl} steps as one run) with every free flight taken as one "
                          "rotation instead of step by step; particle-steps count the steps advanced, not instructions -- "
                          "roofline_frac charges nothing for a flight step; its own tolerance row (DESIGN.md section 4)")
        eng.set_mode(args.mode)

    # ---- the other BASELINE.json configurations ------------------------------------------------------
    extras = {}
    if args.extras:
        # C2 as written: ONE run of 1e6 trajectories per GPU, its drain included
        flush.fill_(4)
        barrier()
        st = eng.run_stats(HE, args.c2_number, first_index=block(40000, 1)[0])
        extras["c2_single_run"] = summarise(st["info"], inrange_ratio, args.mode)
        extras["c2_single_run"]["config"] = (f"BASELINE configs[1] as written: one run of {args.c2_number} trajectories per GPU "
                                             f"(max_time {args.max_time:g}), not streamed with others: the latency-bound end of the "
                                             "run (pairs that circle without meeting, 1e7 dependent steps) is paid in full")
        # C4: laser sweep, 15 points x 1e5 trajectories, every point sharded by index over the ranks (a FIXED job)
        pts = shard_sweep([(i, w, args.c4_number, 0) for i, w in C4_POINTS], rank, world)
        flush.fill_(5)
        barrier()
        sweep = eng.run_sweep(HE, pts)
        c4_ratio = inrange_ratio  # same droplet, same table: the in-range share of evaluated pairs barely moves with lambda
        extras["c4_sweep"] = summarise(sweep[0]["info"], c4_ratio, args.mode)
        extras["c4_sweep"]["config"] = (f"BASELINE configs[3]: intensity {{0.25,0.5,1,2,4}} x pulse width {{50,100,200}} fs at N_He=1e4 "
                                        f"(lambda 0.49..31.5), {args.c4_number} trajectories per point, max_time {args.max_time:g}, one pool run "
                                        f"(heroam_gpu_run_sweep); fixed job, every point sharded by index over {world} GPU(s)")
        extras["c4_sweep"]["scaling"] = "strong"
        # C5 slice: meeting-time histogram run, histograms only, a FIXED number of trajectories split over the ranks
        first5, n5 = shard_range(args.c5_number, rank, world)
        flush.fill_(6)
        barrier()
        st = eng.run_stats(HE, n5, first_index=50 * 10**9 + first5)
        extras["c5_slice"] = summarise(st["info"], inrange_ratio, args.mode)
        extras["c5_slice"]["config"] = (f"BASELINE configs[4], a slice: {args.c5_number} of the 1e8 trajectories (histograms only, T-ref = the "
                                        f"reference-shaped table whose text file is 102 MB), fixed job split by index over {world} GPU(s)")
        extras["c5_slice"]["scaling"] = "strong"
        # C3: large droplet, tens of excitations per trajectory
        he3 = 10**7
        conf3 = make_config(number=args.c3_number, max_time=1.0e5, intensity=0.0125, helium_number=he3)
        with engine.Engine(conf3, table, device=local, mode=args.mode, segment_steps=args.segment) as eng3:
            eng3.run_stats(he3, 256, first_index=10**9)
            flush.fill_(7)
            barrier()
            st = eng3.run_stats(he3, args.c3_number, first_index=rank * args.c3_number)
        r3 = inrange_share(dict(conf=conf3, table=table), he3, 256, 2 * 10**9)
        extras["c3"] = summarise(st["info"], r3, args.mode)
        extras["c3"]["mean_excitations"] = st["particles"] / max(1, st["trajectories"])
        extras["c3"]["config"] = (f"BASELINE configs[2]: N_He=1e7 (R=478 A), intensity 0.0125 (lambda=49.2), {args.c3_number} trajectories "
                                  "per GPU, max_time 1e5")
        # T-large: the same physics on a 26.6 M-entry table (106 MB in memory): the gathers leave L1/L2
        if world == 1 and args.t_large:
            length, grid = 26_600_000, 6.0e-5
            big = synthetic_table(length=length, grid=grid)
            n_l, mt_l = args.t_large, 1.0e6
            conf_l = make_config(number=n_l, max_time=mt_l, force_grid=grid, force_grid_length=length)
            with engine.Engine(conf_l, big, device=local, mode=args.mode) as eng_l:
                eng_l.run_stats(HE, 8192, first_index=10**9)
                flush.fill_(8)
                torch.cuda.synchronize()
                st_l = eng_l.run_stats(HE, n_l, first_index=3 * 10**9)
            del big
            conf_s = make_config(number=n_l, max_time=mt_l)
            with engine.Engine(conf_s, table, device=local, mode=args.mode) as eng_s:
                eng_s.run_stats(HE, 8192, first_index=10**9)
                flush.fill_(9)
                torch.cuda.synchronize()
                st_s = eng_s.run_stats(HE, n_l, first_index=3 * 10**9)
            tl, ts = summarise(st_l["info"], inrange_ratio, args.mode), summarise(st_s["info"], inrange_ratio, args.mode)
            gathers = inrange_ratio * (st_l["info"]["pair_steps"] - st_l["info"]["pair_skipped"])
            extras["t_large"] = {"value": tl["value"], "unit": UNIT, "seconds": tl["seconds"], "roofline_frac": tl["roofline_frac"],
                                 "same_run_on_t_ref": {"value": ts["value"], "seconds": ts["seconds"]},
                                 "slowdown_vs_t_ref": tl["seconds"] / ts["seconds"],
                                 "in_range_gathers_per_sec": gathers / tl["seconds"],
                                 "config": f"{n_l} trajectories of the C2 physics to max_time {mt_l:g} on T-large (26.6 M entries on a 6e-5 A^2 "
                                           "grid = 106 MB in memory, beyond L1 and most of L2 per SM) and, for comparison, on T-ref (6.4 MB)"}
        strong = {}
        for key in ("c4_sweep", "c5_slice"):
            strong[key] = {"n_gpus": world, "seconds": extras[key]["seconds"], "trajectories": extras[key]["trajectories"],
                           "value": extras[key]["value"], "unit": UNIT}
        extras["strong_scaling"] = {"note": "fixed jobs split over the ranks by trajectory index (Philox keyed by global index: "
                                            "identical trajectories at every N); the driver's N = 1, 2, 4, 8 lines give the curve. "
                                            "What limits it is the latency-bound end of a run: the trajectories that reach max_time "
                                            "take 1e7 dependent steps however few they are (DESIGN.md section 5)",
                                    "jobs": strong}

    if rank == 0:
        traffic = traffic_source = None
        prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(prof):
            try:
                rec = json.load(open(prof))
                traffic, traffic_source = rec.get("dram_bytes_per_launch"), rec.get("source")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args),
            "trajectories_per_sec": head["trajectories_per_sec"],
            "wall_ms_per_step": 1e3 * wall_s / args.steps,
            "clocks": clocks,
            "e2e": {"value": e2e_total_ps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d // e2e_steps,
                    "d2h_bytes_per_step": d2h // e2e_steps, "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                    "api": "heroam_gpu_sample_flat + heroam_gpu_simulate_batch (the reference's Particles[] layout, "
                           "INTEGRATION.md section 2), host records in and out"},
            "gpu_launches": tot_launches,
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": peak_tflops * world, "unit": "TFLOP/s",
                         "frac": achieved / (peak_tflops * world), "traffic": traffic, "traffic_source": traffic_source,
                         "peak_three_register_ffma": peak_rrr_tflops * world,
                         "frac_of_three_register_peak": achieved / (peak_rrr_tflops * world),
                         "reference_work_rate_over_peak": tot_flops_ref / dev_s / 1e12 / (peak_tflops * world),
                         "free_flight_share_of_particle_steps": head["free_flight_share_of_particle_steps"],
                         "inrange_share_of_evaluated_pairs": inrange_ratio,
                         "model": "flops actually evaluated: 100 per full particle-step, 37 per free-flight particle-step (the orientation "
                                  "update alone), 10 per evaluated pair-step, +14 per in-range pair-step, over the DEVICE time of the "
                                  "run (first kernel to last, resolve kernels and gaps included); reference_work_rate_over_peak "
                                  "charges the SURVEY.md 8d model 100/10/14 for every step and pair the reference would evaluate -- it "
                                  "is the rate at which the reference's work is retired, not a roofline fraction (most of that work "
                                  "is provably skipped); peak = register-resident FFMA microbenchmark measured in this run "
                                  "(MEASURED_PEAKS.json has no FP32 figure), multiplier in a uniform register; "
                                  "peak_three_register_ffma = the same chains with multiplier and addend in ordinary registers, "
                                  "the form the kernels' FFMAs have",
                         "integrate_share_of_step": integ_s / dev_s, "resolve_share_of_step": resolve_s / dev_s},
            "exact_mode": exact,
            "relaxed_mode": relaxed,
            "closed_mode": closed,
            "outcomes": None if merged is None else {"success": int(merged["done"][0]), "max_time": int(merged["done"][1]),
                                                     "deadlock": int(merged["done"][2]), "trajectories": int(merged["trajectories"])},
        }
        line.update(extras)
        if world == 1 and not args.no_cpu:
            ps, s, n, cores, kind, flavour = cpu_run(table, args.max_time, max(16 * (os.cpu_count() or 1), 64))
            line["cpu_baseline"] = {"value": ps / s, "unit": UNIT, "cores": cores, "kind": kind,
                                    "trajectories_per_sec": n / s,
                                    "sample": f"{n} trajectories of the same workload (libc-rand initial states), "
                                              f"oracle/_ref '{flavour}' build, schedule(dynamic,1), {cores} threads, {s:.1f} s",
                                    "flavours_probed": pick_ref_oracle()[3]}
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
    ap.add_argument("--mode", default="fast", choices=["fast", "exact", "fast_relaxed", "fast_closed"])
    ap.add_argument("--segment", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end arm (0 = all K)")
    ap.add_argument("--warmup-e2e", type=int, default=1)
    ap.add_argument("--exact-step", type=int, default=1)
    ap.add_argument("--relaxed-step", type=int, default=-1, help="steps of the relaxed-mode run (0 = skip, -1 = the same K steps)")
    ap.add_argument("--closed-step", type=int, default=-1, help="steps of the closed-form-flight run (0 = skip, -1 = the same K steps)")
    ap.add_argument("--extras", type=int, default=1, help="0 = skip the c2_single_run / c3 / c4_sweep / c5_slice / t_large legs")
    ap.add_argument("--c2-number", type=int, default=1000000)
    ap.add_argument("--c3-number", type=int, default=10000)
    ap.add_argument("--c4-number", type=int, default=100000, help="trajectories per sweep point (whole job)")
    ap.add_argument("--c5-number", type=int, default=6000000, help="trajectories of the C5 slice (whole job)")
    ap.add_argument("--t-large", type=int, default=262144, help="trajectories of the T-large leg (0 = skip)")
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
