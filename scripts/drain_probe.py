"""Run-level experiments (GPU only): one pooled run of N trajectories of config C2 per setting, same
index block, so every setting must return identical histograms (in the modes that are schedule-independent).

    N=1000000 SETTINGS="on;base=4096;base=16384" MODE=fast python scripts/drain_probe.py

Tokens: on (defaults), base=<segment_steps>.  HEROAM_TRACE=1 prints one line per pass on stderr
(scripts/trace_view.py reads it)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sphere_roaming_b200 import engine
from sphere_roaming_b200.forcetable import synthetic_table
from sphere_roaming_b200.types import make_config

N = int(os.environ.get("N", 1000000)); MT = float(os.environ.get("MT", 1e7)); MODE = os.environ.get("MODE", "fast")
table = synthetic_table()
ref = None
for setting in os.environ.get("SETTINGS", "on").split(";"):
    kw = {}
    for tok in setting.split(","):
        if tok.startswith("base="): kw["segment_steps"] = int(tok[5:])
    with engine.Engine(make_config(number=N, max_time=MT), table, mode=MODE, **kw) as eng:
        if ref is None:
            eng.run_stats(10000, min(N, 100000), first_index=10**9)  # warm-up
        t0 = time.perf_counter()
        st = eng.run_stats(10000, N)
        wall = time.perf_counter() - t0
    i = st["info"]
    key = (st["hist_traj_time"].tobytes(), st["hist_meet_time"].tobytes(), st["done"].tobytes(), i["particle_steps"], i["pair_steps"])
    same = "first" if ref is None else ("IDENTICAL" if key == ref else "DIFFERENT")
    if ref is None: ref = key
    print("RESULT %-28s device_ms %9.1f (integrate %8.1f resolve %7.1f gaps %7.1f) wall %6.2f s  %.4g particle-steps/s  passes %5d launches %6d  %s" % (
        setting, i["device_ms"], i["integrate_ms"], i["resolve_ms"], i["device_ms"] - i["integrate_ms"] - i["resolve_ms"], wall,
        i["particle_steps"] / i["device_ms"] * 1e3, i["passes"], i["kernel_launches"], same), flush=True)
