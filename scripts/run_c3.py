"""Config C3: N_He = 1e7 (R = 478 A), lambda ~ 49, max_time = 1e5; where its time goes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sphere_roaming_b200 import engine
from sphere_roaming_b200.forcetable import synthetic_table
from sphere_roaming_b200.types import make_config
MT = float(os.environ.get("MT", 1e5)); HE = 10**7
table = synthetic_table()
for N in [int(x) for x in os.environ.get("NS", "10000,20000,40000").split(",")]:
    conf = make_config(number=N, max_time=MT, intensity=float(os.environ.get("INTENSITY", 0.0125)), helium_number=HE)
    with engine.Engine(conf, table, mode=os.environ.get("MODE", "fast"), segment_steps=int(os.environ.get("SEG", 0))) as eng:
        eng.run_stats(HE, 256, first_index=10**9)
        st = eng.run_stats(HE, N)
        i = st["info"]
        print("C3 N=%d: mean n %.1f, %.3g particle-steps, device %.1f ms (integrate %.1f, resolve %.1f, gaps %.1f) -> %.3g ps/s; "
              "passes %d launches %d free %.0f%% pairs skipped %.1f%%" % (
                  N, st["particles"] / N, i["particle_steps"], i["device_ms"], i["integrate_ms"], i["resolve_ms"],
                  i["device_ms"] - i["integrate_ms"] - i["resolve_ms"], i["particle_steps"] / (i["device_ms"] * 1e-3), i["passes"],
                  i["kernel_launches"], 100.0 * i["free_steps"] / i["particle_steps"], 100.0 * i["pair_skipped"] / max(1, i["pair_steps"])), flush=True)
