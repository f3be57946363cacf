"""Config C3: N_He = 1e7 (R = 478 A), lambda ~ 49, max_time = 1e5."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sphere_roaming_b200 import engine
from sphere_roaming_b200.forcetable import synthetic_table
from sphere_roaming_b200.types import make_config
N = int(os.environ.get("N", 10000)); MT = float(os.environ.get("MT", 1e5)); HE = 10**7
conf = make_config(number=N, max_time=MT, intensity=float(os.environ.get("INTENSITY", 0.0125)), helium_number=HE)
with engine.Engine(conf, synthetic_table(), mode=os.environ.get("MODE", "fast"), count_work=1) as eng:
    eng.run_stats(HE, 256)
    t = time.time(); st = eng.run_stats(HE, N); dt = time.time() - t
    i = st["info"]
    print("C3: %d trajectories, mean n %.1f, particle-steps %.3g in %.2f s -> %.3g ps/s, %.1f traj/s; pair_steps %.3g inrange %.3g passes %d" % (
        N, st["particles"] / N, i["particle_steps"], dt, i["particle_steps"] / (i["device_ms"] * 1e-3), N / dt, i["pair_steps"], i["inrange_steps"], i["passes"]))
    print("done codes", st["done"][:5], "hist_n peak", int(np.argmax(st["hist_n"])))
