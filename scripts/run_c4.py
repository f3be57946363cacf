"""Config C4: intensity x pulse-width sweep at N_He = 1e4, `N` trajectories per
point -- as ONE run of the pool (heroam_gpu_run_sweep) and as a loop of separate runs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sphere_roaming_b200 import engine
from sphere_roaming_b200.forcetable import synthetic_table
from sphere_roaming_b200.types import make_config
N = int(os.environ.get("N", 100000)); MT = float(os.environ.get("MT", 1e7)); HE = 10000
points = [(i, w, N, 0) for i in (0.25, 0.5, 1.0, 2.0, 4.0) for w in (50.0, 100.0, 200.0)]
table = synthetic_table()
with engine.Engine(make_config(max_time=MT), table, mode="fast") as eng:
    eng.run_stats(HE, 4096)
    t = time.time(); out = eng.run_sweep(HE, points); dt = time.time() - t
    i = out[0]["info"]
    print("C4 one launch : %d points x %d trajectories, %.3g particle-steps in %.2f s (device %.2f s) -> %.3g ps/s, %.0f traj/s, passes %d" % (
        len(points), N, i["particle_steps"], dt, i["device_ms"] * 1e-3, i["particle_steps"] / (i["device_ms"] * 1e-3), len(points) * N / dt, i["passes"]))
    if os.environ.get("LOOP", "1") == "1":
        t = time.time(); ps = 0
        for p in points:
            ps += eng.run_sweep(HE, [p])[0]["info"]["particle_steps"]
        dt2 = time.time() - t
        print("C4 point loop : %.3g particle-steps in %.2f s -> %.3g ps/s, %.0f traj/s  (one launch is %.2fx faster)" % (
            ps, dt2, ps / dt2, len(points) * N / dt2, dt2 / dt))
    for p, st in zip(points, out):
        print("  I=%.2f tau=%3.0f  mean n %.2f  success %d  max_time %d  mean exit %.3g fs" % (
            p[0], p[1], st["particles"] / max(1, st["trajectories"]), st["done"][0], st["done"][1], st["sum_traj_time"] / max(1, st["trajectories"])))
