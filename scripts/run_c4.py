"""Config C4 (laser sweep, 15 points) as one pool run: where its time goes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sphere_roaming_b200 import engine
from sphere_roaming_b200.forcetable import synthetic_table
from sphere_roaming_b200.types import make_config
N = int(os.environ.get("N", 100000)); MT = float(os.environ.get("MT", 1e7)); HE = 10000
points = [(i, w, N, 0) for i in (0.25, 0.5, 1.0, 2.0, 4.0) for w in (50.0, 100.0, 200.0)]
only = os.environ.get("ONLY")
if only:
    points = [points[int(x)] for x in only.split(",")]
with engine.Engine(make_config(max_time=MT), synthetic_table(), mode=os.environ.get("MODE", "fast")) as eng:
    eng.run_stats(HE, 4096, first_index=10**9)
    st = eng.run_sweep(HE, points)
    i = st[0]["info"]
    print("C4 %d points x %d: %.3g particle-steps, device %.1f ms (integrate %.1f, resolve %.1f) -> %.3g ps/s, %.0f traj/s; passes %d launches %d" % (
        len(points), N, i["particle_steps"], i["device_ms"], i["integrate_ms"], i["resolve_ms"], i["particle_steps"] / i["device_ms"] * 1e3,
        len(points) * N / i["device_ms"] * 1e3, i["passes"], i["kernel_launches"]), flush=True)
    for p, s in zip(points, st):
        print("  I=%.2f tau=%3.0f mean n %.2f success %d max_time %d deadlock %d" % (p[0], p[1], s["particles"] / max(1, s["trajectories"]), s["done"][0], s["done"][1], s["done"][2]))
