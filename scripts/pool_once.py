"""One pooled run without warm-up (for ncu: the first passes run at full occupancy).
TLARGE=1: the 26.6 M-entry table (106 MB in memory) instead of T-ref (6.4 MB)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sphere_roaming_b200 import engine
from sphere_roaming_b200.forcetable import synthetic_table
from sphere_roaming_b200.types import make_config
N = int(os.environ.get("N", 1 << 20)); MT = float(os.environ.get("MT", 1e7))
GRID, LENGTH = (6.0e-5, 26_600_000) if int(os.environ.get("TLARGE", 0)) else (0.001, 1_596_000)
with engine.Engine(make_config(number=N, max_time=MT, force_grid=GRID, force_grid_length=LENGTH),
                   synthetic_table(length=LENGTH, grid=GRID), mode=os.environ.get("MODE", "fast"),
                   max_steps=int(os.environ.get("CAP", 0))) as eng:
    st = eng.run_stats(10000, N)
    i = st["info"]
    print("RESULT particle-steps/s %.4g  device_ms %.1f passes %d launches %d" % (
        i["particle_steps"] / i["device_ms"] * 1e3, i["device_ms"], i["passes"], i["kernel_launches"]))
