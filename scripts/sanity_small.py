"""Small end-to-end run touching every kernel family (for compute-sanitizer)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sphere_roaming_b200 import engine
from sphere_roaming_b200.forcetable import synthetic_table
from sphere_roaming_b200.types import make_config
table = synthetic_table()
for he, inten, n, mt in ((10000, 1.0, 600, 3.0e4), (40000, 1.5, 48, 1500.0), (40000, 5.0, 24, 600.0), (10**7, 0.0125, 24, 5000.0)):
    conf = make_config(number=n, max_time=mt, intensity=inten, helium_number=he)
    for mode in ("exact", "fast"):
        with engine.Engine(conf, table, mode=mode, segment_steps=512, pool_slots=max(8, n // 3)) as eng:
            counts, init = eng.sample(he, n)
            final, res = eng.simulate(he, counts, init)
            st = eng.run_stats(he, n)
            print(he, inten, mode, "n max", counts.max(), "done", np.bincount(res["status"], minlength=3)[:3], st["trajectories"], st["info"]["passes"])
print("OK")
