"""Deviation of the fast modes from the strict oracle, from identical initial states (GPU only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.binding import PortOracle
from sphere_roaming_b200 import engine
from sphere_roaming_b200.forcetable import synthetic_table
from sphere_roaming_b200.types import make_config, offsets_of
table = synthetic_table(); port = PortOracle(); HE = 10000
n = int(os.environ.get("N", 4096))
for max_time in (1.0e4, 2.0e5):
    conf = make_config(number=n, max_time=max_time)
    counts, init = port.sample(conf, HE, table, n, source="philox")
    want, wres = port.simulate(conf, HE, table, counts, init)
    off = offsets_of(counts)
    for mode in ("fast", "fast_relaxed"):
        with engine.Engine(conf, table, mode=mode) as eng:
            got, res = eng.simulate(HE, counts, init)
        dt = np.abs(res["time"].astype(np.float64) - wres["time"])
        ok_time = dt <= np.maximum(50.0, 1e-4 * wres["time"])
        same_flags = np.array([np.array_equal(got["success"][off[i]:off[i + 1]], want["success"][off[i]:off[i + 1]]) for i in range(n)])
        same_step = np.repeat(res["steps"] == wres["steps"], counts)
        dr = np.abs(got["position"].astype(np.float64) - want["position"]).max(axis=1)[same_step & (got["time"] == want["time"])]
        q = np.quantile(dr, [0.5, 0.9, 0.99, 0.999])
        print("max_time %.0e %-12s exit-time ok %.4f (differ %d, max |dt| %.0f)  flags same %.4f  |dr| median %.2e  90%% %.2e  99%% %.2e  99.9%% %.2e  max %.2e  (n=%d)" % (
            max_time, mode, ok_time.mean(), int((dt > 0).sum()), dt.max(), same_flags.mean(), q[0], q[1], q[2], q[3], dr.max(), len(dr)))
