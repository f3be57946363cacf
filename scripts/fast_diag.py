"""Fast-mode deviation from the strict oracle, in detail (exploratory)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.binding import PortOracle
from sphere_roaming_b200 import engine
from sphere_roaming_b200.forcetable import synthetic_table
from sphere_roaming_b200.types import make_config, offsets_of
table = synthetic_table(); port = PortOracle(); HE = 10000
n, max_time = int(os.environ.get("N", 4096)), float(os.environ.get("MT", 2e5))
conf = make_config(number=n, max_time=max_time)
counts, init = port.sample(conf, HE, table, n, source="philox")
want, wres = port.simulate(conf, HE, table, counts, init)
with engine.Engine(conf, table, mode=os.environ.get("MODE", "fast")) as eng:
    got, res = eng.simulate(HE, counts, init)
off = offsets_of(counts)
dt = np.abs(res["time"].astype(np.float64) - wres["time"])
print("exit time: differ in", int((dt > 0).sum()), "of", n, "; quantiles of nonzero |dt|:", np.quantile(dt[dt > 0], [0.5, 0.9, 0.99, 1.0]) if (dt > 0).any() else None)
print("within max(50, 1e-4 t):", (dt <= np.maximum(50.0, 1e-4 * wres["time"])).mean())
same_flags = np.array([np.array_equal(got["success"][off[i]:off[i+1]], want["success"][off[i]:off[i+1]]) for i in range(n)])
print("same success pattern:", same_flags.mean(), " status equal:", (res["status"] == wres["status"]).mean())
bad = np.nonzero(~same_flags)[0][:10]
for i in bad:
    print("  traj", i, "n", counts[i], "gpu", got["success"][off[i]:off[i+1]], res[i], "ref", want["success"][off[i]:off[i+1]], wres[i])
same = np.repeat(res["steps"] == wres["steps"], counts) & (got["time"] == want["time"])
dr = np.abs(got["position"].astype(np.float64) - want["position"]).max(axis=1)[same]
print("positions (same exit step): median %.3g  p99 %.3g  max %.3g" % (np.median(dr), np.quantile(dr, 0.99), dr.max()))
dpt = np.abs(got["time"].astype(np.float64) - want["time"])
print("per-particle time: differ", (dpt > 0).mean(), " p99 %.1f max %.1f" % (np.quantile(dpt, 0.99), dpt.max()))
err_all = np.abs(got["position"].astype(np.float64) - want["position"]).max(axis=1)
nn = np.repeat(counts, counts)
flying_like = np.repeat(res["steps"] == wres["steps"], counts)
for k in range(1, 9):
    sel = (nn == k) & flying_like
    if sel.any(): print("n=%d: particles %d median err %.3g  p90 %.3g" % (k, sel.sum(), np.median(err_all[sel]), np.quantile(err_all[sel], 0.9)))
dw = np.abs(got["ang_velocity"].astype(np.float64) - want["ang_velocity"]).max(axis=1) / np.abs(want["ang_velocity"]).max(axis=1)
print("relative |dw|: median %.3g p99 %.3g" % (np.median(dw[flying_like]), np.quantile(dw[flying_like], 0.99)))
