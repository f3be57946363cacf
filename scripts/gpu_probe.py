"""Exploratory GPU check (not a test): parity vs the oracle + rough throughput."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.binding import PortOracle
from sphere_roaming_b200 import engine
from sphere_roaming_b200.forcetable import synthetic_table
from sphere_roaming_b200.types import PARTICLE_DTYPE, make_config

table = synthetic_table()
he = 10000
oracle = PortOracle()
N = int(os.environ.get("N", 512))
MT = float(os.environ.get("MT", 1e5))
conf = make_config(number=N, max_time=MT)
counts, init = oracle.sample(conf, he, table, N, source="philox")
t = time.time(); want, wres, wctr = oracle.simulate(conf, he, table, counts, init, want_counters=True); tc = time.time() - t
print("oracle: %.2fs, particle_steps=%d -> %.3g ps/s (%d threads)" % (tc, wctr["particle_steps"], wctr["particle_steps"]/tc, oracle.max_threads()))
print("n hist", np.bincount(counts)[:16])
with engine.Engine(conf, table, mode="exact") as eng:
    print("fp32 peak", eng.fp32_peak())
    gc, gi = eng.sample(he, N)
    print("sample: counts eq", np.array_equal(gc, counts))
    for name in PARTICLE_DTYPE.names:
        if not np.array_equal(gi[name], init[name]):
            a = gi[name].astype(np.float64); b = init[name].astype(np.float64)
            print("  sample diff", name, "n=", int((a != b).sum()), "maxrel", np.max(np.abs(a-b)/(np.abs(b)+1e-30)))
    t = time.time(); got, res = eng.simulate(he, counts, init); tg = time.time() - t
    info = eng.run_info()
    print("gpu exact: %.3fs wall, info=%s" % (tg, json.dumps(info)))
    print("  ps/s (device_ms): %.3g" % (info["particle_steps"]/(info["device_ms"]*1e-3)))
    ok = True
    for name in PARTICLE_DTYPE.names:
        if not np.array_equal(got[name], want[name]):
            ok = False
            a = got[name].astype(np.float64); b = want[name].astype(np.float64)
            print("  EXACT DIFF", name, "n=", int((a != b).sum()), "max", np.abs(a-b).max())
    print("exact parity:", ok, "steps eq", np.array_equal(res["steps"], wres["steps"]), "status eq", np.array_equal(res["status"], wres["status"]),
          "counters", info["particle_steps"] == wctr["particle_steps"], info["pair_steps"] == wctr["pair_steps"], info["inrange_steps"] == wctr["inrange_steps"])
    print("status hist", np.bincount(res["status"], minlength=5))
    eng.set_mode("fast")
    t = time.time(); fast, fres = eng.simulate(he, counts, init); tf = time.time() - t
    info = eng.run_info()
    print("gpu fast: %.3fs wall, ps/s %.3g, info=%s" % (tf, info["particle_steps"]/(info["device_ms"]*1e-3), json.dumps(info)))
    dpos = np.abs(fast["position"].astype(np.float64) - want["position"]).max(axis=1)
    print("  fast: max|dr| %.3g, mean %.3g; steps differ in %d of %d, max |dsteps| %d; success eq %s" % (
        dpos.max(), dpos.mean(), int((fres["steps"] != wres["steps"]).sum()), N,
        int(np.abs(fres["steps"].astype(np.int64) - wres["steps"]).max()), np.array_equal(fast["success"], want["success"])))
