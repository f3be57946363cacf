import os, sys, subprocess, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    from sphere_roaming_b200 import engine
    from sphere_roaming_b200.forcetable import synthetic_table
    from sphere_roaming_b200.types import make_config
    n = 4096
    conf = make_config(number=n, max_time=float(os.environ.get("MT", 3e4)))
    with engine.Engine(conf, synthetic_table(), mode="fast") as eng:
        counts, init = eng.sample(10000, n)
        final, res = eng.simulate(10000, counts, init)
    np.savez(sys.argv[1], counts=counts, final=final, res=res)
else:
    for la in ("0", "1"):
        subprocess.check_call([sys.executable, __file__, f"/tmp/la{la}.npz"], env=dict(os.environ, HEROAM_LOOKAHEAD=la))
    a, b = np.load("/tmp/la0.npz"), np.load("/tmp/la1.npz")
    fa, fb = a["final"], b["final"]
    for name in fa.dtype.names:
        x, y = fa[name].astype(np.float64), fb[name].astype(np.float64)
        nd = (x != y).reshape(len(x), -1).any(axis=1).sum()
        print(name, "differs in", nd, "records, max abs", np.abs(x - y).max())
    print("steps differ", (a["res"]["steps"] != b["res"]["steps"]).sum())
    off = np.concatenate([[0], np.cumsum(a["counts"])])
    bad = np.nonzero((fa["position"] != fb["position"]).any(axis=1))[0]
    if len(bad):
        t = np.searchsorted(off, bad[0], side="right") - 1
        print("first bad record", bad[0], "trajectory", t, "n", a["counts"][t], a["res"][t], b["res"][t])
        print(fa[off[t]:off[t+1]]["position"], fb[off[t]:off[t+1]]["position"], fa[off[t]:off[t+1]]["success"])
