"""How much would independent, overlapping pass sequences buy?  One pool run of N trajectories against two
contexts on the same GPU running N/2 each at the same time (their resolve kernels and pass tails overlap the
other's integrate kernels).  GPU only."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sphere_roaming_b200 import engine
from sphere_roaming_b200.forcetable import synthetic_table
from sphere_roaming_b200.types import make_config

N = int(os.environ.get("N", 8000000)); HE = 10000
table = synthetic_table()
conf = make_config(number=N, max_time=1e7)
with engine.Engine(conf, table, mode="fast") as eng:
    eng.run_stats(HE, 200000, first_index=10**9)
    t0 = time.perf_counter(); st = eng.run_stats(HE, N); dt = time.perf_counter() - t0
    print("one pool : %.2f s  %.4g particle-steps/s" % (dt, st["info"]["particle_steps"] / dt), flush=True)
for parts in (2, 3):
    engines = [engine.Engine(conf, table, mode="fast", pool_slots=(1 << 22) // parts) for _ in range(parts)]
    for e in engines:
        e.run_stats(HE, 100000, first_index=10**9)
    out = [None] * parts
    def work(k):
        out[k] = engines[k].run_stats(HE, N // parts, first_index=k * (N // parts))
    th = [threading.Thread(target=work, args=(k,)) for k in range(parts)]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    dt = time.perf_counter() - t0
    ps = sum(o["info"]["particle_steps"] for o in out)
    print("%d pools : %.2f s  %.4g particle-steps/s" % (parts, dt, ps / dt), flush=True)
    for e in engines: e.close()
