"""Generate tests/golden/horizon_*.npz from the REFERENCE ITSELF (oracle/_ref) at the FULL
horizon of the BASELINE.json configurations -- the sizes bench.py runs at.

Run in the development container (where /root/reference exists and `make -C oracle ref`
has built oracle/_ref/*.so):

    python tests/golden/make_golden_horizon.py [c2] [c3] [c4lo] [c4hi]

  horizon_c2.npz    C2 (BASELINE configs[1]): repo config.conf physics, 4096 trajectories from
                    reference-exported initial states (initialize_particles, libc rand, the
                    config's seed) run by simulate_particles to max_time = 1e7 fs; strict build
                    (-O2 -ffp-contract=off: the parity anchor) and, for the reference's own
                    build-to-build spread, the release build (meson.build:19-21 flags).
  horizon_c3.npz    C3 (configs[2]): N_He = 1e7, intensity 0.0125 (lambda = 49.2), 256
                    trajectories, max_time = 1e5 fs.
  horizon_c4lo.npz  C4 corner (configs[3]): intensity 0.25, pulse 50 fs (lambda = 0.49), N_He = 1e4,
  horizon_c4hi.npz  C4 corner: intensity 4, pulse 200 fs (lambda = 31.5), 64 trajectories (256 for c4lo),
                    max_time = 1e7 fs.

The reference has no golden vectors of its own; these files pin its
behaviour by executing its unmodified object code.  The GPU box has no /root/reference: tests
there read only the committed .npz files.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.binding import RefOracle  # noqa: E402
from sphere_roaming_b200.forcetable import synthetic_table  # noqa: E402
from sphere_roaming_b200.types import make_config  # noqa: E402


def run(ref, conf, he, table, counts, init, what):
    t0 = time.perf_counter()
    # guarded: the reference itself never returns from a trajectory whose particles are all frozen with
    # its `==` success criterion unmet (71 of 1e6 trajectories of C2) -- see ref_harness.c
    final, dead = ref.simulate_guarded(conf, he, table, counts, init, nthreads=0, dynamic=True)
    dt = time.perf_counter() - t0
    ps = float(final["time"].astype(np.float64).sum()) / conf.dt
    print(f"{what}: {len(counts)} trajectories, {ps:.3e} particle-steps in {dt:.1f} s ({ps / dt:.3e}/s), "
          f"{int(dead.sum())} deadlocked in the reference", flush=True)
    return final, dead


def c2(table):
    he = 10000
    conf = make_config(number=4096, max_time=1.0e7)
    strict, release = RefOracle("strict"), RefOracle("release")
    counts, init = strict.initialize(conf, he, table)
    final, dead = run(strict, conf, he, table, counts, init, "C2 strict")
    rel, _ = run(release, conf, he, table, counts, init, "C2 release")
    np.savez_compressed(os.path.join(HERE, "horizon_c2.npz"), he_number=np.int64(he), max_time=np.float32(conf.max_time),
                        counts=counts, init=init, final=final, deadlocked=dead, release_time=rel["time"],
                        release_success=rel["success"], release_position=rel["position"])


def generic(table, name, he, number, max_time, **over):
    conf = make_config(number=number, max_time=max_time, **over)
    strict = RefOracle("strict")
    counts, init = strict.initialize(conf, he, table)
    final, dead = run(strict, conf, he, table, counts, init, name)
    extra = {k: np.float32(v) for k, v in over.items()}
    np.savez_compressed(os.path.join(HERE, f"horizon_{name}.npz"), he_number=np.int64(he), max_time=np.float32(max_time),
                        counts=counts, init=init, final=final, deadlocked=dead, **extra)


def main():
    table = synthetic_table()
    which = sys.argv[1:] or ["c2", "c3", "c4lo", "c4hi"]
    if "c3" in which:
        generic(table, "c3", 10_000_000, 256, 1.0e5, intensity=0.0125)
    if "c4lo" in which:
        generic(table, "c4lo", 10000, 256, 1.0e7, intensity=0.25, pulse_width=50.0)
    if "c4hi" in which:
        generic(table, "c4hi", 10000, 64, 1.0e7, intensity=4.0, pulse_width=200.0)
    if "c2" in which:
        c2(table)
    for f in sorted(os.listdir(HERE)):
        if f.startswith("horizon_"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
