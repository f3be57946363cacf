"""Generate tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref, strict build).

Run in the development container, where /root/reference exists and
`make -C oracle ref` has built oracle/_ref/libheroam_ref_strict.so:

    python tests/golden/make_golden.py

The reference has no tests, fixtures or golden vectors of its own (SURVEY.md
section 4); these files pin its behaviour by executing its unmodified object code.
The GPU box has no /root/reference: tests there read only the committed .npz files.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.binding import RefOracle  # noqa: E402
from sphere_roaming_b200.forcetable import synthetic_table  # noqa: E402
from sphere_roaming_b200.types import PARTICLE_DTYPE, make_config  # noqa: E402

_fp = C.POINTER(C.c_float)


def fptr(a):
    return a.ctypes.data_as(_fp)


def vector_kats(ref: RefOracle, rng: np.random.Generator, n: int = 256) -> dict:
    """Known answers of every vector.c function on the hot path."""
    L = ref.lib
    a = (rng.standard_normal((n, 3)) * rng.choice([1e-5, 1.0, 47.8], (n, 1))).astype(np.float32)
    b = (rng.standard_normal((n, 3)) * rng.choice([1e-6, 1.0, 47.8], (n, 1))).astype(np.float32)
    s = rng.standard_normal(n).astype(np.float32)
    q1 = rng.standard_normal((n, 4)).astype(np.float32)
    q2 = rng.standard_normal((n, 4)).astype(np.float32)
    q1[: n // 2] /= np.linalg.norm(q1[: n // 2], axis=1, keepdims=True).astype(np.float32)
    a[0] = (0.0, 0.0, 47.8)  # the pole branch of create_perpend_vector
    out = dict(a=a, b=b, s=s, q1=q1, q2=q2)
    cross = np.zeros_like(a); add = np.zeros_like(a); scal = np.zeros_like(a); perp = np.zeros_like(a)
    nrm = np.zeros(n, np.float32); nsq = np.zeros(n, np.float32)
    qprod = np.zeros_like(q1); qinv = np.zeros_like(q1); qpos = np.zeros_like(a); qnorm = np.zeros(n, np.float32)
    for i in range(n):
        L.cross_product(fptr(a[i]), fptr(b[i]), fptr(cross[i]))
        L.add_vectors(fptr(a[i]), fptr(b[i]), fptr(add[i]))
        L.scalar_multiply(C.c_float(float(s[i])), fptr(a[i]), fptr(scal[i]))
        L.create_perpend_vector(fptr(a[i]), fptr(perp[i]))
        nrm[i] = L.norm(fptr(a[i]))
        nsq[i] = L.norm_sq(fptr(a[i]))
        L.quat_product(fptr(q1[i]), fptr(q2[i]), fptr(qprod[i]))
        L.quat_inverse(fptr(q1[i]), fptr(qinv[i]))
        L.quat_to_pos(fptr(q1[i]), fptr(qpos[i]))
        qnorm[i] = L.quat_norm(fptr(q1[i]))
    out.update(cross=cross, add=add, scal=scal, perp=perp, norm=nrm, norm_sq=nsq, qprod=qprod, qinv=qinv,
               qpos=qpos, qnorm=qnorm)
    # file-static helpers of simulation.c
    mass, radius, hdt = np.float32(4.0), np.float32(47.82845), np.float32(0.5)
    acc = np.zeros_like(a); wnew = np.zeros_like(a)
    uq = q1.copy(); upos = np.zeros_like(a)
    w = (rng.standard_normal((n, 3)) * 1e-5).astype(np.float32)
    for i in range(n):
        L.ref_find_accelration(mass, radius, fptr(a[i]), fptr(b[i]), fptr(acc[i]))
        L.ref_update_angvel(hdt, fptr(w[i]), fptr(acc[i]), fptr(wnew[i]))
        L.ref_update_orientation(radius, np.float32(1.0), fptr(uq[i]), fptr(w[i]), fptr(upos[i]))
    out.update(mass=mass, radius=radius, hdt=hdt, w=w, acc=acc, wnew=wnew, uq=uq, upos=upos)
    return out


def force_cases(ref: RefOracle, table: np.ndarray) -> dict:
    """calculate_force / find_success on hand-built states."""
    conf = make_config()
    R = 47.82845

    def on_sphere(lon_deg, lat_deg):
        lo, la = np.radians(lon_deg), np.radians(lat_deg)
        return (R * np.cos(la) * np.cos(lo), R * np.cos(la) * np.sin(lo), R * np.sin(la))

    def mk(points, frozen=()):
        p = np.zeros(len(points), dtype=PARTICLE_DTYPE)
        for i, pt in enumerate(points):
            p[i]["index"] = i
            p[i]["position"] = pt
            p[i]["force"] = (9.0, 9.0, 9.0)  # must be overwritten
            p[i]["force_mag"] = 7.0
            p[i]["success"] = 1 if i in frozen else 0
        return p

    deg = np.degrees(5.0 / R)  # about 5 A of arc
    cases = {
        "single": mk([on_sphere(0, 0)]),
        "far_pair": mk([on_sphere(0, 0), on_sphere(120, 10)]),
        "in_range_pair": mk([on_sphere(0, 0), on_sphere(20, 5)]),
        "meeting_pair": mk([on_sphere(0, 0), on_sphere(deg, 0)]),
        # 0 meets 1 and 2 in one sweep: three flags (Q3); 3 is near 2 but 2 is taken when 3's turn comes
        "triple_flag": mk([on_sphere(0, 0), on_sphere(deg, 0), on_sphere(-deg, 0), on_sphere(-2 * deg, 0), on_sphere(90, 0)]),
        # a frozen particle neither meets nor exerts force (Q8)
        "frozen_neighbour": mk([on_sphere(0, 0), on_sphere(deg, 0), on_sphere(25, 0)], frozen=(1,)),
        "six_mixed": mk([on_sphere(0, 0), on_sphere(15, 3), on_sphere(30, -8), on_sphere(50, 20), on_sphere(52, 24), on_sphere(200, -40)]),
        "last_index_force_mag": mk([on_sphere(0, 0), on_sphere(18, 0), on_sphere(36, 0)]),
    }
    out = {}
    for name, p in cases.items():
        out[f"{name}__in"] = p
        res = ref.calculate_force(conf, table, p)
        out[f"{name}__out"] = res
        out[f"{name}__success"] = np.int32(ref.find_success(res))
    return out


def trajectory_goldens(ref: RefOracle, table: np.ndarray) -> dict:
    """96 trajectories of the repository config: sampler output (libc rand, seed of
    config.conf), states after k steps (simulate_particles with max_time = k*dt), final."""
    he = 10000
    conf = make_config(number=96, max_time=3.0e5)
    counts, init = ref.initialize(conf, he, table)
    out = dict(he_number=np.int64(he), counts=counts, init=init, max_time=np.float32(conf.max_time))
    for k in (1, 10, 100, 1000, 10000):
        ck = make_config(number=96, max_time=float(k))
        out[f"step_{k}"] = ref.simulate(ck, he, table, counts, init)
    out["final"] = ref.simulate(conf, he, table, counts, init)
    # a dense droplet state: many early meetings, odd/even termination, 3-flag sweeps
    conf_d = make_config(number=64, max_time=2.0e4, intensity=6.0)
    he_d = 1000
    c_d, i_d = ref.initialize(conf_d, he_d, table)
    out.update(dense_he_number=np.int64(he_d), dense_counts=c_d, dense_init=i_d,
               dense_intensity=np.float32(6.0), dense_max_time=np.float32(conf_d.max_time),
               dense_final=ref.simulate(conf_d, he_d, table, c_d, i_d))
    return out


def sampler_streams(ref: RefOracle) -> dict:
    """libc-rand driven draws of the two scalar samplers (simulation.c:238-275)."""
    conf = make_config()
    ref.lib.ref_srand(12345)
    counts = np.array([ref.lib.ref_generate_no_of_excitation(3.9337242843335494) for _ in range(2000)], dtype=np.int32)
    ref.lib.ref_srand(54321)
    speeds = np.array([ref.lib.ref_generate_velocity(C.byref(conf)) for _ in range(2000)], dtype=np.float32)
    ref.lib.ref_srand(777)
    quats = np.zeros((500, 4), np.float32)
    for i in range(500):
        ref.lib.quat_random(fptr(quats[i]))
    return dict(poisson_seed=np.int64(12345), poisson_lambda=np.float64(3.9337242843335494), poisson=counts,
                speed_seed=np.int64(54321), speeds=speeds, quat_seed=np.int64(777), quats=quats)


def particlefile_golden(ref: RefOracle, table: np.ndarray) -> None:
    """particlefile text written by the reference's own save_particles (simulation.c:415-436)
    under the header of main.c:65-69, for the first 6 golden trajectories."""
    from sphere_roaming_b200.types import ParticlesC, offsets_of

    g = np.load(os.path.join(HERE, "trajectories.npz"))
    counts, final = g["counts"][:6], g["final"]
    off = offsets_of(g["counts"])
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    libc.fputs.argtypes = [C.c_char_p, C.c_void_p]
    ref.lib.save_particles.argtypes = [C.c_void_p, C.POINTER(ParticlesC)]
    ref.lib.save_particles.restype = None
    path = os.path.join(HERE, "particlefile_after_golden.tsv")
    f = libc.fopen(path.encode(), b"w")
    libc.fputs(b"Index\tNumber\tSuccess\tTime\tX\tY\tZ\tVX\tVY\tVZ\tFX\tFY\tFZ\n", f)
    keep = []
    for i in range(6):
        rec = np.ascontiguousarray(final[off[i]:off[i + 1]])
        keep.append(rec)
        ps = ParticlesC(int(counts[i]), i, rec.ctypes.data)
        ref.lib.save_particles(f, C.byref(ps))
    libc.fclose(f)


def main() -> None:
    ref = RefOracle("strict")
    assert ref.sizes() == (84, 16, 152)
    table = synthetic_table()
    rng = np.random.default_rng(20260119)
    np.savez_compressed(os.path.join(HERE, "vector_kats.npz"), **vector_kats(ref, rng))
    np.savez_compressed(os.path.join(HERE, "force_cases.npz"), **force_cases(ref, table))
    np.savez_compressed(os.path.join(HERE, "trajectories.npz"), **trajectory_goldens(ref, table))
    np.savez_compressed(os.path.join(HERE, "sampler_streams.npz"), **sampler_streams(ref))
    np.savez_compressed(os.path.join(HERE, "layout.npz"), sizes=np.array(ref.sizes()),
                        particle_offsets=np.array(ref.particle_offsets()), config_offsets=np.array(ref.config_offsets()),
                        table_head=table[:8], table_tail=table[-8:], table_sum=np.float64(table.astype(np.float64).sum()))
    particlefile_golden(ref, table)
    for f in sorted(os.listdir(HERE)):
        if f.endswith((".npz", ".tsv")):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
