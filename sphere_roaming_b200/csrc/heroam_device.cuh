// heroam_device.cuh -- device-side state layout and arithmetic of the trajectory
// integrator (sm_100a).  Two arithmetic policies:
//
//   Exact : every float operation of the reference's step (simulation.c:478-509,
//           :158-236, vector.c) in the reference's order through the never-contracted
//           __f*_rn intrinsics; sqrt / reciprocal / divide correctly rounded (a double
//           sqrt or divide of float operands narrowed to float equals the correctly
//           rounded float result, so (float)sqrt((double)x) == __fsqrt_rn(x) and
//           (float)(1.0/(double)x) == __frcp_rn(x)); the one genuinely double
//           expression, Force_list[i]/sqrt(d) (simulation.c:220), in double.
//           Terms the reference multiplies by a literal 0 or 1 are dropped: they change
//           at most the sign of an exact zero, never a value.
//   Fast  : same state and integrator; FMA contraction, closed-form pole image,
//           MUFU rsqrt.  The GPU counterpart of the reference's -ffast-math build.
//   Turbo : Fast with the loop-invariant factors folded into the state kept by the hot
//           kernels (they hold w' = (dt/2) w and F' = (dt/2)^2/(m R^2) F and gather from a
//           table pre-multiplied by that factor, so the half kick is a bare cross product and
//           the orientation update needs no scaling), raw MUFU.RSQ, one-FMA table position,
//           and optionally no work counting.  Used by the integrate kernels in fast mode;
//           the once-per-pass resolve kernel keeps the plain Fast arithmetic.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/heroam_gpu.h"

namespace heroam {

// Trajectory status while it is being worked on (device side only).
enum : uint32_t {
  ST_LIVE = 0,     // clean state at a step boundary
  ST_MIDSTEP = 1,  // a pair came within icd_dist during the step: drift done, greedy
                   // meeting sweep + force + second kick still to do (resolve kernel)
  ST_FLYING = 2,   // scheduled into the free-flight kernel for [steps, step_stop): the next
                   // resolve pass commits steps/time (the kernel never writes TrajMeta)
  ST_DONE = 16     // ST_DONE + HEROAM_DONE_* once finished
};

struct TrajMeta {      // 60 bytes per trajectory
  uint32_t first;      // first record of this trajectory in the flat particle array
  uint32_t no;         // particles in the trajectory
  uint32_t steps;      // completed iterations of the step loop (of the core, see below)
  float time;          // the loop's `time` variable (simulation.c:467, :512-517)
  uint32_t n_flagged;  // particles with success == 1
  uint32_t n_active;   // particles still moving that the integrate kernels work on (the "core")
  uint32_t status;     // ST_*
  uint32_t step_stop;  // the integrate kernel pauses when steps reaches this
  // Loners: moving particles that provably cannot come within reach of any other particle
  // before loner_until.  They fly on their own (k_free_flight) and may be ahead of the core.
  uint32_t n_loners;
  uint32_t loner_from;   // step at which their current flight started
  uint32_t loner_until;  // step they have reached; the core catches up before the next analysis
  float loner_time0;     // the clock at loner_from
  uint32_t cross_from;   // core step up to which core-loner / loner-loner pair-steps are accounted
  uint32_t core_slots;   // act[first .. first+core_slots) = core, then n_loners loner entries
  uint32_t point;        // sweep point the trajectory belongs to (streaming pool; 0 otherwise)
};

// Loop-invariant scalars, rounded on the host exactly where the reference rounds them.
struct Consts {
  float dt;          // simulation.c:443
  float half_dt;     // (float)(dt / 2.0)                       :486, :502, :406
  float radius;      // (float)(2.22 * pow((float)N, 1.0/3.0))  :448
  float acc_scale;   // (float)(1.0 / mass / radius / radius)   :351
  float inv_grid;    // (float)(1.0 / force_grid)               :162
  float grid_start;  // :163
  float icd2;        // icd_dist * icd_dist (float product)     :164
  float max_time;    // :469
  int grid_len;      // :219 (clamped to INT_MAX on the host)
  uint32_t max_steps;   // steps until the float clock reaches max_time (host-iterated)
  uint32_t clock_exact; // 1: after k steps the reference's clock (time += dt, k times) equals (float)k * dt for every
                        // k <= max_steps (checked on the host): a flight may then set it at landing instead of
                        // adding dt in every step
  // fast-mode folded constants
  float far2;        // squared distance beyond which a pair can neither feel force nor meet (with margin)
  float kick;        // acc_scale * half_dt
  float radius2;     // 2 * radius
  // Turbo folded constants
  float kappa;       // acc_scale * half_dt^2 : force scale of the hot kernels' state and table copy
  float inv_kappa;
  float inv_half_dt;
  float pos_bias;    // -grid_start * inv_grid
  // rigorous reach bounds (partner lists)
  float f_max;       // max |table| over squared distances >= icd^2 (unmet pairs are never closer)
  float inv_mr;      // 1 / (mass * radius): |angular acceleration| <= |F| * inv_mr
};

struct Counters {     // roofline work counters
  unsigned long long particle_steps;
  unsigned long long pair_steps;
  unsigned long long inrange_steps;
  unsigned long long unsupported;   // trajectories refused for having too many moving particles
  unsigned long long pair_skipped;  // pair-steps of free-flight segments: counted in pair_steps, never evaluated
  unsigned long long free_steps;    // particle-steps taken in free flight (orientation update only)
};

constexpr uint32_t NBR_ALL = 0xffffffffu;

struct DevView {       // everything a kernel needs, passed by value
  Consts c;
  const float *__restrict__ table;    // grid_len + 1 entries, the last one 0 (out-of-table bin)
  const float *__restrict__ table_k;  // the same multiplied by kappa (Turbo)
  heroam_particle *particles;
  TrajMeta *meta;
  uint32_t *act;       // act[first + s] = index (within the trajectory) of its s-th active particle
  float4 *ckpt_q;      // [first + j] orientation of particle j when it last took off as a loner
  // partner lists of the warp-per-trajectory kernels (rebuilt every pass by the resolve kernel)
  uint8_t *nbr_pool;   // concatenated ascending partner slots
  uint32_t *nbr_start; // [first + s] offset of slot s's partners in nbr_pool
  uint32_t *nbr_cnt;   // [first + s] number of partners; NBR_ALL = iterate over every other particle
  unsigned long long *nbr_used;
  unsigned long long nbr_cap;
  Counters *counters;
  uint32_t n_traj;
};

// ---------------------------------------------------------------------------------
// arithmetic policies
// ---------------------------------------------------------------------------------
struct Exact {
  static constexpr bool closed_flight = false; // free flight one step at a time, as the reference takes them
  static constexpr unsigned renorm_every = 1;  // steps between renormalisations of the orientation
  static constexpr bool exact = true;
  static constexpr bool turbo = false;
  static constexpr bool count = true;
  __device__ __forceinline__ static float mul(float a, float b) { return __fmul_rn(a, b); }
  __device__ __forceinline__ static float add(float a, float b) { return __fadd_rn(a, b); }
  __device__ __forceinline__ static float sub(float a, float b) { return __fsub_rn(a, b); }
  // a*b - c*d and a*b + c*d with three roundings, as written in the reference
  __device__ __forceinline__ static float mm_sub(float a, float b, float c, float d) {
    return __fsub_rn(__fmul_rn(a, b), __fmul_rn(c, d));
  }
  __device__ __forceinline__ static float mm_add(float a, float b, float c, float d) {
    return __fadd_rn(__fmul_rn(a, b), __fmul_rn(c, d));
  }
  // acc + a*b / acc - a*b (two roundings)
  __device__ __forceinline__ static float acc_add(float acc, float a, float b) {
    return __fadd_rn(acc, __fmul_rn(a, b));
  }
  __device__ __forceinline__ static float acc_sub(float acc, float a, float b) {
    return __fsub_rn(acc, __fmul_rn(a, b));
  }
  __device__ __forceinline__ static float dot3(float x, float y, float z) {  // x*x + y*y + z*z
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
  }
  __device__ __forceinline__ static float dot4(float a, float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(x, x)), __fmul_rn(y, y)),
                     __fmul_rn(z, z));
  }
};

struct Fast {
  static constexpr bool closed_flight = false;
  static constexpr unsigned renorm_every = 1;
  static constexpr bool exact = false;
  static constexpr bool turbo = false;
  static constexpr bool count = true;
  __device__ __forceinline__ static float mul(float a, float b) { return a * b; }
  __device__ __forceinline__ static float add(float a, float b) { return a + b; }
  __device__ __forceinline__ static float sub(float a, float b) { return a - b; }
  __device__ __forceinline__ static float mm_sub(float a, float b, float c, float d) {
    return fmaf(a, b, -(c * d));
  }
  __device__ __forceinline__ static float mm_add(float a, float b, float c, float d) {
    return fmaf(a, b, c * d);
  }
  __device__ __forceinline__ static float acc_add(float acc, float a, float b) { return fmaf(a, b, acc); }
  __device__ __forceinline__ static float acc_sub(float acc, float a, float b) { return fmaf(-a, b, acc); }
  __device__ __forceinline__ static float dot3(float x, float y, float z) {
    return fmaf(z, z, fmaf(y, y, x * x));
  }
  __device__ __forceinline__ static float dot4(float a, float x, float y, float z) {
    return fmaf(z, z, fmaf(y, y, fmaf(x, x, a * a)));
  }
};

struct Turbo : Fast {
  static constexpr bool turbo = true;
  static constexpr bool count = false;
};
struct TurboCount : Fast {
  static constexpr bool turbo = true;
  static constexpr bool count = true;
};
// Flight policy of HEROAM_MODE_FAST_CLOSED: a force-free particle turns rigidly about its fixed axis, so a
// whole flight is ONE rotation (rotate_free below) instead of one orientation update per step.  Only the
// free-flight code is instantiated with it; the integrate kernels of that mode are Turbo's.
struct TurboClosed : Turbo {
  static constexpr bool closed_flight = true;
};
// Turbo that renormalises the orientation after every 8th step only (HEROAM_MODE_FAST_RELAXED)
struct Turbo8 : Turbo {
  static constexpr unsigned renorm_every = 8;
};

__device__ __forceinline__ float rsqrt_raw(float x) {  // one MUFU.RSQ, operands are never denormal here
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Conversions between the particle records (w, F) and the state a kernel keeps.
template <class P>
__device__ __forceinline__ void load_w(const Consts &c, const float *src, float w[3]) {
  w[0] = P::turbo ? __fmul_rn(src[0], c.half_dt) : src[0];
  w[1] = P::turbo ? __fmul_rn(src[1], c.half_dt) : src[1];
  w[2] = P::turbo ? __fmul_rn(src[2], c.half_dt) : src[2];
}
template <class P>
__device__ __forceinline__ void unscale_w(const Consts &c, const float w[3], float out[3]) {
  out[0] = P::turbo ? w[0] * c.inv_half_dt : w[0];
  out[1] = P::turbo ? w[1] * c.inv_half_dt : w[1];
  out[2] = P::turbo ? w[2] * c.inv_half_dt : w[2];
}
template <class P>
__device__ __forceinline__ void load_f(const Consts &c, const float *src, float f[3]) {
  f[0] = P::turbo ? __fmul_rn(src[0], c.kappa) : src[0];
  f[1] = P::turbo ? __fmul_rn(src[1], c.kappa) : src[1];
  f[2] = P::turbo ? __fmul_rn(src[2], c.kappa) : src[2];
}
template <class P>
__device__ __forceinline__ void unscale_f(const Consts &c, const float f[3], float out[3]) {
  out[0] = P::turbo ? f[0] * c.inv_kappa : f[0];
  out[1] = P::turbo ? f[1] * c.inv_kappa : f[1];
  out[2] = P::turbo ? f[2] * c.inv_kappa : f[2];
}

// ---------------------------------------------------------------------------------
// per-particle pieces of a step
// ---------------------------------------------------------------------------------
// (dt/2) * angular acceleration: find_accelration (simulation.c:343-352) followed by the
// scalar_multiply of update_angvel (:354-361).  The same value serves the second kick of
// one step and the first kick of the next (same position, same force).
template <class P>
__device__ __forceinline__ void half_kick(const Consts &c, const float r[3], const float f[3],
                                          float hk[3]) {
  const float tx = P::mm_sub(r[1], f[2], r[2], f[1]);
  const float ty = P::mm_sub(r[2], f[0], r[0], f[2]);
  const float tz = P::mm_sub(r[0], f[1], r[1], f[0]);
  if (P::exact) {
    hk[0] = P::mul(P::mul(tx, c.acc_scale), c.half_dt);
    hk[1] = P::mul(P::mul(ty, c.acc_scale), c.half_dt);
    hk[2] = P::mul(P::mul(tz, c.acc_scale), c.half_dt);
  } else if (P::turbo) {  // F' already carries (dt/2)^2/(m R^2): this is (dt/2) * (half kick)
    hk[0] = tx; hk[1] = ty; hk[2] = tz;
  } else {
    hk[0] = __fmul_rn(tx, c.kick);  // not contractable into w + hk
    hk[1] = __fmul_rn(ty, c.kick);
    hk[2] = __fmul_rn(tz, c.kick);
  }
}

// update_orientation (simulation.c:394-413) in two halves:
//   advance_orientation : q += (dt/2) (0,w) q ; q /= |q|
//   pole_position       : r = R * q (0,0,1) q^-1   (vector.c:94-110, :191-205, :182-188)
// The position depends on nothing but the new q, so a particle that is guaranteed to feel
// no force (free flight) only needs the first half per step and the second once at the end.
// A policy with renorm_every = 8 (Turbo8) renormalises the orientation only after every 8th
// step of a trajectory (absolute step index k with (k + 1) % 8 == 0).  q -> q + (dt/2) (0,w) q
// is linear in q, so the direction of q -- all the position depends on -- is the same whether
// the scale is taken out every step or every eighth; in between |q| wanders by float rounding
// only (~1e-7: the exact growth per step is 1 + |w dt/2|^2 / 2 ~ 1 + 1e-11), which moves the
// closed-form pole image by ~1e-5 A.  Keyed to the ABSOLUTE step, the result does not depend
// on how a trajectory's steps are cut into segments and kernels.  What it does change is the
// rounding pattern: the per-step roundings no longer mirror the reference's one for one, so the
// two orientations random-walk apart at the rate at which each random-walks away from the
// exact solution (measured: ~1.6e-3 A median after 2e5 steps instead of 2e-5 A).
constexpr uint32_t RENORM_ALIGN = 8;  // segment ends are multiples of this (the largest renorm_every)

// true if the step with absolute index `step` ends with a renormalisation
template <class P>
__device__ __forceinline__ bool renorm_at(uint32_t step) {
  return P::renorm_every == 1u || ((step + 1u) & (P::renorm_every - 1u)) == 0u;
}

template <class P>
__device__ __forceinline__ void advance_orientation(const Consts &c, const float w[3], float q[4], const bool renorm) {
  const float qa = q[0], qx = q[1], qy = q[2], qz = q[3];
  const float ox = w[0], oy = w[1], oz = w[2];
  if (P::exact) {
    // (0,w) (x) q, vector.c:166-171 with the zero scalar part dropped
    const float ta = P::sub(P::sub(-P::mul(ox, qx), P::mul(oy, qy)), P::mul(oz, qz));
    const float tx = P::sub(P::add(P::mul(ox, qa), P::mul(oy, qz)), P::mul(oz, qy));
    const float ty = P::add(P::sub(P::mul(oy, qa), P::mul(ox, qz)), P::mul(oz, qx));
    const float tz = P::add(P::sub(P::mul(ox, qy), P::mul(oy, qx)), P::mul(oz, qa));
    float na = P::add(qa, P::mul(ta, c.half_dt));
    float nx = P::add(qx, P::mul(tx, c.half_dt));
    float ny = P::add(qy, P::mul(ty, c.half_dt));
    float nz = P::add(qz, P::mul(tz, c.half_dt));
    const float inv = __frcp_rn(__fsqrt_rn(P::dot4(na, nx, ny, nz)));
    q[0] = P::mul(na, inv); q[1] = P::mul(nx, inv); q[2] = P::mul(ny, inv); q[3] = P::mul(nz, inv);
  } else if (P::turbo) {
    // w' = (dt/2) w.  The increment w' (x) q (~1e-6 |q|) is formed on its own, at full relative
    // precision, and added to q with ONE rounding, exactly as the reference does: accumulating
    // the three products straight into q would round three times per step at ulp(q) and
    // quadruple the random walk of the orientation (measured: 1.6e-3 A instead of 4e-4 A after
    // 2000 steps on R = 47.8 A).
    const float na = qa - fmaf(oz, qz, fmaf(oy, qy, ox * qx));
    const float nx = qx + fmaf(-oz, qy, fmaf(oy, qz, ox * qa));
    const float ny = qy + fmaf(oz, qx, fmaf(-ox, qz, oy * qa));
    const float nz = qz + fmaf(oz, qa, fmaf(-oy, qx, ox * qy));
    q[0] = na; q[1] = nx; q[2] = ny; q[3] = nz;
    if (renorm) {  // a real branch: uniform across the warp whenever its trajectories are in phase
      const float inv = rsqrt_raw(fmaf(nz, nz, fmaf(ny, ny, fmaf(nx, nx, na * na))));
      // __fmul_rn: a plain product could be contracted by the compiler into whatever addition
      // consumes it next (differently in different kernels); the state must not depend on that
      q[0] = __fmul_rn(na, inv); q[1] = __fmul_rn(nx, inv); q[2] = __fmul_rn(ny, inv); q[3] = __fmul_rn(nz, inv);
    }
  } else {
    const float ta = -fmaf(oz, qz, fmaf(oy, qy, ox * qx));
    const float tx = fmaf(-oz, qy, fmaf(oy, qz, ox * qa));
    const float ty = fmaf(oz, qx, fmaf(-ox, qz, oy * qa));
    const float tz = fmaf(oz, qa, fmaf(-oy, qx, ox * qy));
    const float na = fmaf(ta, c.half_dt, qa);
    const float nx = fmaf(tx, c.half_dt, qx);
    const float ny = fmaf(ty, c.half_dt, qy);
    const float nz = fmaf(tz, c.half_dt, qz);
    const float inv = rsqrtf(P::dot4(na, nx, ny, nz));
    q[0] = __fmul_rn(na, inv); q[1] = __fmul_rn(nx, inv); q[2] = __fmul_rn(ny, inv); q[3] = __fmul_rn(nz, inv);
  }
}

// k force-free steps at once (HEROAM_MODE_FAST_CLOSED).  One step maps q to normalize(q + W q) with
// W = (0, w'), w' = (dt/2) w: since q + W q = (1 + W) q and 1 + W = sqrt(1 + |w'|^2) (cos t + n sin t) with
// n = w'/|w'| and tan t = |w'|, the normalised step is the rotor (cos t, n sin t) applied from the left,
// and k steps are the rotor (cos kt, n sin kt).  This is the exact solution of the map the reference
// iterates; what it does not reproduce is the reference's per-step rounding, so positions agree with
// the stepped arithmetic to the size of its own rounding walk (tests/horizon_common.py, "fast_closed").
// Single precision suffices: |w'| ~ 1e-5, k t <= ~1e2 rad over a whole trajectory, so the angle is
// good to ~1e-5 rad (5e-4 A on R = 47.8 A) in the worst case and to ~1e-7 rad for a typical flight.
__device__ __forceinline__ void rotate_free(const float w[3], float q[4], uint32_t k) {
  const float s2 = fmaf(w[2], w[2], fmaf(w[1], w[1], w[0] * w[0]));
  if (!(s2 > 0.0f) || k == 0u) return;
  const float s = sqrtf(s2);
  float sn, cs;
  sincosf(__uint2float_rn(k) * atanf(s), &sn, &cs);
  const float g = sn / s, ax = w[0] * g, ay = w[1] * g, az = w[2] * g;
  const float qa = q[0], qx = q[1], qy = q[2], qz = q[3];
  const float na = fmaf(cs, qa, -fmaf(az, qz, fmaf(ay, qy, ax * qx)));
  const float nx = fmaf(cs, qx, fmaf(-az, qy, fmaf(ay, qz, ax * qa)));
  const float ny = fmaf(cs, qy, fmaf(az, qx, fmaf(-ax, qz, ay * qa)));
  const float nz = fmaf(cs, qz, fmaf(az, qa, fmaf(-ay, qx, ax * qy)));
  const float inv = rsqrtf(fmaf(nz, nz, fmaf(ny, ny, fmaf(nx, nx, na * na))));
  q[0] = __fmul_rn(na, inv); q[1] = __fmul_rn(nx, inv); q[2] = __fmul_rn(ny, inv); q[3] = __fmul_rn(nz, inv);
}

// the renormalisation on its own (Turbo): lets a kernel that holds several particles take ONE
// branch per step around all of them
__device__ __forceinline__ void renormalise(float q[4]) {
  const float inv = rsqrt_raw(fmaf(q[3], q[3], fmaf(q[2], q[2], fmaf(q[1], q[1], q[0] * q[0]))));
  q[0] = __fmul_rn(q[0], inv); q[1] = __fmul_rn(q[1], inv); q[2] = __fmul_rn(q[2], inv); q[3] = __fmul_rn(q[3], inv);
}

template <class P>
__device__ __forceinline__ void pole_position(const Consts &c, const float q[4], float r[3]) {
  const float na = q[0], nx = q[1], ny = q[2], nz = q[3];
  if (P::exact) {
    // q^-1 = conj(q)/|q|^2 ; t = q (x) (0,0,0,1) = (-qz, qy, -qx, qa) ; res = t (x) q^-1
    const float n2 = P::dot4(na, nx, ny, nz);
    const float ia = __fdiv_rn(na, n2);
    const float ix = -__fdiv_rn(nx, n2);
    const float iy = -__fdiv_rn(ny, n2);
    const float iz = -__fdiv_rn(nz, n2);
    const float ua = -nz, ux = ny, uy = -nx, uz = na;
    const float rx = P::sub(P::add(P::add(P::mul(ua, ix), P::mul(ux, ia)), P::mul(uy, iz)), P::mul(uz, iy));
    const float ry = P::add(P::add(P::sub(P::mul(ua, iy), P::mul(ux, iz)), P::mul(uy, ia)), P::mul(uz, ix));
    const float rz = P::add(P::sub(P::add(P::mul(ua, iz), P::mul(ux, iy)), P::mul(uy, ix)), P::mul(uz, ia));
    r[0] = P::mul(rx, c.radius);
    r[1] = P::mul(ry, c.radius);
    r[2] = P::mul(rz, c.radius);
  } else if (P::turbo) {  // unit q: a^2 - x^2 - y^2 + z^2 = 1 - 2 (x^2 + y^2)
    r[0] = __fmul_rn(c.radius2, fmaf(nx, nz, na * ny));  // never contracted into the pair separations
    r[1] = __fmul_rn(c.radius2, fmaf(ny, nz, -(na * nx)));
    r[2] = fmaf(-c.radius2, fmaf(nx, nx, ny * ny), c.radius);
  } else {
    // unit q: q (0,0,1) q^-1 = (2(xz + ay), 2(yz - ax), a^2 - x^2 - y^2 + z^2)
    r[0] = __fmul_rn(c.radius2, fmaf(nx, nz, na * ny));
    r[1] = __fmul_rn(c.radius2, fmaf(ny, nz, -(na * nx)));
    r[2] = __fmul_rn(c.radius, __fsub_rn(fmaf(na, na, nz * nz), fmaf(nx, nx, ny * ny)));
  }
}

template <class P>
__device__ __forceinline__ void drift(const Consts &c, const float w[3], float q[4], float r[3], const bool renorm) {
  advance_orientation<P>(c, w, q, renorm);
  pole_position<P>(c, q, r);
}

// One pair (simulation.c:192-197 and :216-229).  Returns d; adds the table force to
// fj / subtracts it from fk when the squared distance falls inside the table.
// Out-of-table (either side) means "no force" (SURVEY.md Q5).
template <class P>
__device__ __forceinline__ float pair_force(const Consts &c, const float *__restrict__ table,
                                            const float rj[3], const float rk[3], float fj[3],
                                            float fk[3], uint32_t &inrange) {
  const float dx = P::sub(rj[0], rk[0]);
  const float dy = P::sub(rj[1], rk[1]);
  const float dz = P::sub(rj[2], rk[2]);
  const float d = P::dot3(dx, dy, dz);
  // bin = (int)((d - start) * inv_grid) in float, as the reference (Q6); a negative or too
  // large bin means "outside the table"
  const float pos = __fmul_rn(__fsub_rn(d, c.grid_start), c.inv_grid);
  const unsigned bin = (unsigned)__float2int_rz(pos);
  if (bin < (unsigned)c.grid_len) {
    ++inrange;
    const float tv = __ldg(table + bin);
    float scale;
    if (P::exact) {
      scale = __double2float_rn(__ddiv_rn((double)tv, __dsqrt_rn((double)d)));
      const float fx = P::mul(scale, dx), fy = P::mul(scale, dy), fz = P::mul(scale, dz);
      fj[0] = P::add(fj[0], fx); fj[1] = P::add(fj[1], fy); fj[2] = P::add(fj[2], fz);
      fk[0] = P::sub(fk[0], fx); fk[1] = P::sub(fk[1], fy); fk[2] = P::sub(fk[2], fz);
    } else {
      scale = tv * rsqrtf(d);
      fj[0] = fmaf(scale, dx, fj[0]); fj[1] = fmaf(scale, dy, fj[1]); fj[2] = fmaf(scale, dz, fj[2]);
      fk[0] = fmaf(-scale, dx, fk[0]); fk[1] = fmaf(-scale, dy, fk[1]); fk[2] = fmaf(-scale, dz, fk[2]);
    }
  }
  return d;
}

// The same pair split in three branch-free stages so that a kernel can issue the table
// gathers of many pairs back to back and only then consume them (memory-level parallelism
// instead of one exposed L1/L2 latency per pair):
//   pair_geom   : separation, squared distance, table bin (clamped to 0 when outside)
//   pair_gather : the table value, 0 outside the table
//   pair_scale  : Force_list[bin] / sqrt(d)   (simulation.c:220)
// A pair outside the table contributes scale == 0, i.e. adds +-0 to both forces, which
// leaves every value as the reference's "skip" does.
struct PairGeom {
  float dx, dy, dz, d;
  unsigned bin;  // table index, clamped to grid_len (the zero entry appended to the table) when outside
  bool inside;   // only maintained when the policy counts work
};

template <class P>
__device__ __forceinline__ PairGeom pair_geom(const Consts &c, const float rj[3], const float rk[3]) {
  PairGeom g;
  g.dx = P::sub(rj[0], rk[0]);
  g.dy = P::sub(rj[1], rk[1]);
  g.dz = P::sub(rj[2], rk[2]);
  g.d = P::dot3(g.dx, g.dy, g.dz);
  // Exact / Fast: (d - start) * inv_grid in float exactly as the reference (Q6).  The float ->
  // int conversion truncates (a position in (-1, 0) is bin 0, as in the reference); seen as
  // unsigned, a negative bin is huge, so one unsigned min sends both sides of the table to
  // the appended zero entry.
  const float pos = P::turbo ? fmaf(g.d, c.inv_grid, c.pos_bias) : __fmul_rn(__fsub_rn(g.d, c.grid_start), c.inv_grid);
  g.bin = min((unsigned)__float2int_rz(pos), (unsigned)c.grid_len);
  g.inside = P::count ? g.bin < (unsigned)c.grid_len : false;
  return g;
}

// the table value (0 outside the table); Turbo reads the copy scaled by kappa
template <class P>
__device__ __forceinline__ float pair_gather(const DevView &v, const PairGeom &g) {
  return __ldg((P::turbo ? v.table_k : v.table) + g.bin);
}

template <class P>
__device__ __forceinline__ float pair_scale(float tv, float d) {
  if (P::exact) return __double2float_rn(__ddiv_rn((double)tv, __dsqrt_rn((double)d)));
  if (P::turbo) return tv * rsqrt_raw(d);
  return tv * rsqrtf(d);
}

// F_j += scale * (r_j - r_k) ; F_k -= scale * (r_j - r_k)   (simulation.c:221-229)
template <class P>
__device__ __forceinline__ void pair_apply(float scale, const PairGeom &g, float fj[3], float fk[3]) {
  if (P::exact) {
    const float fx = P::mul(scale, g.dx), fy = P::mul(scale, g.dy), fz = P::mul(scale, g.dz);
    fj[0] = P::add(fj[0], fx); fj[1] = P::add(fj[1], fy); fj[2] = P::add(fj[2], fz);
    fk[0] = P::sub(fk[0], fx); fk[1] = P::sub(fk[1], fy); fk[2] = P::sub(fk[2], fz);
  } else {
    fj[0] = fmaf(scale, g.dx, fj[0]); fj[1] = fmaf(scale, g.dy, fj[1]); fj[2] = fmaf(scale, g.dz, fj[2]);
    fk[0] = fmaf(-scale, g.dx, fk[0]); fk[1] = fmaf(-scale, g.dy, fk[1]); fk[2] = fmaf(-scale, g.dz, fk[2]);
  }
}
// one-sided form for kernels in which every particle accumulates its own force
template <class P>
__device__ __forceinline__ void pair_apply_own(float scale, const PairGeom &g, float fj[3]) {
  if (P::exact) {
    fj[0] = P::add(fj[0], P::mul(scale, g.dx));
    fj[1] = P::add(fj[1], P::mul(scale, g.dy));
    fj[2] = P::add(fj[2], P::mul(scale, g.dz));
  } else {
    fj[0] = fmaf(scale, g.dx, fj[0]); fj[1] = fmaf(scale, g.dy, fj[1]); fj[2] = fmaf(scale, g.dz, fj[2]);
  }
}

// squared distance only (meeting sweep, simulation.c:192-196)
template <class P>
__device__ __forceinline__ float pair_dist_sq(const float rj[3], const float rk[3]) {
  return P::dot3(P::sub(rj[0], rk[0]), P::sub(rj[1], rk[1]), P::sub(rj[2], rk[2]));
}

// stored velocity r x w (simulation.c:504) and its square (:506)
template <class P>
__device__ __forceinline__ void stored_velocity(const float r[3], const float w[3], float v[3], float &v2) {
  v[0] = P::mm_sub(r[1], w[2], r[2], w[1]);
  v[1] = P::mm_sub(r[2], w[0], r[0], w[2]);
  v[2] = P::mm_sub(r[0], w[1], r[1], w[0]);
  v2 = P::dot3(v[0], v[1], v[2]);
}

// |f| as norm() of vector.c:24-26
__device__ __forceinline__ float vec_norm(const float f[3]) {
  return __fsqrt_rn(Exact::dot3(f[0], f[1], f[2]));
}

// find_success (simulation.c:33-46): equality, not >= (Q4)
__device__ __forceinline__ bool all_paired(uint32_t no, uint32_t flagged) {
  return (no & 1u) ? (flagged == no - 1u) : (flagged == no);
}

}  // namespace heroam
