// heroam_kernels.cuh -- CUDA kernels of the trajectory engine (sm_100a).
//
// Work decomposition:
//   * a trajectory whose `a` particles are still moving behaves exactly like an
//     a-particle system: frozen particles neither exert nor feel force and cannot be
//     met again (reference simulation.c:181, :191, :207, :215);
//   * k_integrate_thread<A>: ONE THREAD PER TRAJECTORY, the a == A active particles
//     entirely in registers, all loops unrolled, no shared memory, no shuffles: the
//     all-pairs loop runs in the reference's (j,k) order so forces accumulate in the
//     same order.  It advances a trajectory by up to `segment` steps and stops early
//     the moment any pair comes within icd_dist;
//   * that step (rare) is completed by k_resolve_and_bin, which runs the reference's
//     greedy order-dependent meeting sweep (Q3), freezes particles, evaluates the
//     termination rule (Q4) and sorts live trajectories into bins by active count for
//     the next pass;
//   * trajectories with more than THREAD_MAX_A active particles go to a wider kernel.
#pragma once

#include "heroam_device.cuh"

namespace heroam {

// Bins: a live trajectory is binned by the number of its particles still moving.
constexpr int THREAD_MAX_A = 8;   // bins 1..8  : k_integrate_thread<A>, one thread per trajectory, all state in registers
constexpr int HYBRID_MAX_A = 12;  // bins 9..12 : k_integrate_hybrid<A>, one thread per trajectory, positions and forces in
                                  //              registers, orientation / angular velocities in shared memory
constexpr int BIN_SMEM16 = 13;    // 13..16 active: k_integrate_smem<16>, one thread per trajectory, state in shared memory
constexpr int BIN_WARP1 = 14;     // 17..32 active: k_integrate_warp<1>, one warp per trajectory
constexpr int BIN_WARP2 = 15;     // 33..64 active: k_integrate_warp<2> (two particles per lane)
constexpr int BIN_WARP4 = 16;     // 65..128 active: k_integrate_warp<4>
// every pair provably out of reach for a whole segment: k_free_flight, one thread per PARTICLE.
// Three tiers of FIXED segment length (so that all lanes of a warp run the same trip count):
constexpr int BIN_FREE = 17;      // BIN_FREE + 0: short, + 1: medium, + 2: long (whole trajectory in free flight)
constexpr int FREE_SPLIT = 3;     // BIN_FREE + 3: loners split off a trajectory whose core keeps interacting
constexpr int N_FREE_TIERS = 4;
constexpr int BIN_CORE = BIN_FREE + N_FREE_TIERS;  // BIN_CORE + a: cores of split trajectories, a = 1..8 (k_integrate_thread<a>)
constexpr int N_BINS = BIN_CORE + 9;
constexpr int LIVE_SLOT = N_BINS;         // bin_count[LIVE_SLOT]: live trajectories seen by the resolve kernel
constexpr int LAG_SLOT = N_BINS + 1;      // bin_count[LAG_SLOT]: 0xffffffff - (fewest steps done by a live trajectory)
constexpr int PART_LIVE_SLOT = N_BINS + 2; // bin_count[PART_LIVE_SLOT]: live trajectories of this part of a partitioned pool
constexpr int N_BIN_COUNTERS = N_BINS + 3;
constexpr int MAX_PARTS = 32;              // parts of a partitioned pool (a droplet-size sweep as one run)
constexpr int FREE_MAX_A = 8;     // the free-flight analysis is done for trajectories with <= 8 moving particles
constexpr int MAX_ACTIVE = 128;   // wider trajectories are refused (HEROAM_DONE_UNSUPPORTED)
constexpr int INTEGRATE_TPB = 128;  // 64 was measured too (round 2): no gain
constexpr int SMEM_TPB = 64;          // threads (= trajectories) per block of k_integrate_smem / k_integrate_hybrid
constexpr int WARP_KERNEL_WARPS = 4;  // warps (= trajectories) per block of k_integrate_warp

__host__ __device__ inline int bin_of(uint32_t active) {
  if (active <= (uint32_t)HYBRID_MAX_A) return (int)active;
  if (active <= 16u) return BIN_SMEM16;
  if (active <= 32u) return BIN_WARP1;
  if (active <= 64u) return BIN_WARP2;
  return BIN_WARP4;
}

__device__ __forceinline__ void ld3(const float *src, float dst[3]) {
  dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
}
__device__ __forceinline__ void st3(float *dst, const float src[3]) {
  dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------
// Record formats written by the integrate kernels (w, wh, f are in the kernel's state
// units: see load_w / load_f).
// ---------------------------------------------------------------------------------
// A step abandoned after the drift because a pair came within icd_dist (see finish_step).
template <class P>
__device__ __forceinline__ void write_midstep(const Consts &c, heroam_particle *p, const float q[4], const float r[3],
                                              const float w[3], const float wh[3], const float r_old[3], float time,
                                              bool stepped_before) {
  float w_rec[3], wh_rec[3];
  unscale_w<P>(c, w, w_rec);
  unscale_w<P>(c, wh, wh_rec);
  p->orientation[0] = q[0]; p->orientation[1] = q[1]; p->orientation[2] = q[2]; p->orientation[3] = q[3];
  st3(p->position, r);
  st3(p->ang_velocity, w_rec);  // w(t): what a particle frozen in this step keeps
  st3(p->force, wh_rec);        // scratch: w(t + dt/2)
  p->time = time;
  if (stepped_before) {         // else the imported velocity fields stand
    float vel[3], v2;
    stored_velocity<P>(r_old, w_rec, vel, v2);
    st3(p->velocity, vel);
    p->velocity_sq = v2;
  }
}
// Clean state at a step boundary, every field as the reference holds it.
template <class P>
__device__ __forceinline__ void write_clean(const Consts &c, heroam_particle *p, const float q[4], const float r[3],
                                            const float w[3], const float f[3], float time, bool has_force_mag) {
  float w_rec[3], f_rec[3], vel[3], v2;
  unscale_w<P>(c, w, w_rec);
  unscale_f<P>(c, f, f_rec);
  p->orientation[0] = q[0]; p->orientation[1] = q[1]; p->orientation[2] = q[2]; p->orientation[3] = q[3];
  st3(p->position, r);
  st3(p->ang_velocity, w_rec);
  st3(p->force, f_rec);
  stored_velocity<P>(r, w_rec, vel, v2);
  st3(p->velocity, vel);
  p->velocity_sq = v2;
  p->time = time;
  // force_mag only exists for indices below no - 1 (simulation.c:205, :232; Q9)
  p->force_mag = has_force_mag ? vec_norm(f_rec) : 0.0f;
}

// ---------------------------------------------------------------------------------
// k_prepare: imported records -> working state.  time := 0 (Q1); frozen-at-import
// particles get the zero force the first calculate_force call would give them
// (simulation.c:170-176).
// ---------------------------------------------------------------------------------
__global__ void k_prepare(DevView v, const uint32_t *__restrict__ first,
                          const uint32_t *__restrict__ counts) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= v.n_traj) return;
  TrajMeta m;
  m.first = first[t];
  m.no = counts[t];
  m.steps = 0;
  m.time = 0.0f;
  m.n_flagged = 0;
  for (uint32_t j = 0; j < m.no; ++j) {
    heroam_particle *p = v.particles + m.first + j;
    p->time = 0.0f;
    if (p->success) {
      p->success = 1;
      ++m.n_flagged;
      if (v.c.max_steps > 0) {
        p->force[0] = 0.0f; p->force[1] = 0.0f; p->force[2] = 0.0f;
        p->force_mag = 0.0f;
      }
    }
  }
  m.n_active = m.no - m.n_flagged;
  m.status = ST_LIVE;
  m.step_stop = 0;
  m.n_loners = 0; m.loner_from = 0; m.loner_until = 0; m.loner_time0 = 0.0f; m.cross_from = 0; m.core_slots = m.n_active;
  m.point = 0;
  v.meta[t] = m;
}

// ---------------------------------------------------------------------------------
// finish_step: complete a step that k_integrate_* abandoned after the drift because a
// pair came within icd_dist.  Record format on entry (per active particle):
//   orientation, position : after the drift            ang_velocity : w(t)  (pre-step)
//   velocity, velocity_sq : r(t) x w(t) and its square (or the imported values if the
//                           trajectory has not completed a step yet)
//   force                 : scratch, holds w(t + dt/2)   time : t
// Does calculate_force (simulation.c:158-236) and the second half of the step
// (:496-509) straight on the records, in the reference's order.
// ---------------------------------------------------------------------------------
template <class P>
__device__ void finish_step(const DevView &v, TrajMeta &m, Counters &work) {
  const Consts &c = v.c;
  heroam_particle *base = v.particles + m.first;
  const uint32_t *act = v.act + m.first;
  const uint32_t a = m.n_active;
  // pass 1: greedy meeting sweep (:178-202)
  for (uint32_t js = 0; js < a; ++js) {
    heroam_particle *pj = base + act[js];
    if (pj->success) continue;
    for (uint32_t ks = js + 1; ks < a; ++ks) {
      heroam_particle *pk = base + act[ks];
      if (pk->success) continue;
      if (pair_dist_sq<P>(pj->position, pk->position) < c.icd2) {
        pj->success = 1;
        pk->success = 1;
      }
    }
  }
  // survivors: w(t+dt/2) moves from the scratch slot to ang_velocity; everybody's force := 0
  uint32_t remaining = 0;
  for (uint32_t s = 0; s < a; ++s) {
    heroam_particle *p = base + act[s];
    if (!p->success) {
      ++remaining;
      p->ang_velocity[0] = p->force[0];
      p->ang_velocity[1] = p->force[1];
      p->ang_velocity[2] = p->force[2];
    }
    p->force[0] = 0.0f; p->force[1] = 0.0f; p->force[2] = 0.0f;
    p->force_mag = 0.0f;
  }
  // pass 2: forces among the survivors (:205-233)
  uint32_t inrange = 0, pairs = 0;
  for (uint32_t js = 0; js < a; ++js) {
    heroam_particle *pj = base + act[js];
    if (pj->success) continue;
    if (act[js] + 1u >= m.no) continue;  // the loop bound j < no - 1
    for (uint32_t ks = js + 1; ks < a; ++ks) {
      heroam_particle *pk = base + act[ks];
      if (pk->success) continue;
      ++pairs;
      float fj[3], fk[3];
      ld3(pj->force, fj);
      ld3(pk->force, fk);
      pair_force<P>(c, v.table, pj->position, pk->position, fj, fk, inrange);
      st3(pj->force, fj);
      st3(pk->force, fk);
    }
    pj->force_mag = vec_norm(pj->force);
  }
  // second kick, stored velocity, clock (:496-509)
  const float t_next = __fadd_rn(m.time, c.dt);
  for (uint32_t s = 0; s < a; ++s) {
    heroam_particle *p = base + act[s];
    if (p->success) continue;
    float hk[3], w[3], vel[3], v2;
    half_kick<P>(c, p->position, p->force, hk);
    w[0] = P::add(p->ang_velocity[0], hk[0]);
    w[1] = P::add(p->ang_velocity[1], hk[1]);
    w[2] = P::add(p->ang_velocity[2], hk[2]);
    st3(p->ang_velocity, w);
    stored_velocity<P>(p->position, w, vel, v2);
    st3(p->velocity, vel);
    p->velocity_sq = v2;
    p->time = t_next;
  }
  m.n_flagged += a - remaining;
  m.n_active = remaining;
  m.steps += 1;
  if (remaining > 0 || m.n_loners > 0) m.time = t_next;  // :512-517, max over particle clocks (loners move too)
  m.status = ST_LIVE;
  work.particle_steps += remaining;
  work.pair_steps += pairs;
  work.inrange_steps += inrange;
}

// ---------------------------------------------------------------------------------
// Pieces of the per-pass bookkeeping, shared by k_resolve_and_bin (fixed batch) and
// k_pool_resolve (streaming pool, heroam_sampler.cuh).
// ---------------------------------------------------------------------------------
constexpr uint32_t NOT_DONE = 0xffffffffu;

// The loop condition `!success && time < max_time` (simulation.c:469) plus the defined
// exits (deadlock, step cap, unsupported width).  Returns a HEROAM_DONE_* code or NOT_DONE.
__device__ inline uint32_t evaluate_exit(const Consts &c, TrajMeta &m, uint32_t step_cap) {
  const uint32_t moving = m.n_active + m.n_loners;
  uint32_t done = NOT_DONE;
  if (m.steps == 0) {
    if (!(0.0f < c.max_time)) {
      done = HEROAM_DONE_MAXTIME;  // the reference's loop body never runs
    } else if (step_cap == 0u) {
      done = HEROAM_DONE_STEPCAP;
    } else if (moving == 0) {
      // a step with nobody to move: forces are zeroed at import, no clock advances
      m.steps = 1;
      done = all_paired(m.no, m.n_flagged) ? HEROAM_DONE_SUCCESS : HEROAM_DONE_DEADLOCK;
    }
  } else if (all_paired(m.no, m.n_flagged)) {
    done = HEROAM_DONE_SUCCESS;
  } else if (!(m.time < c.max_time)) {
    done = HEROAM_DONE_MAXTIME;
  } else if (moving == 0) {
    done = HEROAM_DONE_DEADLOCK;
  } else if (m.steps >= step_cap) {
    done = HEROAM_DONE_STEPCAP;
  }
  if (done == NOT_DONE && moving > (uint32_t)MAX_ACTIVE) done = HEROAM_DONE_UNSUPPORTED;
  return done;
}

// Steps a trajectory may advance per pass, by bin.  All bin kernels of a pass run concurrently
// and the pass ends with the slowest, so bins whose step has a long dependency chain (many
// particles per thread) get shorter segments and the cheap free-flight steps longer ones.
struct Segments {
  uint32_t of_bin[N_BINS];
  uint32_t free_stride;   // entries reserved per whole-trajectory free-flight tier in the free list
  uint32_t drain;         // 1 while the run is draining (few live trajectories, latency bound):
                          // the short free-flight tier would only slow per-pass progress
  uint32_t no_flight;     // 1: per-step capture -- every particle's record must be current at every step
                          // boundary, so nobody flies and nothing is split off
  uint32_t nosplit_max;   // trajectories with at most this many moving particles are not split into core and loners
};

__host__ inline Segments make_segments(uint32_t base, bool drain, uint32_t split_window_max) {
  Segments sg;
  auto scaled = [base](uint32_t num, uint32_t den) { uint32_t v = (uint32_t)((unsigned long long)base * num / den); return v ? v : 1u; };
  sg.of_bin[0] = base;
  if (!drain) {
    // throughput-bound regime (the GPU is full): pass time is the sum of the kernels' work
    sg.of_bin[1] = sg.of_bin[2] = scaled(2, 1);
    sg.of_bin[3] = sg.of_bin[4] = base;
    sg.of_bin[5] = sg.of_bin[6] = scaled(1, 2);
    sg.of_bin[7] = sg.of_bin[8] = scaled(1, 4);
    for (int b = THREAD_MAX_A + 1; b <= HYBRID_MAX_A; ++b) sg.of_bin[b] = scaled(1, 8);
    sg.of_bin[BIN_SMEM16] = scaled(1, 8);
    sg.of_bin[BIN_WARP1] = sg.of_bin[BIN_WARP2] = sg.of_bin[BIN_WARP4] = scaled(1, 8);
  } else {
    // latency-bound regime (less than a wave of threads): every kernel lasts segment x
    // chain latency of one step and the pass lasts as long as the slowest, so give each bin
    // about the same wall time (~0.3 us per step at 2 particles, growing with the pair count)
    sg.of_bin[1] = sg.of_bin[2] = sg.of_bin[3] = scaled(2, 1);
    sg.of_bin[4] = scaled(3, 2);
    sg.of_bin[5] = sg.of_bin[6] = base;
    sg.of_bin[7] = sg.of_bin[8] = scaled(1, 2);
    for (int b = THREAD_MAX_A + 1; b <= HYBRID_MAX_A; ++b) sg.of_bin[b] = scaled(1, 4);
    sg.of_bin[BIN_SMEM16] = scaled(1, 8);
    sg.of_bin[BIN_WARP1] = sg.of_bin[BIN_WARP2] = sg.of_bin[BIN_WARP4] = scaled(1, 8);
  }
  // free flight: ~73 ns per step; the short tier is dropped while draining (schedule_segment)
  sg.of_bin[BIN_FREE + 0] = scaled(1, 4);
  sg.of_bin[BIN_FREE + 1] = scaled(2, 1);
  // While draining a pass lasts as long as its slowest kernel, and every trajectory advances one
  // segment per pass: a long tier of 8x (65 536 steps x 73 ns = 4.8 ms for a hundred flyers) would hold
  // up the thousands of two-particle trajectories whose segment takes half of that.
  sg.of_bin[BIN_FREE + 2] = drain ? scaled(4, 1) : scaled(8, 1);
  // split trajectories: core and loners advance by the same fixed window.  The reach bound grows
  // with the square of the window (possible acceleration), so beyond split_window_max =
  // O(sqrt(m / f_max) / dt) everything would look linked to everything.
  uint32_t window = scaled(1, 2);
  while (window > split_window_max && window > 64u) window /= 2u;
  sg.of_bin[BIN_FREE + FREE_SPLIT] = window;
  for (int a = 0; a <= THREAD_MAX_A; ++a) sg.of_bin[BIN_CORE + a] = window;
  sg.free_stride = 0;
  sg.drain = drain ? 1u : 0u;
  sg.nosplit_max = drain ? (uint32_t)THREAD_MAX_A : 0u;
  sg.no_flight = 0;
  return sg;
}

// per-step capture: every bin advances by the same number of steps (one chunk of snapshots)
__host__ inline Segments make_capture_segments(uint32_t chunk) {
  Segments sg;
  for (int b = 0; b < N_BINS; ++b) sg.of_bin[b] = chunk;
  sg.free_stride = 0;
  sg.drain = 0;
  sg.nosplit_max = 0;
  sg.no_flight = 1;
  return sg;
}

// ---------------------------------------------------------------------------------
// Reach analysis.  A particle with angular velocity w moves its position by at most |w| R dt
// per step (chord <= arc; the rotation angle of a step is 2 atan(|w| dt / 2) <= |w| dt) plus
// rounding of the re-projection (a few ulp of R); under forces its angular speed can grow by
// at most dt * f_max / (m R) per step and partner within reach (|torque| <= R |F|, and an
// unmet pair is never closer than icd_dist, so |F| <= f_max).  Over W steps it travels at most
//     reach = R dt (|w| W + dt deg f_max / (m R) W^2 / 2) + slack,
// deg = number of partners that could come within D_far of it, which itself follows from the
// reaches: link(j,k) <=> dist(j,k) - D_far <= reach(j) + reach(k); iterate to the (monotone)
// fixed point.  Pairs that are not linked are provably out of the table and out of icd_dist
// for the whole window: they contribute exact zeros and cannot meet.  Evaluated in double
// with 1 % slack on speeds and forces.
// ---------------------------------------------------------------------------------
struct Reach {
  const heroam_particle *base;
  const uint32_t *act;
  uint32_t a;
  // single precision is enough: every bound carries 1 % (speeds, forces) plus absolute slack,
  // orders of magnitude above float rounding (1e-7 relative, < 1e-4 A on these distances)
  float RdtW, accel, slack, d_far;
  float w0[MAX_ACTIVE];
  uint8_t deg[MAX_ACTIVE];
  // Candidate pairs: the pairs that would be linked if every particle had the largest reach it can have (every other
  // particle a partner).  Only they can ever be linked, so the fixed point is iterated over them alone: one sweep over
  // all pairs per analysis instead of one per iteration.  On a large droplet (config C3: 49 particles on R = 478 A,
  // 99 % of the pairs out of reach) that is a handful of pairs out of 1176, and the resolve kernel -- one thread per
  // trajectory, 59 % of the run before -- no longer dominates it.  Dense trajectories that overflow the list keep the
  // all-pairs iteration.
  static constexpr uint32_t CAND_CAP = 320;
  // positions copied once into the thread's own (lane-interleaved, hence coalesced) local memory: the sweep over
  // all pairs would otherwise re-read the 84-byte records of 32 different trajectories per warp instruction
  float px[MAX_ACTIVE], py[MAX_ACTIVE], pz[MAX_ACTIVE];
  uint8_t cj[CAND_CAP], ck[CAND_CAP];
  float cgap[CAND_CAP];
  uint32_t n_cand;
  bool dense;  // the candidates did not fit: iterate over all pairs

  __device__ float gap_of(uint32_t j, uint32_t k) const {
    const float dx = px[j] - px[k], dy = py[j] - py[k], dz = pz[j] - pz[k];
    return sqrtf(dx * dx + dy * dy + dz * dz) - d_far;
  }
  __device__ void init(const DevView &v, const TrajMeta &m, uint32_t count, uint32_t window) {
    base = v.particles + m.first;
    act = v.act + m.first;
    a = count;
    const float R = v.c.radius, dt = v.c.dt, W = (float)window;
    d_far = sqrtf(v.c.far2) * 1.00001f;
    RdtW = R * dt * W * 1.01f;                                          // path per unit angular speed
    accel = 0.5f * R * dt * dt * v.c.f_max * v.c.inv_mr * W * W * 1.02f; // extra path per linked partner
    slack = 2.0e-6f * R * W + 2.0e-3f;
    for (uint32_t s = 0; s < a; ++s) {
      const heroam_particle *p = base + act[s];
      const float *w = p->ang_velocity;
      w0[s] = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
      px[s] = p->position[0]; py[s] = p->position[1]; pz[s] = p->position[2];
      deg[s] = 0;
    }
    // one sweep over all pairs, squared distances against the squared largest possible limit (with margin: the
    // exact test below decides, this only has to be a superset)
    n_cand = 0;
    dense = false;
    const float most = (float)(a - 1u) * accel + slack;
    for (uint32_t j = 0; j + 1 < a && !dense; ++j) {
      const float xj = px[j], yj = py[j], zj = pz[j], lj = d_far + fmaf(w0[j], RdtW, most);
      for (uint32_t k = j + 1; k < a; ++k) {
        const float dx = xj - px[k], dy = yj - py[k], dz = zj - pz[k];
        const float lim = (lj + fmaf(w0[k], RdtW, most)) * 1.0001f;
        const float d2 = dx * dx + dy * dy + dz * dz;
        if (d2 <= lim * lim) {
          if (n_cand == CAND_CAP) { dense = true; break; }
          cj[n_cand] = (uint8_t)j;
          ck[n_cand] = (uint8_t)k;
          cgap[n_cand] = sqrtf(d2) - d_far;
          ++n_cand;
        }
      }
    }
  }
  __device__ float reach(uint32_t s) const { return fmaf(w0[s], RdtW, fmaf((float)deg[s], accel, slack)); }
  __device__ bool linked(uint32_t j, uint32_t k) const { return gap_of(j, k) <= reach(j) + reach(k); }
  __device__ void solve() {
    for (int iter = 0; iter < 16; ++iter) {
      bool grew = false;
      // count links with the current reaches
      uint8_t cnt[MAX_ACTIVE];
      for (uint32_t s = 0; s < a; ++s) cnt[s] = 0;
      if (!dense) {
        for (uint32_t i = 0; i < n_cand; ++i)
          if (cgap[i] <= reach(cj[i]) + reach(ck[i])) { ++cnt[cj[i]]; ++cnt[ck[i]]; }
      } else {
        for (uint32_t j = 0; j + 1 < a; ++j) {
          const float rj = reach(j);
          for (uint32_t k = j + 1; k < a; ++k)
            if (gap_of(j, k) <= rj + reach(k)) { ++cnt[j]; ++cnt[k]; }
        }
      }
      for (uint32_t s = 0; s < a; ++s)
        if (cnt[s] > deg[s]) { deg[s] = cnt[s]; grew = true; }
      if (!grew) return;
    }
    for (uint32_t s = 0; s < a; ++s) deg[s] = (uint8_t)(a - 1);  // no fixed point in time: link everything
  }
  // the partners of slot s at the fixed point, ascending, at most deg[s] of them; returns how many
  __device__ uint32_t partners(uint32_t s, uint8_t *out) const {
    uint32_t n = 0;
    if (!dense) {
      // the candidates are in (j, k) order: those with k == s come first (ascending j < s), then those with j == s
      for (uint32_t i = 0; i < n_cand && n < deg[s]; ++i) {
        const uint32_t j = cj[i], k = ck[i];
        if ((j == s || k == s) && cgap[i] <= reach(j) + reach(k)) out[n++] = (uint8_t)(j == s ? k : j);
      }
    } else {
      for (uint32_t k = 0; k < a && n < deg[s]; ++k)
        if (k != s && linked(s < k ? s : k, s < k ? k : s)) out[n++] = (uint8_t)k;
    }
    return n;
  }
  __device__ bool force_free(uint32_t s) const {
    const float *f = base[act[s]].force;
    return f[0] == 0.0f && f[1] == 0.0f && f[2] == 0.0f;
  }
};

// Free-flight horizon of a whole trajectory (<= 8 moving particles): the number of steps for
// which NO pair can come within reach; forces must be exactly zero on entry (they are whenever
// the state came out of a step in which every pair was out of range).  Without forces the
// speeds are constant, so the bound is linear: (D - D_far) / (b_j + b_k).
__device__ inline uint32_t free_flight_horizon(const DevView &v, const TrajMeta &m) {
  const uint32_t a = m.n_active;
  if (a == 0 || a > (uint32_t)FREE_MAX_A) return 0;
  const heroam_particle *base = v.particles + m.first;
  const uint32_t *act = v.act + m.first;
  float px[FREE_MAX_A], py[FREE_MAX_A], pz[FREE_MAX_A], reach[FREE_MAX_A];
  const float R = v.c.radius, dt = v.c.dt;
  for (uint32_t s = 0; s < a; ++s) {
    const heroam_particle *p = base + act[s];
    if (p->force[0] != 0.0f || p->force[1] != 0.0f || p->force[2] != 0.0f) return 0;
    px[s] = p->position[0]; py[s] = p->position[1]; pz[s] = p->position[2];
    const float wx = p->ang_velocity[0], wy = p->ang_velocity[1], wz = p->ang_velocity[2];
    reach[s] = sqrtf(wx * wx + wy * wy + wz * wz) * R * dt * 1.01f + 2.0e-6f * R;  // per step
  }
  if (a == 1) return 0xffffffffu;
  const float d_far = sqrtf(v.c.far2) * 1.00001f + 1.0e-3f;
  float horizon = 4.0e9f;
  for (uint32_t j = 0; j + 1 < a; ++j)
    for (uint32_t k = j + 1; k < a; ++k) {
      const float dx = px[j] - px[k], dy = py[j] - py[k], dz = pz[j] - pz[k];
      const float gap = sqrtf(dx * dx + dy * dy + dz * dz) - d_far;
      if (!(gap > 0.0f)) return 0;
      const float h = gap / (reach[j] + reach[k]) * 0.999f - 2.0f;
      if (h < horizon) horizon = h;
    }
  return horizon < 1.0f ? 0u : (horizon > 4.0e9f ? 0xffffffffu : (uint32_t)horizon);
}

// Partner lists for the warp-per-trajectory kernels: slot s's list holds, in ascending order,
// every other moving particle linked to it for the next `window` steps (see Reach).
__device__ inline void build_partner_lists(const DevView &v, const TrajMeta &m, uint32_t window) {
  Reach rc;
  rc.init(v, m, m.n_active, window);
  rc.solve();
  const uint32_t a = m.n_active;
  unsigned long long total = 0;
  for (uint32_t s = 0; s < a; ++s) total += rc.deg[s];
  const unsigned long long at = atomicAdd(v.nbr_used, total);
  if (at + total > v.nbr_cap) {  // pool exhausted: this trajectory iterates over all partners
    for (uint32_t s = 0; s < a; ++s) v.nbr_cnt[m.first + s] = NBR_ALL;
    return;
  }
  unsigned long long cursor = at;
  for (uint32_t j = 0; j < a; ++j) {
    v.nbr_start[m.first + j] = (uint32_t)cursor;
    v.nbr_cnt[m.first + j] = rc.partners(j, v.nbr_pool + cursor);
    cursor += rc.deg[j];
  }
}

// What a resolve thread wants launched for its trajectory this pass.
struct Schedule {
  int bin = -1;            // bin of the trajectory (or of its core); -1: nothing for the integrate kernels
  int fly_tier = -1;       // free-flight list to append to; -1: nobody flies
  uint32_t fly_first = 0;  // first slot of the act list that flies
  uint32_t fly_count = 0;
};

// core-loner and loner-loner pair-steps (never evaluated: provably force-free) for the core
// steps taken since the last call
__device__ inline void account_cross_pairs(TrajMeta &m, Counters &work) {
  if (m.n_loners && m.steps > m.cross_from) {
    const unsigned long long l = m.n_loners, c = m.n_active;
    const unsigned long long n = (unsigned long long)(m.steps - m.cross_from) * (l * c + l * (l - 1) / 2);
    work.pair_steps += n;
    work.pair_skipped += n;
  }
  m.cross_from = m.steps;
}

// Commit a whole-trajectory free-flight segment (the kernel leaves TrajMeta alone): the
// trajectory is now at step_stop; its clock is the one the particles carry.
__device__ inline void land_free_flight(const DevView &v, TrajMeta &m, Counters &work) {
  const unsigned long long k = m.step_stop - m.steps;
  m.steps = m.step_stop;
  m.time = v.particles[m.first + v.act[m.first]].time;
  m.status = ST_LIVE;
  m.cross_from = m.steps;
  const unsigned long long a = m.n_active;
  work.particle_steps += k * a;
  work.pair_steps += k * (a * (a - 1) / 2);
  work.pair_skipped += k * (a * (a - 1) / 2);
  work.free_steps += k * a;
}

// The trajectory ends at step m.steps (success at a core meeting) while its loners are ahead
// at loner_until: the loop of the reference stops for everybody at that step, so put the
// loners where they were then -- re-fly them from their take-off orientation.  (With the
// `==` criterion at most one particle is still moving at that point.)
template <class P>
__device__ inline void rewind_loners(const DevView &v, TrajMeta &m, Counters &work) {
  const Consts &c = v.c;
  heroam_particle *base = v.particles + m.first;
  const uint32_t *act = v.act + m.first;
  for (uint32_t i = 0; i < m.n_loners; ++i) {
    const uint32_t j = act[m.core_slots + i];
    heroam_particle *p = base + j;
    const float4 q4 = v.ckpt_q[m.first + j];
    float q[4] = {q4.x, q4.y, q4.z, q4.w}, w[3], w_rec[3], r[3], vel[3], v2;
    ld3(p->ang_velocity, w_rec);
    load_w<P>(c, w_rec, w);
    float time = m.loner_time0;
    if (P::closed_flight) {
      rotate_free(w, q, m.steps - m.loner_from);
      for (uint32_t k = m.loner_from; k < m.steps; ++k) time = __fadd_rn(time, c.dt);
    } else {
      for (uint32_t k = m.loner_from; k < m.steps; ++k) {
        advance_orientation<P>(c, w, q, renorm_at<P>(k));
        time = __fadd_rn(time, c.dt);
      }
    }
    pole_position<P>(c, q, r);
    stored_velocity<P>(r, w_rec, vel, v2);
    p->orientation[0] = q[0]; p->orientation[1] = q[1]; p->orientation[2] = q[2]; p->orientation[3] = q[3];
    st3(p->position, r);
    st3(p->velocity, vel);
    p->velocity_sq = v2;
    p->time = time;
  }
  const unsigned long long undone = (unsigned long long)(m.loner_until - m.steps) * m.n_loners;
  work.particle_steps -= undone;  // unsigned wrap-around: the per-run sums come out right
  work.free_steps -= undone;
  m.loner_until = m.steps;
}

// Everything a resolve thread does for a live trajectory before scheduling: finish a pending
// meeting step, land a whole-trajectory flight, account work, let an emptied core jump to its
// loners.
template <class P>
__device__ inline void settle(const DevView &v, TrajMeta &m, Counters &work) {
  if (m.status == ST_MIDSTEP) {
    account_cross_pairs(m, work);   // steps before the meeting, old core size
    finish_step<P>(v, m, work);
    account_cross_pairs(m, work);   // the meeting step itself, surviving core
  } else if (m.status == ST_FLYING) {
    land_free_flight(v, m, work);
  } else {
    account_cross_pairs(m, work);
  }
}

// List the moving particles, decide who flies and who is integrated, choose the bin, and fix
// the step at which the segment ends.
__device__ inline Schedule schedule_segment(const DevView &v, TrajMeta &m, const Segments &sg, uint32_t step_cap,
                                            Counters &work) {
  Schedule out;
  heroam_particle *base = v.particles + m.first;
  uint32_t *act = v.act + m.first;
  // ---- loners ahead of the core: the core catches up on its own -------------------------------
  if (m.n_loners && m.steps < m.loner_until) {
    uint32_t lon[MAX_ACTIVE];
    for (uint32_t i = 0; i < m.n_loners; ++i) lon[i] = act[m.core_slots + i];
    uint32_t s = 0;
    for (uint32_t j = 0; j < m.no; ++j) {
      if (base[j].success) continue;
      bool is_loner = false;
      for (uint32_t i = 0; i < m.n_loners; ++i) is_loner = is_loner || lon[i] == j;
      if (!is_loner) act[s++] = j;
    }
    m.core_slots = s;  // == m.n_active
    for (uint32_t i = 0; i < m.n_loners; ++i) act[s + i] = lon[i];
    if (m.n_active > 0) {
      uint32_t stop = m.steps + sg.of_bin[BIN_CORE + (int)m.n_active];
      if (stop >= m.steps && stop - m.steps >= 2u * RENORM_ALIGN) stop &= ~(RENORM_ALIGN - 1u);
      if (stop < m.steps || stop > m.loner_until) stop = m.loner_until;
      m.step_stop = stop;
      out.bin = BIN_CORE + (int)m.n_active;
      return out;
    }
    // nobody left in the core: its remaining steps up to the loners are empty steps
    account_cross_pairs(m, work);
    const unsigned long long l = m.n_loners;
    const unsigned long long n = (unsigned long long)(m.loner_until - m.steps) * (l * (l - 1) / 2);
    work.pair_steps += n;
    work.pair_skipped += n;
    m.steps = m.loner_until;
    m.cross_from = m.steps;
    m.time = base[lon[0]].time;
  }
  // ---- everybody at the same step: fresh analysis -----------------------------------------------
  m.n_active += m.n_loners;
  m.n_loners = 0;
  uint32_t a = 0;
  for (uint32_t j = 0; j < m.no; ++j)
    if (!base[j].success) act[a++] = j;
  m.core_slots = a;
  const bool one_step_only = m.steps == 0 && all_paired(m.no, m.n_flagged);  // Q2
  auto clip = [&](uint32_t stop) {
    if (stop < m.steps) stop = 0xffffffffu;
    // segments end on renormalisation boundaries, so that the trajectories of a warp stay in phase
    if (stop - m.steps >= 2u * RENORM_ALIGN) stop &= ~(RENORM_ALIGN - 1u);
    if (one_step_only) stop = 1;
    if (stop > v.c.max_steps) stop = v.c.max_steps;
    if (stop > step_cap) stop = step_cap;
    return stop;
  };
  if (sg.no_flight) {
    out.bin = bin_of(a);
    m.step_stop = clip(m.steps + sg.of_bin[out.bin]);
    if (out.bin >= BIN_WARP1 && out.bin <= BIN_WARP4)  // no partner lists either: every pair is evaluated
      for (uint32_t s = 0; s < a; ++s) v.nbr_cnt[m.first + s] = NBR_ALL;
    return out;
  }
  // (1) the whole trajectory in free flight: the longest fixed-length tier its horizon covers
  const uint32_t horizon = free_flight_horizon(v, m);
  for (int tier = 2; tier >= (sg.drain ? 1 : 0); --tier)
    if (horizon >= sg.of_bin[BIN_FREE + tier]) {
      m.step_stop = clip(m.steps + sg.of_bin[BIN_FREE + tier]);
      m.status = ST_FLYING;
      out.fly_tier = tier;
      out.fly_first = 0;
      out.fly_count = a;
      return out;
    }
  // (2) split: particles linked to nobody for the window fly alone, the rest is the core
  // While draining, a narrow trajectory advances further per pass in the kernel of its size
  // (longer segments) than split with the bounded window; wide ones always gain from splitting.
  const uint32_t window = sg.of_bin[BIN_FREE + FREE_SPLIT];
  if (a >= 2 && a <= (uint32_t)MAX_ACTIVE && window >= 64u && a > sg.nosplit_max) {
    Reach rc;
    rc.init(v, m, a, window);
    rc.solve();
    uint32_t n_core = 0, n_lon = 0;
    uint32_t core[MAX_ACTIVE], lon[MAX_ACTIVE];
    for (uint32_t s = 0; s < a; ++s) {
      if (rc.deg[s] == 0 && rc.force_free(s)) lon[n_lon++] = act[s];
      else core[n_core++] = act[s];
    }
    if (n_lon >= 1 && n_core <= (uint32_t)THREAD_MAX_A) {
      for (uint32_t s = 0; s < n_core; ++s) act[s] = core[s];
      for (uint32_t i = 0; i < n_lon; ++i) {
        act[n_core + i] = lon[i];
        const float *o = base[lon[i]].orientation;
        v.ckpt_q[m.first + lon[i]] = make_float4(o[0], o[1], o[2], o[3]);
      }
      const uint32_t stop = clip(m.steps + window);
      m.n_active = n_core;
      m.core_slots = n_core;
      m.n_loners = n_lon;
      m.loner_from = m.steps;
      m.loner_until = stop;
      m.loner_time0 = m.time;
      m.cross_from = m.steps;
      m.step_stop = stop;
      const unsigned long long flown = (unsigned long long)(stop - m.steps) * n_lon;
      work.particle_steps += flown;
      work.free_steps += flown;
      out.bin = n_core ? BIN_CORE + (int)n_core : -1;
      out.fly_tier = FREE_SPLIT;
      out.fly_first = n_core;
      out.fly_count = n_lon;
      return out;
    }
  }
  // (3) the whole trajectory in the kernel of its moving count
  out.bin = bin_of(a);
  m.step_stop = clip(m.steps + sg.of_bin[out.bin]);
  if (out.bin >= BIN_WARP1 && out.bin <= BIN_WARP4 && m.step_stop > m.steps) build_partner_lists(v, m, m.step_stop - m.steps);
  return out;
}

// Append what a resolve thread scheduled.  Integrate bins: warp-aggregated, one atomic per
// (warp, bin); must be called by all 32 lanes.  Flyers: one entry per PARTICLE, (trajectory,
// slot in its act list).
// Where a resolve thread puts what it scheduled.  An ordinary run is one part: its counters are the pass's
// counters and its lists start at the beginning of every bin's list.  A partitioned pool (a droplet-size sweep as
// one run: every size has its own sphere radius, hence its own constants and its own scaled table) gives each
// part its own bin counters and its own stretch of every list, so that each (part, bin) is launched with the
// part's constants in the constant bank.
struct PartSlice {
  uint32_t *bin_count;     // this part's counters (N_BIN_COUNTERS of them)
  uint32_t *global_count;  // the counters of part 0: LIVE_SLOT (cursor of the live list) and LAG_SLOT are per pass
  uint32_t part;           // key for the warp-level aggregation: lanes of different parts never share an atomic
  uint32_t list_offset;    // start of this part inside every bin's list
  uint32_t free_offset;    // start of this part inside every free-flight tier's list
};

__device__ inline void append_schedule(const Schedule &sc, uint32_t t, uint32_t stride, const PartSlice &ps,
                                       uint32_t *__restrict__ bin_list, uint2 *__restrict__ free_list,
                                       uint32_t free_stride, uint32_t *__restrict__ live_out, uint32_t steps_done) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned peers = __match_any_sync(0xffffffffu, sc.bin >= 0 ? (int)(ps.part * 64u) + sc.bin : -1);
  {  // live trajectories (anything scheduled): counted and listed for the next pass, one atomic per warp
    const bool is_live = sc.bin >= 0 || sc.fly_tier >= 0;
    const unsigned alive = __ballot_sync(0xffffffffu, is_live);
    if (alive) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(ps.global_count + LIVE_SLOT, (uint32_t)__popc(alive));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (is_live) live_out[base + __popc(alive & ((1u << lane) - 1u))] = t;
      // the trajectory furthest behind: the longest chain of steps still to come bounds the end of the run
      const uint32_t lag = __reduce_max_sync(0xffffffffu, is_live ? 0xffffffffu - steps_done : 0u);
      if (lane == 0) atomicMax(ps.global_count + LAG_SLOT, lag);
      // live trajectories per part (its free slots staying free tells the host that the part's index space is used up)
      const unsigned same_part = __match_any_sync(0xffffffffu, is_live ? (int)ps.part : -1);
      if (is_live && (int)lane == __ffs(same_part) - 1) atomicAdd(ps.bin_count + PART_LIVE_SLOT, (uint32_t)__popc(same_part));
    }
  }
  if (sc.bin >= 0) {
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(ps.bin_count + sc.bin, (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    bin_list[(size_t)sc.bin * stride + ps.list_offset + base + rank] = t;
  }
  if (sc.fly_tier >= 0 && sc.fly_count) {
    const uint32_t base = atomicAdd(ps.bin_count + BIN_FREE + sc.fly_tier, sc.fly_count);
    uint2 *list = free_list + (size_t)sc.fly_tier * free_stride + ps.free_offset;
    for (uint32_t s = 0; s < sc.fly_count; ++s) list[base + s] = make_uint2(t, sc.fly_first + s);
  }
}

__device__ inline void flush_work(const DevView &v, const Counters &work) {
  const unsigned long long ps = warp_sum_u64(work.particle_steps);
  const unsigned long long pr = warp_sum_u64(work.pair_steps);
  const unsigned long long ir = warp_sum_u64(work.inrange_steps);
  const unsigned long long un = warp_sum_u64(work.unsupported);
  const unsigned long long sk = warp_sum_u64(work.pair_skipped);
  const unsigned long long fs = warp_sum_u64(work.free_steps);
  if ((threadIdx.x & 31u) == 0 && (ps | pr | ir | un | fs)) {
    atomicAdd(&v.counters->particle_steps, ps);
    atomicAdd(&v.counters->pair_steps, pr);
    atomicAdd(&v.counters->inrange_steps, ir);
    if (un) atomicAdd(&v.counters->unsupported, un);
    if (sk) atomicAdd(&v.counters->pair_skipped, sk);
    if (fs) atomicAdd(&v.counters->free_steps, fs);
  }
}

// First half of a resolve visit of a live trajectory: finish what the kernels left pending and
// evaluate the exit.  Returns true if the trajectory is over (m.status = ST_DONE + code).
// P: arithmetic of the meeting step; H: arithmetic of the integrate kernels (a rewound loner
// must be re-flown exactly as k_free_flight<H> flew it).
template <class P, class H>
__device__ inline bool settle_and_exit(const DevView &v, TrajMeta &m, uint32_t step_cap, Counters &work) {
  settle<P>(v, m, work);
  const uint32_t done = evaluate_exit(v.c, m, step_cap);
  if (done == NOT_DONE) return false;
  if (m.n_loners && m.loner_until > m.steps) rewind_loners<H>(v, m, work);
  m.status = ST_DONE + done;
  if (done == HEROAM_DONE_UNSUPPORTED) work.unsupported += 1;
  return true;
}

// One resolve visit of a live trajectory: returns what to launch for it (nothing if it ended).
template <class P, class H>
__device__ inline Schedule resolve_live(const DevView &v, TrajMeta &m, const Segments &sg, uint32_t step_cap,
                                        Counters &work) {
  if (settle_and_exit<P, H>(v, m, step_cap, work)) return Schedule();
  return schedule_segment(v, m, sg, step_cap, work);
}

// ---------------------------------------------------------------------------------
// k_resolve_and_bin: one thread per trajectory, once per pass (fixed batch).
// ---------------------------------------------------------------------------------
// The trajectories to visit: all of them (live_in == nullptr, the first pass) or the ones the
// previous pass listed as still live -- late in a run they are a small fraction of the batch.
struct LiveLists {
  const uint32_t *in;   // nullptr: visit trajectory t = thread index, t < count
  uint32_t count;
  uint32_t *out;        // the trajectories still live after this pass, in no particular order
};

template <class P, class H>
__global__ void k_resolve_and_bin(DevView v, Segments sg, uint32_t step_cap, uint32_t *__restrict__ bin_count,
                                  uint32_t *__restrict__ bin_list, uint2 *__restrict__ free_list, LiveLists live) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t t = idx;
  Counters work = {};
  Schedule sc;
  uint32_t steps_done = 0;
  if (idx < live.count) {
    if (live.in) t = live.in[idx];
    TrajMeta m = v.meta[t];
    if (m.status < ST_DONE) {
      sc = resolve_live<P, H>(v, m, sg, step_cap, work);
      v.meta[t] = m;
      steps_done = m.steps;
    }
  }
  const PartSlice whole = {bin_count, bin_count, 0u, 0u, 0u};
  append_schedule(sc, t, v.n_traj, whole, bin_list, free_list, sg.free_stride, live.out, steps_done);
  flush_work(v, work);
}

// ---------------------------------------------------------------------------------
// k_free_flight: one thread per flying PARTICLE.  Zero force: the half kicks add +-0, w stays
// what it is, and a step is nothing but the orientation update; the position (only needed by
// the pair loop) is rebuilt once at the end.  Per step this performs exactly the operations
// the full kernels perform on that particle, so the records it leaves are bit-identical.
// Flyers are either all moving particles of a trajectory (they fly [steps, step_stop)) or the
// loners of a split trajectory (act slots >= n_active; they fly [loner_from, loner_until)).
// TrajMeta is only read, and only fields no concurrent core kernel changes.
// ---------------------------------------------------------------------------------
struct Flyer {
  heroam_particle *p;
  float q[4], w[3], w_rec[3], time;
  uint32_t from, until;
};

template <class P>
__device__ __forceinline__ void load_flyer(const DevView &v, const TrajMeta &m, uint32_t slot, Flyer &fl) {
  const uint32_t first = m.first;
  const bool loner = m.n_loners != 0u;  // whole-trajectory flights have none
  fl.from = loner ? m.loner_from : m.steps;
  fl.until = loner ? m.loner_until : m.step_stop;
  fl.time = loner ? m.loner_time0 : m.time;
  fl.p = v.particles + first + v.act[first + slot];
  fl.q[0] = fl.p->orientation[0]; fl.q[1] = fl.p->orientation[1];
  fl.q[2] = fl.p->orientation[2]; fl.q[3] = fl.p->orientation[3];
  ld3(fl.p->ang_velocity, fl.w_rec);
  load_w<P>(v.c, fl.w_rec, fl.w);
}

// N flyers of ONE trajectory (same from / until / clock) advanced together: their steps are
// independent, so interleaving them hides each other's dependency chains.  Per flyer exactly
// the operations of the full kernels, in their order.  CLOCK_PRODUCT (Consts::clock_exact): the
// clock is set at the end, (float)until * dt, instead of being advanced in every step.
template <class P, int N, bool CLOCK_PRODUCT>
__device__ __forceinline__ void fly_together(const Consts &c, Flyer (&fl)[N]) {
  uint32_t k = fl[0].from;
  const uint32_t until = fl[0].until;
  float time = fl[0].time;
  if (P::closed_flight) {  // the whole flight as one rotation; only the clock still counts the steps
#pragma unroll
    for (int i = 0; i < N; ++i) rotate_free(fl[i].w, fl[i].q, until - k);
    if (!CLOCK_PRODUCT)
      for (; k < until; ++k) time = __fadd_rn(time, c.dt);
    k = until;
  }
  if (P::renorm_every > 1u) {
    // renormalisation after every renorm_every-th absolute step: single steps up to the next
    // boundary, then whole blocks without any test, then the remainder
    for (; k < until && (k & (P::renorm_every - 1u)) != 0u; ++k) {
#pragma unroll
      for (int i = 0; i < N; ++i) advance_orientation<P>(c, fl[i].w, fl[i].q, renorm_at<P>(k));
      if (!CLOCK_PRODUCT) time = __fadd_rn(time, c.dt);
    }
    for (; k + P::renorm_every <= until; k += P::renorm_every) {
#pragma unroll
      for (uint32_t u = 0; u < P::renorm_every; ++u) {
#pragma unroll
        for (int i = 0; i < N; ++i) advance_orientation<P>(c, fl[i].w, fl[i].q, u + 1u == P::renorm_every);
        if (!CLOCK_PRODUCT) time = __fadd_rn(time, c.dt);
      }
    }
  }
  for (; k < until; ++k) {
#pragma unroll
    for (int i = 0; i < N; ++i) advance_orientation<P>(c, fl[i].w, fl[i].q, renorm_at<P>(k));
    if (!CLOCK_PRODUCT) time = __fadd_rn(time, c.dt);
  }
  if (CLOCK_PRODUCT && until > fl[0].from) time = __fmul_rn(__uint2float_rn(until), c.dt);
#pragma unroll
  for (int i = 0; i < N; ++i) fl[i].time = time;
}

template <class P, bool CLOCK_PRODUCT>
__device__ __forceinline__ void fly_alone(const Consts &c, Flyer &fl) {
  Flyer one[1] = {fl};
  fly_together<P, 1, CLOCK_PRODUCT>(c, one);
  fl = one[0];
}

template <class P>
__device__ __forceinline__ void land_flyer(const Consts &c, const Flyer &fl) {
  float r[3], vel[3], v2;
  pole_position<P>(c, fl.q, r);
  stored_velocity<P>(r, fl.w_rec, vel, v2);
  heroam_particle *p = fl.p;
  p->orientation[0] = fl.q[0]; p->orientation[1] = fl.q[1]; p->orientation[2] = fl.q[2]; p->orientation[3] = fl.q[3];
  st3(p->position, r);
  st3(p->velocity, vel);
  p->velocity_sq = v2;
  p->time = fl.time;
}

template <class P, bool CLOCK_PRODUCT>
__global__ void __launch_bounds__(128) k_free_flight(DevView v, const uint2 *__restrict__ list, uint32_t count) {
  const uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= count) return;
  Flyer fl;
  const uint2 e = list[item];
  load_flyer<P>(v, v.meta[e.x], e.y, fl);
  fly_alone<P, CLOCK_PRODUCT>(v.c, fl);
  land_flyer<P>(v.c, fl);
}

// ---------------------------------------------------------------------------------
// Turbo keeps the forces scaled by kappa, which is not a power of two: a force that went out
// to the record (F = F' / kappa, rounded) and came back (F * kappa, rounded) need not be the
// F' the kernel held, and the state would then depend on where the segment boundaries fall.
// The force is a function of the positions, which do round-trip exactly, so a Turbo kernel
// rebuilds it on entry with the very operations of its step.  (The angular velocity is scaled
// by dt/2: exact for the usual power-of-two time steps.)
// ---------------------------------------------------------------------------------
template <int A, class P>
__device__ __forceinline__ void rebuild_forces(const DevView &v, const float (&r)[A][3], float (&f)[A][3]) {
#pragma unroll
  for (int s = 0; s < A; ++s) { f[s][0] = 0.0f; f[s][1] = 0.0f; f[s][2] = 0.0f; }
#pragma unroll
  for (int j = 0; j < A - 1; ++j) {
#pragma unroll
    for (int k = j + 1; k < A; ++k) {
      const PairGeom g = pair_geom<P>(v.c, r[j], r[k]);
      pair_apply<P>(pair_scale<P>(pair_gather<P>(v, g), g.d), g, f[j], f[k]);
    }
  }
}

// ---------------------------------------------------------------------------------
// k_integrate_thread<A>: one thread per trajectory, A active particles in registers.
//
// LOOKAHEAD (launched only while a run drains, A <= 4): when few warps are resident nothing
// hides the L2 round trip of the table gather, which sits in the middle of the step's
// dependency chain (a pair's table position moves ~18 bins = 72 bytes per step, so every step
// touches a new 32-byte sector and L1 does not help on its own; ncu: 59 % of a 650-cycle step).
// A pair's table position moves smoothly (second difference ~1e-3 bin), so the position TWO
// steps ahead is extrapolated and its sector pulled into L1 with a cp.async whose destination
// nobody reads: no register and so no scoreboard slot is held, nothing has to be selected or
// checked afterwards, and a wrong guess only means an ordinary miss.  A run of 1e5 trajectories
// (all of it drain) takes 4.2 s with it, 4.65 s without, 5.1 s with the one-step look-ahead
// into registers it replaced (DESIGN.md section 7 has the other variants).
// ---------------------------------------------------------------------------------
// One segment of a trajectory with A moving particles, by one thread: loads the records, runs
// steps [m.steps, m.step_stop) or up to a meeting, leaves the records (clean, or mid-step) and
// updates m (steps, clock, status).  Returns the number of steps completed.
template <int A, class P, bool LOOKAHEAD>
__device__ __forceinline__ uint32_t integrate_thread_segment(const DevView &v, TrajMeta &m, const unsigned sink_addr,
                                                             unsigned long long &inrange_total) {
  const Consts &c = v.c;
  heroam_particle *rec[A];
  float q[A][4], w[A][3], r[A][3], f[A][3], hk[A][3], wh[A][3], r_old[A][3];
#pragma unroll
  for (int s = 0; s < A; ++s) {
    rec[s] = v.particles + m.first + v.act[m.first + s];
    const heroam_particle *p = rec[s];
    q[s][0] = p->orientation[0]; q[s][1] = p->orientation[1];
    q[s][2] = p->orientation[2]; q[s][3] = p->orientation[3];
    ld3(p->position, r[s]);
    load_w<P>(c, p->ang_velocity, w[s]);
    load_f<P>(c, p->force, f[s]);
  }
  if (P::turbo) rebuild_forces<A, P>(v, r, f);
#pragma unroll
  for (int s = 0; s < A; ++s) half_kick<P>(c, r[s], f[s], hk[s]);
  uint32_t steps = m.steps;
  float time = m.time;
  bool met = false;
  constexpr int NP = A * (A - 1) / 2;
  float pos_prev[LOOKAHEAD ? NP + 1 : 1];
#pragma unroll
  for (int i = 0; i < (LOOKAHEAD ? NP : 0); ++i) pos_prev[i] = 0.0f;
  while (steps < m.step_stop) {
    const bool rn = renorm_at<P>(steps);
    // first half kick + drift (simulation.c:478-490)
#pragma unroll
    for (int s = 0; s < A; ++s) {
      r_old[s][0] = r[s][0]; r_old[s][1] = r[s][1]; r_old[s][2] = r[s][2];
      wh[s][0] = P::add(w[s][0], hk[s][0]);
      wh[s][1] = P::add(w[s][1], hk[s][1]);
      wh[s][2] = P::add(w[s][2], hk[s][2]);
      advance_orientation<P>(c, wh[s], q[s], P::renorm_every == 1u);
    }
    if (P::renorm_every > 1u && rn) {  // one branch per step for all particles; uniform when the warp is in phase
#pragma unroll
      for (int s = 0; s < A; ++s) renormalise(q[s]);
    }
#pragma unroll
    for (int s = 0; s < A; ++s) {
      pole_position<P>(c, q[s], r[s]);
      f[s][0] = 0.0f; f[s][1] = 0.0f; f[s][2] = 0.0f;
    }
    // all pairs in the reference's (j,k) order (:178-233)
    // Branch-free and staged row by row: geometry and gathers of a whole row are
    // issued before any of them is consumed.
    float dmin = 3.0e38f;
    uint32_t in_step = 0;
#pragma unroll
    for (int j = 0; j < A - 1; ++j) {
      PairGeom g[A];
      float tv[A];
#pragma unroll
      for (int k = j + 1; k < A; ++k) {
        g[k] = pair_geom<P>(c, r[j], r[k]);
        tv[k] = pair_gather<P>(v, g[k]);
        if (LOOKAHEAD) {
          // L1 prefetch by cp.async: the sector around the table position predicted for two steps
          // from now is copied to a dummy slot of shared memory that nobody reads; the copy passes
          // through L1, so the ordinary gather of that step finds its sector there
          const int pi = j * A - j * (j + 1) / 2 + (k - j - 1);  // index of pair (j,k), static after unrolling
          const float pos = P::turbo ? fmaf(g[k].d, c.inv_grid, c.pos_bias) : __fmul_rn(__fsub_rn(g[k].d, c.grid_start), c.inv_grid);
          const float guess = fmaf(3.0f, pos, -2.0f * pos_prev[pi]);
          pos_prev[pi] = pos;
          const unsigned at = min((unsigned)__float2int_rz(guess), (unsigned)c.grid_len);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sink_addr), "l"((P::turbo ? v.table_k : v.table) + at));
        }
        dmin = fminf(dmin, g[k].d);
        if (P::count) in_step += g[k].inside ? 1u : 0u;
      }
#pragma unroll
      for (int k = j + 1; k < A; ++k) pair_apply<P>(pair_scale<P>(tv[k], g[k].d), g[k], f[j], f[k]);
    }
    if (LOOKAHEAD) {
      asm volatile("cp.async.commit_group;");
      asm volatile("cp.async.wait_group 3;");
    }
    if (dmin < c.icd2) { met = true; break; }
    // second half kick (:496-509); velocity and clock fields are written at the end
#pragma unroll
    for (int s = 0; s < A; ++s) {
      half_kick<P>(c, r[s], f[s], hk[s]);
      w[s][0] = P::add(wh[s][0], hk[s][0]);
      w[s][1] = P::add(wh[s][1], hk[s][1]);
      w[s][2] = P::add(wh[s][2], hk[s][2]);
    }
    time = __fadd_rn(time, c.dt);
    ++steps;
    inrange_total += in_step;
  }
  if (LOOKAHEAD) asm volatile("cp.async.wait_all;");
  const uint32_t done_steps = steps - m.steps;
  if (met) {
#pragma unroll
    for (int s = 0; s < A; ++s) write_midstep<P>(c, rec[s], q[s], r[s], w[s], wh[s], r_old[s], time, steps > 0);
    m.status = ST_MIDSTEP;
  } else if (done_steps > 0) {
#pragma unroll
    for (int s = 0; s < A; ++s) {
      const uint32_t j = (uint32_t)(rec[s] - (v.particles + m.first));
      write_clean<P>(c, rec[s], q[s], r[s], w[s], f[s], time, j + 1u < m.no);
    }
  }
  m.steps = steps;
  m.time = time;
  return done_steps;
}

// Registers: ptxas left to itself takes 80 for A = 2; capped at 72 (seven blocks per SM instead of six) the
// two-particle kernel is 15 % faster at full occupancy (ncu, round 2: 4.35 against 5.14 ms for 3.8 k blocks).
// The wider ones lose from a cap: with fewer registers fewer gathers are in flight (A = 3, 4, 6 at
// 96 / 128 / 168 registers instead of 108 / 144 / 228: 6-60 % slower, long_scoreboard 2.3-2.9 cycles per
// instruction instead of 0.4-1.3; profiles/r2_ncu_summary.md sections 3 and 3b).
// The body lives in integrate_thread_segment rather than in the kernel itself: same arithmetic, but ptxas
// then carries r_old without half of the register copies (loop of A = 2: 128 instead of 135 instructions
// and no spill under the 72-register cap; A = 6: 595 instead of 613; A = 8: 950 instead of 974).
template <int A, class P, bool LOOKAHEAD = false>
__global__ void __launch_bounds__(INTEGRATE_TPB, (A == 2 && !LOOKAHEAD) ? 7 : 1)
k_integrate_thread(DevView v, const uint32_t *__restrict__ list, uint32_t count) {
  __shared__ float prefetch_sink[LOOKAHEAD ? INTEGRATE_TPB : 1];
  const unsigned sink_addr = (unsigned)__cvta_generic_to_shared(prefetch_sink + (LOOKAHEAD ? threadIdx.x : 0));
  const uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long done_steps = 0, inrange_total = 0;
  if (item < count) {
    const uint32_t t = list[item];
    TrajMeta m = v.meta[t];
    done_steps = integrate_thread_segment<A, P, LOOKAHEAD>(v, m, sink_addr, inrange_total);
    v.meta[t] = m;
  }
  const unsigned long long ds = warp_sum_u64(done_steps);
  const unsigned long long ir = warp_sum_u64(inrange_total);
  if ((threadIdx.x & 31u) == 0 && ds) {
    atomicAdd(&v.counters->particle_steps, ds * A);
    atomicAdd(&v.counters->pair_steps, ds * (A * (A - 1) / 2));
    atomicAdd(&v.counters->inrange_steps, ir);
  }
}

// ---------------------------------------------------------------------------------
// k_integrate_hybrid<A>, A = 9..12: one thread per trajectory.  Positions and forces (the
// operands of the A(A-1)/2-pair loop, which dominates at this size) stay in registers and
// the pair loop is fully unrolled exactly as in k_integrate_thread; what is touched only
// once per particle and step -- orientation, angular velocity, its half-step value and the
// pre-drift position kept for a possible meeting -- lives in the thread's private column
// of shared memory, [array][particle][thread]:
//     qs = orientation     ws = (w.xyz, r_old.x)     hs = (w(t+dt/2).xyz, r_old.y)     zs = r_old.z
// Same operations in the same order as the register kernels: bit-identical results.
// ---------------------------------------------------------------------------------
template <int A>
constexpr size_t hybrid_smem_bytes() { return (size_t)A * SMEM_TPB * (3 * sizeof(float4) + sizeof(float)); }

template <int A, class P>
__global__ void __launch_bounds__(SMEM_TPB) k_integrate_hybrid(DevView v, const uint32_t *__restrict__ list, uint32_t count) {
  extern __shared__ float4 sm4[];
  const uint32_t tid = threadIdx.x;
  float4 *qs = sm4 + tid, *ws = qs + A * SMEM_TPB, *hs = ws + A * SMEM_TPB;
  float *zs = reinterpret_cast<float *>(sm4 + 3 * A * SMEM_TPB) + tid;
  const uint32_t item = blockIdx.x * SMEM_TPB + tid;
  unsigned long long done_steps = 0, inrange_total = 0;
  if (item < count) {
    const Consts &c = v.c;
    const uint32_t t = list[item];
    TrajMeta m = v.meta[t];
    heroam_particle *base = v.particles + m.first;
    const uint32_t *act = v.act + m.first;
    float r[A][3], f[A][3];
#pragma unroll
    for (int s = 0; s < A; ++s) {
      const heroam_particle *p = base + act[s];
      qs[s * SMEM_TPB] = make_float4(p->orientation[0], p->orientation[1], p->orientation[2], p->orientation[3]);
      float w0[3];
      load_w<P>(c, p->ang_velocity, w0);
      ws[s * SMEM_TPB] = make_float4(w0[0], w0[1], w0[2], 0.0f);
      ld3(p->position, r[s]);
      load_f<P>(c, p->force, f[s]);
    }
    if (P::turbo) rebuild_forces<A, P>(v, r, f);
    uint32_t steps = m.steps;
    float time = m.time;
    bool met = false;
    while (steps < m.step_stop) {
      const bool rn = renorm_at<P>(steps);
#pragma unroll
      for (int s = 0; s < A; ++s) {
        const float4 w4 = ws[s * SMEM_TPB], q4 = qs[s * SMEM_TPB];
        float hk[3], wh[3], q[4] = {q4.x, q4.y, q4.z, q4.w};
        half_kick<P>(c, r[s], f[s], hk);
        wh[0] = P::add(w4.x, hk[0]); wh[1] = P::add(w4.y, hk[1]); wh[2] = P::add(w4.z, hk[2]);
        ws[s * SMEM_TPB] = make_float4(w4.x, w4.y, w4.z, r[s][0]);
        hs[s * SMEM_TPB] = make_float4(wh[0], wh[1], wh[2], r[s][1]);
        zs[s * SMEM_TPB] = r[s][2];
        drift<P>(c, wh, q, r[s], rn);
        qs[s * SMEM_TPB] = make_float4(q[0], q[1], q[2], q[3]);
        f[s][0] = 0.0f; f[s][1] = 0.0f; f[s][2] = 0.0f;
      }
      float dmin = 3.0e38f;
      uint32_t in_step = 0;
#pragma unroll
      for (int j = 0; j < A - 1; ++j) {
        PairGeom g[A];
        float tv[A];
#pragma unroll
        for (int k = j + 1; k < A; ++k) {
          g[k] = pair_geom<P>(c, r[j], r[k]);
          tv[k] = pair_gather<P>(v, g[k]);
          dmin = fminf(dmin, g[k].d);
          if (P::count) in_step += g[k].inside ? 1u : 0u;
        }
#pragma unroll
        for (int k = j + 1; k < A; ++k) pair_apply<P>(pair_scale<P>(tv[k], g[k].d), g[k], f[j], f[k]);
      }
      if (dmin < c.icd2) { met = true; break; }
#pragma unroll
      for (int s = 0; s < A; ++s) {
        const float4 h4 = hs[s * SMEM_TPB];
        float hk[3];
        half_kick<P>(c, r[s], f[s], hk);
        ws[s * SMEM_TPB] = make_float4(P::add(h4.x, hk[0]), P::add(h4.y, hk[1]), P::add(h4.z, hk[2]), 0.0f);
      }
      time = __fadd_rn(time, c.dt);
      ++steps;
      inrange_total += in_step;
    }
    done_steps = steps - m.steps;
#pragma unroll
    for (int s = 0; s < A; ++s) {
      heroam_particle *p = base + act[s];
      const float4 q4 = qs[s * SMEM_TPB], w4 = ws[s * SMEM_TPB];
      const float q[4] = {q4.x, q4.y, q4.z, q4.w}, w[3] = {w4.x, w4.y, w4.z};
      if (met) {
        const float4 h4 = hs[s * SMEM_TPB];
        const float wh[3] = {h4.x, h4.y, h4.z}, r_old[3] = {w4.w, h4.w, zs[s * SMEM_TPB]};
        write_midstep<P>(c, p, q, r[s], w, wh, r_old, time, steps > 0);
      } else if (done_steps > 0) {
        write_clean<P>(c, p, q, r[s], w, f[s], time, act[s] + 1u < m.no);
      }
    }
    if (met) m.status = ST_MIDSTEP;
    m.steps = steps;
    m.time = time;
    v.meta[t] = m;
  }
  if (done_steps) {
    atomicAdd(&v.counters->particle_steps, done_steps * A);
    atomicAdd(&v.counters->pair_steps, done_steps * (A * (A - 1) / 2));
    atomicAdd(&v.counters->inrange_steps, inrange_total);
  }
}

// ---------------------------------------------------------------------------------
// k_integrate_smem<CAP>: one thread per trajectory with 9..CAP moving particles.  Too many
// for registers, so the per-particle state lives in the thread's private column of shared
// memory, five float4 per particle laid out [array][particle][thread] (consecutive threads
// -> consecutive 16-byte words: conflict-free 128-bit accesses):
//     pos = (r.xyz, r_old.x)   frc = (F.xyz, r_old.y)   w = (w.xyz, r_old.z)
//     q   = orientation        wh  = (w(t + dt/2), -)
// Loops are rolled (any 9 <= a <= CAP), the pair loop is the reference's (j,k) order and the
// accumulators start from the value already in memory, so results are bit-identical to the
// register kernels.  Compared with the warp-per-trajectory kernel every pair is evaluated
// once and no lane idles.
// ---------------------------------------------------------------------------------
template <int CAP, class P>
__global__ void __launch_bounds__(SMEM_TPB) k_integrate_smem(DevView v, const uint32_t *__restrict__ list, uint32_t count) {
  extern __shared__ float4 sm4[];
  const uint32_t tid = threadIdx.x;
  float4 *pos = sm4 + tid, *frc = pos + CAP * SMEM_TPB, *qq = frc + CAP * SMEM_TPB, *ww = qq + CAP * SMEM_TPB,
         *whh = ww + CAP * SMEM_TPB;
#define COL(arr, s) (arr)[(s) * SMEM_TPB]
  const uint32_t item = blockIdx.x * SMEM_TPB + tid;
  unsigned long long done_steps = 0, inrange_total = 0;
  uint32_t a = 0;
  if (item < count) {
    const Consts &c = v.c;
    const uint32_t t = list[item];
    TrajMeta m = v.meta[t];
    a = m.n_active;
    heroam_particle *base = v.particles + m.first;
    const uint32_t *act = v.act + m.first;
    for (uint32_t s = 0; s < a; ++s) {
      const heroam_particle *p = base + act[s];
      COL(qq, s) = make_float4(p->orientation[0], p->orientation[1], p->orientation[2], p->orientation[3]);
      COL(pos, s) = make_float4(p->position[0], p->position[1], p->position[2], 0.0f);
      float w0[3], f0[3];
      load_w<P>(c, p->ang_velocity, w0);
      load_f<P>(c, p->force, f0);
      COL(ww, s) = make_float4(w0[0], w0[1], w0[2], 0.0f);
      COL(frc, s) = make_float4(f0[0], f0[1], f0[2], 0.0f);
    }
    // All pairs in the reference's (j,k) order: adds the pair forces of the positions in `pos` to
    // the accumulators in `frc` (whose xyz must be zero on entry).
    auto pair_pass = [&](float &dmin, uint32_t &in_step) {
      for (uint32_t j = 0; j + 1 < a; ++j) {
      const float4 rj4 = COL(pos, j);
      float4 fj4 = COL(frc, j);
      const float rj[3] = {rj4.x, rj4.y, rj4.z};
      float fj[3] = {fj4.x, fj4.y, fj4.z};
      // partners four at a time: all loads, then geometry and gathers, then the ordered
      // accumulation, then all stores -- so that four pairs are in flight together
      for (uint32_t k0 = j + 1; k0 < a; k0 += 4) {
        float4 rk4[4], fk4[4];
        PairGeom g[4];
        float tv[4];
        bool valid[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          valid[u] = k0 + u < a;
          const uint32_t kk = valid[u] ? k0 + u : a - 1u;
          rk4[u] = COL(pos, kk);
          fk4[u] = COL(frc, kk);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float rk[3] = {rk4[u].x, rk4[u].y, rk4[u].z};
          g[u] = pair_geom<P>(c, rj, rk);
          if (!valid[u]) { g[u].bin = (unsigned)c.grid_len; g[u].inside = false; }  // the zero entry
          tv[u] = pair_gather<P>(v, g[u]);
          dmin = fminf(dmin, valid[u] ? g[u].d : 3.0e38f);
          if (P::count) in_step += g[u].inside ? 1u : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float fk[3] = {fk4[u].x, fk4[u].y, fk4[u].z};
          pair_apply<P>(pair_scale<P>(tv[u], g[u].d), g[u], fj, fk);  // an invalid slot adds +-0 to fj
          fk4[u] = make_float4(fk[0], fk[1], fk[2], fk4[u].w);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (valid[u]) COL(frc, k0 + u) = fk4[u];
      }
      COL(frc, j) = make_float4(fj[0], fj[1], fj[2], fj4.w);
    }
    };
    if (P::turbo) {  // rebuild the scaled forces from the positions (see rebuild_forces)
      for (uint32_t s = 0; s < a; ++s) COL(frc, s) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      float dmin_unused = 3.0e38f;
      uint32_t in_unused = 0;
      pair_pass(dmin_unused, in_unused);
    }
    uint32_t steps = m.steps;
    float time = m.time;
    bool met = false;
    while (steps < m.step_stop) {
      const bool rn = renorm_at<P>(steps);
      // first half kick + drift
#pragma unroll 2
      for (uint32_t s = 0; s < a; ++s) {
        const float4 r4 = COL(pos, s), f4 = COL(frc, s), q4 = COL(qq, s), w4 = COL(ww, s);
        const float r[3] = {r4.x, r4.y, r4.z}, f[3] = {f4.x, f4.y, f4.z};
        float hk[3], wh[3], q[4] = {q4.x, q4.y, q4.z, q4.w}, r_new[3];
        half_kick<P>(c, r, f, hk);
        wh[0] = P::add(w4.x, hk[0]); wh[1] = P::add(w4.y, hk[1]); wh[2] = P::add(w4.z, hk[2]);
        drift<P>(c, wh, q, r_new, rn);
        COL(qq, s) = make_float4(q[0], q[1], q[2], q[3]);
        COL(pos, s) = make_float4(r_new[0], r_new[1], r_new[2], r4.x);
        COL(frc, s) = make_float4(0.0f, 0.0f, 0.0f, r4.y);
        COL(ww, s) = make_float4(w4.x, w4.y, w4.z, r4.z);
        COL(whh, s) = make_float4(wh[0], wh[1], wh[2], 0.0f);
      }
      // all pairs, (j,k) order
      float dmin = 3.0e38f;
      uint32_t in_step = 0;
      pair_pass(dmin, in_step);
      if (dmin < c.icd2) { met = true; break; }
      // second half kick
#pragma unroll 2
      for (uint32_t s = 0; s < a; ++s) {
        const float4 r4 = COL(pos, s), f4 = COL(frc, s), wh4 = COL(whh, s);
        const float r[3] = {r4.x, r4.y, r4.z}, f[3] = {f4.x, f4.y, f4.z};
        float hk[3];
        half_kick<P>(c, r, f, hk);
        COL(ww, s) = make_float4(P::add(wh4.x, hk[0]), P::add(wh4.y, hk[1]), P::add(wh4.z, hk[2]), 0.0f);
      }
      time = __fadd_rn(time, c.dt);
      ++steps;
      inrange_total += in_step;
    }
    done_steps = steps - m.steps;
    for (uint32_t s = 0; s < a; ++s) {
      heroam_particle *p = base + act[s];
      const float4 r4 = COL(pos, s), f4 = COL(frc, s), q4 = COL(qq, s), w4 = COL(ww, s), wh4 = COL(whh, s);
      const float q[4] = {q4.x, q4.y, q4.z, q4.w}, r[3] = {r4.x, r4.y, r4.z}, w[3] = {w4.x, w4.y, w4.z};
      if (met) {
        const float wh[3] = {wh4.x, wh4.y, wh4.z}, r_old[3] = {r4.w, f4.w, w4.w};
        write_midstep<P>(c, p, q, r, w, wh, r_old, time, steps > 0);
      } else if (done_steps > 0) {
        const float f[3] = {f4.x, f4.y, f4.z};
        write_clean<P>(c, p, q, r, w, f, time, act[s] + 1u < m.no);
      }
    }
    if (met) m.status = ST_MIDSTEP;
    m.steps = steps;
    m.time = time;
    v.meta[t] = m;
  }
#undef COL
  if (done_steps) {
    atomicAdd(&v.counters->particle_steps, done_steps * a);
    atomicAdd(&v.counters->pair_steps, done_steps * ((unsigned long long)a * (a - 1) / 2));
    atomicAdd(&v.counters->inrange_steps, inrange_total);
  }
}

// ---------------------------------------------------------------------------------
// k_integrate_warp<PPL>: one warp per trajectory, up to 32*PPL active particles; lane l
// owns active particles l, l+32, ...  Post-drift positions are staged in shared memory
// (16 B per particle, broadcast reads).  Every lane accumulates the force on ITS particle
// over all partners in ascending partner order, which is exactly the order in which the
// reference's (j,k) double loop touches that particle (first the F_k -= f of every j < k,
// then the F_j += f of every k > j, simulation.c:205-231); since -(a-b) == (b-a) and
// s*(-x) == -(s*x) in IEEE arithmetic the result has the reference's bits without any
// cross-lane reduction.  Each pair is therefore evaluated by both of its lanes.
// ---------------------------------------------------------------------------------
template <int PPL, class P>
__global__ void __launch_bounds__(32 * WARP_KERNEL_WARPS)
k_integrate_warp(DevView v, const uint32_t *__restrict__ list, uint32_t count, bool use_lists) {
  __shared__ float4 s_pos[WARP_KERNEL_WARPS][32 * PPL];
  const unsigned lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  const uint32_t item = blockIdx.x * WARP_KERNEL_WARPS + wib;
  if (item >= count) return;  // whole warp leaves together
  const Consts &c = v.c;
  const uint32_t t = list[item];
  TrajMeta m = v.meta[t];
  const uint32_t a = m.n_active;
  float4 *pos = s_pos[wib];
  heroam_particle *rec[PPL];
  bool own[PPL];
  float q[PPL][4], w[PPL][3], r[PPL][3], f[PPL][3], hk[PPL][3], wh[PPL][3], r_old[PPL][3];
#pragma unroll
  for (int i = 0; i < PPL; ++i) {
    const uint32_t s = lane + 32u * i;
    own[i] = s < a;
    rec[i] = v.particles + m.first + (own[i] ? v.act[m.first + s] : 0u);
    if (own[i]) {
      const heroam_particle *p = rec[i];
      q[i][0] = p->orientation[0]; q[i][1] = p->orientation[1];
      q[i][2] = p->orientation[2]; q[i][3] = p->orientation[3];
      ld3(p->position, r[i]);
      load_w<P>(c, p->ang_velocity, w[i]);
      load_f<P>(c, p->force, f[i]);
      half_kick<P>(c, r[i], f[i], hk[i]);
    } else {
      q[i][0] = 1.0f; q[i][1] = q[i][2] = q[i][3] = 0.0f;
#pragma unroll
      for (int d = 0; d < 3; ++d) { r[i][d] = 0.0f; w[i][d] = 0.0f; f[i][d] = 0.0f; hk[i][d] = 0.0f; }
    }
  }
  // partner lists built by the resolve kernel for exactly this segment
  const uint8_t *nbr[PPL];
  uint32_t n_nbr[PPL];
  uint32_t my_len = 0, my_pairs = 0;
  bool listed_lane = true;
#pragma unroll
  for (int i = 0; i < PPL; ++i) {
    const uint32_t s = lane + 32u * i;
    n_nbr[i] = !use_lists ? NBR_ALL : (own[i] ? v.nbr_cnt[m.first + s] : 0u);
    nbr[i] = v.nbr_pool + ((use_lists && own[i]) ? v.nbr_start[m.first + s] : 0u);
    if (n_nbr[i] == NBR_ALL) { listed_lane = false; n_nbr[i] = 0; }
    my_len = max(my_len, n_nbr[i]);
    my_pairs += n_nbr[i];
  }
  const bool listed = __all_sync(0xffffffffu, listed_lane);
  const uint32_t loop_len = listed ? __reduce_max_sync(0xffffffffu, my_len) : a;
  // pairs evaluated per step (each listed pair sits in two lists)
  const unsigned long long evaluated = listed ? __reduce_add_sync(0xffffffffu, my_pairs) / 2u
                                              : (unsigned long long)a * (a - 1) / 2;
  // Every lane adds to f[i] the forces of the (listed) partners of its particles, positions taken
  // from `pos`; f[i] must be zero on entry.
  auto partner_pass = [&](float &dmin, uint32_t &in_step) {
    // partners in ascending order, four at a time: geometry and gathers first, then the
    // (ordered) accumulation; everything branch-free, invalid slots contribute zero.
    // With partner lists (the usual case) slot k of the loop is the k-th listed partner of
    // each lane's particle; without, it is particle k itself.
    constexpr int UNR = 4;
    for (uint32_t p0 = 0; p0 < loop_len; p0 += UNR) {
      PairGeom g[UNR][PPL];
      float tv[UNR][PPL];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
#pragma unroll
        for (int i = 0; i < PPL; ++i) {
          const uint32_t s = lane + 32u * i;
          const uint32_t k = p0 + u;
          uint32_t p = k;
          bool valid = own[i];
          if (listed) {
            valid = valid && k < n_nbr[i];
            p = valid ? (uint32_t)nbr[i][k] : s;
          } else {
            valid = valid && k < a && s != k;
            p = k < a ? k : a - 1u;
          }
          const float4 rp4 = pos[own[i] ? p : 0u];
          const float rp[3] = {rp4.x, rp4.y, rp4.z};
          g[u][i] = pair_geom<P>(c, r[i], rp);  // own particle first: sign of (r_s - r_p)
          if (!valid) { g[u][i].bin = (unsigned)c.grid_len; g[u][i].inside = false; }  // the zero entry
          tv[u][i] = pair_gather<P>(v, g[u][i]);
          dmin = fminf(dmin, valid ? g[u][i].d : 3.0e38f);
          if (P::count) in_step += (g[u][i].inside && s < p) ? 1u : 0u;  // count each pair once
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u)
#pragma unroll
        for (int i = 0; i < PPL; ++i) {
          // a masked slot may have d == 0 (the particle itself): keep sqrt away from it
          const float dsafe = g[u][i].d > 0.0f ? g[u][i].d : 1.0f;
          pair_apply_own<P>(pair_scale<P>(tv[u][i], dsafe), g[u][i], f[i]);
        }
    }
  };
  if (P::turbo) {  // rebuild the scaled forces from the positions (see rebuild_forces)
#pragma unroll
    for (int i = 0; i < PPL; ++i) {
      if (own[i]) pos[lane + 32u * i] = make_float4(r[i][0], r[i][1], r[i][2], 0.0f);
      f[i][0] = 0.0f; f[i][1] = 0.0f; f[i][2] = 0.0f;
    }
    __syncwarp();
    float dmin_unused = 3.0e38f;
    uint32_t in_unused = 0;
    partner_pass(dmin_unused, in_unused);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < PPL; ++i)
      if (own[i]) half_kick<P>(c, r[i], f[i], hk[i]);
  }
  uint32_t steps = m.steps;
  float time = m.time;
  bool met = false;
  unsigned long long inrange_total = 0;
  while (steps < m.step_stop) {
    const bool rn = renorm_at<P>(steps);  // one trajectory per warp: uniform
#pragma unroll
    for (int i = 0; i < PPL; ++i) {
      if (own[i]) {
        r_old[i][0] = r[i][0]; r_old[i][1] = r[i][1]; r_old[i][2] = r[i][2];
        wh[i][0] = P::add(w[i][0], hk[i][0]);
        wh[i][1] = P::add(w[i][1], hk[i][1]);
        wh[i][2] = P::add(w[i][2], hk[i][2]);
        drift<P>(c, wh[i], q[i], r[i], rn);
        pos[lane + 32u * i] = make_float4(r[i][0], r[i][1], r[i][2], 0.0f);
      }
      f[i][0] = 0.0f; f[i][1] = 0.0f; f[i][2] = 0.0f;
    }
    __syncwarp();
    float dmin = 3.0e38f;
    uint32_t in_step = 0;
    partner_pass(dmin, in_step);
    __syncwarp();
    if (__any_sync(0xffffffffu, dmin < c.icd2)) { met = true; break; }
#pragma unroll
    for (int i = 0; i < PPL; ++i) {
      if (own[i]) {
        half_kick<P>(c, r[i], f[i], hk[i]);
        w[i][0] = P::add(wh[i][0], hk[i][0]);
        w[i][1] = P::add(wh[i][1], hk[i][1]);
        w[i][2] = P::add(wh[i][2], hk[i][2]);
      }
    }
    time = __fadd_rn(time, c.dt);
    ++steps;
    inrange_total += in_step;
  }
  const unsigned long long done_steps = steps - m.steps;
#pragma unroll
  for (int i = 0; i < PPL; ++i) {
    if (!own[i]) continue;
    heroam_particle *p = rec[i];
    if (met) {
      write_midstep<P>(c, p, q[i], r[i], w[i], wh[i], r_old[i], time, steps > 0);
    } else if (done_steps > 0) {
      const uint32_t j = (uint32_t)(p - (v.particles + m.first));
      write_clean<P>(c, p, q[i], r[i], w[i], f[i], time, j + 1u < m.no);
    }
  }
  const unsigned long long ir = warp_sum_u64(inrange_total);
  if (lane == 0) {
    if (met) m.status = ST_MIDSTEP;
    m.steps = steps;
    m.time = time;
    v.meta[t] = m;
    if (done_steps) {
      atomicAdd(&v.counters->particle_steps, done_steps * a);
      atomicAdd(&v.counters->pair_steps, done_steps * ((unsigned long long)a * (a - 1) / 2));
      atomicAdd(&v.counters->pair_skipped, done_steps * ((unsigned long long)a * (a - 1) / 2 - evaluated));
      atomicAdd(&v.counters->inrange_steps, ir);
    }
  }
}

// ---------------------------------------------------------------------------------
// k_integrate_records: any active count, one thread per trajectory, state left in the
// global records (slow; the correctness anchor for the wide kernels).
// ---------------------------------------------------------------------------------
template <class P>
__global__ void k_integrate_records(DevView v, const uint32_t *__restrict__ list, uint32_t count) {
  const uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= count) return;
  const Consts &c = v.c;
  const uint32_t t = list[item];
  TrajMeta m = v.meta[t];
  heroam_particle *base = v.particles + m.first;
  const uint32_t *act = v.act + m.first;
  const uint32_t a = m.n_active;
  unsigned long long done_steps = 0, inrange_total = 0;
  while (m.steps < m.step_stop) {
    for (uint32_t s = 0; s < a; ++s) {
      heroam_particle *p = base + act[s];
      float hk[3], wh[3], q[4], r[3];
      half_kick<P>(c, p->position, p->force, hk);
      wh[0] = P::add(p->ang_velocity[0], hk[0]);
      wh[1] = P::add(p->ang_velocity[1], hk[1]);
      wh[2] = P::add(p->ang_velocity[2], hk[2]);
      q[0] = p->orientation[0]; q[1] = p->orientation[1]; q[2] = p->orientation[2]; q[3] = p->orientation[3];
      drift<P>(c, wh, q, r, renorm_at<P>(m.steps));
      p->orientation[0] = q[0]; p->orientation[1] = q[1]; p->orientation[2] = q[2]; p->orientation[3] = q[3];
      st3(p->position, r);
      st3(p->force, wh);  // scratch slot: w(t + dt/2)
    }
    bool met = false;
    for (uint32_t js = 0; js < a && !met; ++js)
      for (uint32_t ks = js + 1; ks < a; ++ks)
        if (pair_dist_sq<P>(base[act[js]].position, base[act[ks]].position) < c.icd2) { met = true; break; }
    if (met) { m.status = ST_MIDSTEP; break; }
    for (uint32_t s = 0; s < a; ++s) {
      heroam_particle *p = base + act[s];
      st3(p->velocity, p->force);  // keep w(t + dt/2) while the forces are rebuilt
      p->force[0] = 0.0f; p->force[1] = 0.0f; p->force[2] = 0.0f;
    }
    uint32_t in_step = 0;
    for (uint32_t js = 0; js + 1 < a; ++js) {
      heroam_particle *pj = base + act[js];
      for (uint32_t ks = js + 1; ks < a; ++ks) {
        heroam_particle *pk = base + act[ks];
        float fj[3], fk[3];
        ld3(pj->force, fj);
        ld3(pk->force, fk);
        pair_force<P>(c, v.table, pj->position, pk->position, fj, fk, in_step);
        st3(pj->force, fj);
        st3(pk->force, fk);
      }
    }
    m.time = __fadd_rn(m.time, c.dt);
    for (uint32_t s = 0; s < a; ++s) {
      heroam_particle *p = base + act[s];
      float hk[3], w[3], vel[3], v2;
      half_kick<P>(c, p->position, p->force, hk);
      w[0] = P::add(p->velocity[0], hk[0]);
      w[1] = P::add(p->velocity[1], hk[1]);
      w[2] = P::add(p->velocity[2], hk[2]);
      st3(p->ang_velocity, w);
      stored_velocity<P>(p->position, w, vel, v2);
      st3(p->velocity, vel);
      p->velocity_sq = v2;
      p->time = m.time;
      p->force_mag = (act[s] + 1u < m.no) ? vec_norm(p->force) : 0.0f;
    }
    ++m.steps;
    ++done_steps;
    inrange_total += in_step;
  }
  v.meta[t] = m;
  if (done_steps) {
    atomicAdd(&v.counters->particle_steps, done_steps * a);
    atomicAdd(&v.counters->pair_steps, done_steps * ((unsigned long long)a * (a - 1) / 2));
    atomicAdd(&v.counters->inrange_steps, inrange_total);
  }
}

// ---------------------------------------------------------------------------------
// k_integrate_capture: the saveflag = 1 path (simulation.c:460-476).  One thread per
// trajectory, state in the global records exactly as k_integrate_records keeps it (every
// field of every record is the reference's at each step boundary); at the top of every loop
// iteration all `no` records of the trajectory -- frozen ones included -- are copied to the
// snapshot buffer, which is what the reference hands to save_timestep_to_hdf5 there.
// Snapshot i of trajectory t in this pass lives at snap[(first * chunk) + i * no .. + no).
// ---------------------------------------------------------------------------------
struct CaptureView {
  heroam_particle *snap;   // n_particles * chunk records
  uint32_t *taken;         // [t] snapshots written in this pass
  uint32_t *from_step;     // [t] loop iteration of the first of them
  uint32_t chunk;
};

template <class P>
__global__ void k_integrate_capture(DevView v, CaptureView cap, const uint32_t *__restrict__ list, uint32_t count) {
  const uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= count) return;
  const Consts &c = v.c;
  const uint32_t t = list[item];
  TrajMeta m = v.meta[t];
  heroam_particle *base = v.particles + m.first;
  const uint32_t *act = v.act + m.first;
  const uint32_t a = m.n_active;
  heroam_particle *snap = cap.snap + (size_t)m.first * cap.chunk;
  uint32_t taken = 0;
  cap.from_step[t] = m.steps;
  unsigned long long done_steps = 0, inrange_total = 0;
  while (m.steps < m.step_stop && taken < cap.chunk) {
    {  // the state at the top of iteration m.steps
      const uint32_t *src = reinterpret_cast<const uint32_t *>(base);
      uint32_t *dst = reinterpret_cast<uint32_t *>(snap + (size_t)taken * m.no);
      const uint32_t words = m.no * (uint32_t)(sizeof(heroam_particle) / sizeof(uint32_t));
      for (uint32_t i = 0; i < words; ++i) dst[i] = src[i];
      ++taken;
    }
    for (uint32_t s = 0; s < a; ++s) {
      heroam_particle *p = base + act[s];
      float hk[3], wh[3], q[4], r[3];
      half_kick<P>(c, p->position, p->force, hk);
      wh[0] = P::add(p->ang_velocity[0], hk[0]);
      wh[1] = P::add(p->ang_velocity[1], hk[1]);
      wh[2] = P::add(p->ang_velocity[2], hk[2]);
      q[0] = p->orientation[0]; q[1] = p->orientation[1]; q[2] = p->orientation[2]; q[3] = p->orientation[3];
      drift<P>(c, wh, q, r, renorm_at<P>(m.steps));
      p->orientation[0] = q[0]; p->orientation[1] = q[1]; p->orientation[2] = q[2]; p->orientation[3] = q[3];
      st3(p->position, r);
      st3(p->force, wh);  // scratch slot: w(t + dt/2)
    }
    bool met = false;
    for (uint32_t js = 0; js < a && !met; ++js)
      for (uint32_t ks = js + 1; ks < a; ++ks)
        if (pair_dist_sq<P>(base[act[js]].position, base[act[ks]].position) < c.icd2) { met = true; break; }
    if (met) { m.status = ST_MIDSTEP; break; }  // the resolve kernel finishes this iteration
    for (uint32_t s = 0; s < a; ++s) {
      heroam_particle *p = base + act[s];
      st3(p->velocity, p->force);  // keep w(t + dt/2) while the forces are rebuilt
      p->force[0] = 0.0f; p->force[1] = 0.0f; p->force[2] = 0.0f;
    }
    uint32_t in_step = 0;
    for (uint32_t js = 0; js + 1 < a; ++js) {
      heroam_particle *pj = base + act[js];
      for (uint32_t ks = js + 1; ks < a; ++ks) {
        heroam_particle *pk = base + act[ks];
        float fj[3], fk[3];
        ld3(pj->force, fj);
        ld3(pk->force, fk);
        pair_force<P>(c, v.table, pj->position, pk->position, fj, fk, in_step);
        st3(pj->force, fj);
        st3(pk->force, fk);
      }
    }
    m.time = __fadd_rn(m.time, c.dt);
    for (uint32_t s = 0; s < a; ++s) {
      heroam_particle *p = base + act[s];
      float hk[3], w[3], vel[3], v2;
      half_kick<P>(c, p->position, p->force, hk);
      w[0] = P::add(p->velocity[0], hk[0]);
      w[1] = P::add(p->velocity[1], hk[1]);
      w[2] = P::add(p->velocity[2], hk[2]);
      st3(p->ang_velocity, w);
      stored_velocity<P>(p->position, w, vel, v2);
      st3(p->velocity, vel);
      p->velocity_sq = v2;
      p->time = m.time;
      p->force_mag = (act[s] + 1u < m.no) ? vec_norm(p->force) : 0.0f;
    }
    ++m.steps;
    ++done_steps;
    inrange_total += in_step;
  }
  cap.taken[t] = taken;
  v.meta[t] = m;
  if (done_steps) {
    atomicAdd(&v.counters->particle_steps, done_steps * a);
    atomicAdd(&v.counters->pair_steps, done_steps * ((unsigned long long)a * (a - 1) / 2));
    atomicAdd(&v.counters->inrange_steps, inrange_total);
  }
}

// ---------------------------------------------------------------------------------
// k_results: per-trajectory outcome for the host
// ---------------------------------------------------------------------------------
__global__ void k_results(DevView v, heroam_traj_result *__restrict__ out) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= v.n_traj) return;
  const TrajMeta m = v.meta[t];
  heroam_traj_result r;
  r.steps = m.steps;
  r.status = m.status >= ST_DONE ? m.status - ST_DONE : 0xffffffffu;
  r.n_flagged = m.n_flagged;
  r.time = m.time;
  out[t] = r;
}

// the Turbo kernels' table: every entry times kappa (the appended zero stays zero)
__global__ void k_scale_table(float *__restrict__ dst, const float *__restrict__ src, unsigned n, float kappa) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i] * kappa;
}

// ---------------------------------------------------------------------------------
// k_fma_peak: register-resident FFMA chains, 8 independent accumulators per thread.
// The measured FP32 ceiling that roofline fractions are quoted against.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fma_peak(float *out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  const float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 12345.678f) out[0] = s;  // never true in practice; keeps the chains alive
}

// The same with three REGISTER operands per FFMA (multiplier and addend come from memory, not
// from the constant bank): the form almost every FFMA of the integrate kernels has.
__global__ void __launch_bounds__(256) k_fma_peak_rrr(float *out, const float *__restrict__ coef, int iters) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  const float a0 = coef[0], a1 = coef[1], a2 = coef[2], a3 = coef[3], b0 = coef[4], b1 = coef[5], b2 = coef[6], b3 = coef[7];
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fmaf(x0, a0, b0); x1 = fmaf(x1, a1, b1); x2 = fmaf(x2, a2, b2); x3 = fmaf(x3, a3, b3);
      x4 = fmaf(x4, a0, b1); x5 = fmaf(x5, a1, b2); x6 = fmaf(x6, a2, b3); x7 = fmaf(x7, a3, b0);
    }
  }
  const float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 12345.678f) out[0] = s;
}

// Packed FP32 (FFMA2 on 64-bit register pairs, sm_100): the same chains, two per instruction.
__global__ void __launch_bounds__(256) k_fma_peak_x2(float *out, const float *__restrict__ coef, int iters) {
  const float t = threadIdx.x;
  float2 x0 = make_float2(t, t + 1), x1 = make_float2(t + 2, t + 3), x2 = make_float2(t + 4, t + 5), x3 = make_float2(t + 6, t + 7);
  float2 x4 = make_float2(t + 8, t + 9), x5 = make_float2(t + 10, t + 11), x6 = make_float2(t + 12, t + 13), x7 = make_float2(t + 14, t + 15);
  const float2 a0 = make_float2(coef[0], coef[1]), a1 = make_float2(coef[2], coef[3]);
  const float2 b0 = make_float2(coef[4], coef[5]), b1 = make_float2(coef[6], coef[7]);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = __ffma2_rn(x0, a0, b0); x1 = __ffma2_rn(x1, a1, b1); x2 = __ffma2_rn(x2, a0, b1); x3 = __ffma2_rn(x3, a1, b0);
      x4 = __ffma2_rn(x4, a0, b0); x5 = __ffma2_rn(x5, a1, b1); x6 = __ffma2_rn(x6, a0, b1); x7 = __ffma2_rn(x7, a1, b0);
    }
  }
  const float s = ((x0.x + x1.x) + (x2.x + x3.x)) + ((x4.x + x5.x) + (x6.x + x7.x)) + ((x0.y + x1.y) + (x2.y + x3.y)) + ((x4.y + x5.y) + (x6.y + x7.y));
  if (s == 12345.678f) out[0] = s;
}

}  // namespace heroam
