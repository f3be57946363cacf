// heroam_sampler.cuh -- Philox sampling stage and fused statistics (included by heroam_gpu.cu).
//
// The sampling stage restates initialize_particles (reference simulation.c:277-337 with
// :238-275 and vector.c:74-134) with the libc rand() stream replaced by a counter-based
// Philox4x32-10 stream per trajectory: key = (seed, he_number high word), counter =
// (block, trajectory index, he_number low word).  Each 32-bit output is shifted to the 31
// bits of glibc's rand() and then goes through the reference's own conversions, so the
// arithmetic downstream of the integer draw is the reference's.  oracle/heroam_oracle.c
// (orc_sample, source = 1) is the CPU twin of this file.
#pragma once

namespace heroam {

struct Philox {
  uint32_t key0, key1, c1, c2, c3;
  uint32_t block;   // next block number (counter word 0)
  uint32_t buf[4];
  int have;

  __device__ __forceinline__ void begin(long long seed, unsigned long long he_number, unsigned long long traj) {
    key0 = (uint32_t)((unsigned long long)seed & 0xffffffffu);
    key1 = (uint32_t)((unsigned long long)seed >> 32) ^ (uint32_t)(he_number >> 32);
    c1 = (uint32_t)(traj & 0xffffffffu);
    c2 = (uint32_t)(traj >> 32);
    c3 = (uint32_t)(he_number & 0xffffffffu);
    block = 0;
    have = 0;
  }
  __device__ __forceinline__ void refill() {
    uint32_t x0 = block, x1 = c1, x2 = c2, x3 = c3, k0 = key0, k1 = key1;
#pragma unroll
    for (int round = 0; round < 10; ++round) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
      const uint32_t n0 = hi1 ^ x1 ^ k0, n2 = hi0 ^ x3 ^ k1;
      x0 = n0; x1 = lo1; x2 = n2; x3 = lo0;
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    buf[0] = x0; buf[1] = x1; buf[2] = x2; buf[3] = x3;
    ++block;
    have = 4;
  }
  // 31-bit integer with the range of glibc rand()
  __device__ __forceinline__ int next() {
    if (have == 0) refill();
    const uint32_t word = have == 4 ? buf[0] : have == 3 ? buf[1] : have == 2 ? buf[2] : buf[3];
    --have;
    return (int)(word >> 1);
  }
  // resume after `consumed` draws
  __device__ __forceinline__ void skip_to(uint32_t consumed) {
    block = consumed >> 2;
    have = 0;
    const int within = (int)(consumed & 3u);
    if (within) {
      refill();
      have = 4 - within;
    }
  }
  __device__ __forceinline__ uint32_t consumed() const { return block * 4u - (uint32_t)have; }
};

// (float)rand() / RAND_MAX: int -> float, RAND_MAX -> float (2^31), float divide
__device__ __forceinline__ float unit_f(Philox &g) { return __fdiv_rn(__int2float_rn(g.next()), 2147483648.0f); }
// draws that feed a logarithm come from (0,1] (the reference can feed 0, simulation.c:249-251)
__device__ __forceinline__ float unit_f_pos(Philox &g) {
  int r = g.next();
  if (r == 0) r = 1;
  return __fdiv_rn(__int2float_rn(r), 2147483648.0f);
}

// zero-truncated Poisson by Knuth's product of uniforms (simulation.c:257-275, :302-304)
__device__ inline uint32_t sample_count(Philox &g, double limit) {
  int n;
  do {
    n = 0;
    double p = 1.0;
    do {
      ++n;
      const double u = __ddiv_rn((double)g.next(), 2147483647.0);
      p = __dmul_rn(p, u);
    } while (p > limit);
    --n;
  } while (n < 1);
  return (uint32_t)n;
}

// q (0,0,1) q^-1 exactly as the reference evaluates it (vector.c:94-110, :182-205)
__device__ inline void pole_image_exact(const float q[4], float out[3]) {
  typedef Exact P;
  const float n2 = P::dot4(q[0], q[1], q[2], q[3]);
  const float ia = __fdiv_rn(q[0], n2), ix = -__fdiv_rn(q[1], n2), iy = -__fdiv_rn(q[2], n2), iz = -__fdiv_rn(q[3], n2);
  const float ua = -q[3], ux = q[2], uy = -q[1], uz = q[0];
  out[0] = P::sub(P::add(P::add(P::mul(ua, ix), P::mul(ux, ia)), P::mul(uy, iz)), P::mul(uz, iy));
  out[1] = P::add(P::add(P::sub(P::mul(ua, iy), P::mul(ux, iz)), P::mul(uy, ia)), P::mul(uz, ix));
  out[2] = P::add(P::sub(P::add(P::mul(ua, iz), P::mul(ux, iy)), P::mul(uy, ix)), P::mul(uz, ia));
}

// Marsaglia (1972) point on S^3 (vector.c:112-134)
__device__ inline void sample_orientation(Philox &g, float q[4]) {
  float x, y, z, u, v, w;
  do {
    x = __fsub_rn(__fmul_rn(2.0f, unit_f(g)), 1.0f);
    y = __fsub_rn(__fmul_rn(2.0f, unit_f(g)), 1.0f);
    z = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));
  } while (z > 1.0f);
  do {
    u = __fsub_rn(__fmul_rn(2.0f, unit_f(g)), 1.0f);
    v = __fsub_rn(__fmul_rn(2.0f, unit_f(g)), 1.0f);
    w = __fadd_rn(__fmul_rn(u, u), __fmul_rn(v, v));
  } while (w > 1.0f);
  const float sc = __fsqrt_rn(__fdiv_rn(__fsub_rn(1.0f, z), w));
  q[0] = x; q[1] = y; q[2] = __fmul_rn(sc, u); q[3] = __fmul_rn(sc, v);
  const float mag = __fsqrt_rn(Exact::dot4(q[0], q[1], q[2], q[3]));
  q[0] = __fdiv_rn(q[0], mag); q[1] = __fdiv_rn(q[1], mag);
  q[2] = __fdiv_rn(q[2], mag); q[3] = __fdiv_rn(q[3], mag);
}

// sigma * chi_3 through Box-Muller in double, as simulation.c:238-255
__device__ inline float sample_speed(Philox &g, float sigma) {
  const double pi = (double)3.14159274101257324f;  // float pi = 4*atan(1)
  const float u1 = unit_f_pos(g), u2 = unit_f(g), u3 = unit_f_pos(g), u4 = unit_f(g);
  const double r1 = sqrt(-2.0 * log((double)u1)), a1 = 2.0 * pi * (double)u2;
  const double r3 = sqrt(-2.0 * log((double)u3)), a3 = 2.0 * pi * (double)u4;
  const float z1 = (float)(r1 * cos(a1));
  const float z2 = (float)(r1 * sin(a1));
  const float z3 = (float)(r3 * cos(a3));
  return (float)(sqrt((double)Exact::dot3(z1, z2, z3)) * (double)sigma);
}

struct SampleParams {
  long long seed;
  unsigned long long he_number;
  unsigned long long first_index;
  double limit;        // exp(-lambda), host libm (simulation.c:263)
  double radius_init;  // 2.22 * pow((double)N, 1/3), double (simulation.c:297)
  float sigma;         // conf->velocity
};

// stage 1: excitation count per trajectory and how many draws it consumed
__global__ void k_sample_counts(SampleParams sp, uint32_t n_traj, uint32_t *__restrict__ counts,
                                uint32_t *__restrict__ consumed) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_traj) return;
  Philox g;
  g.begin(sp.seed, sp.he_number, sp.first_index + t);
  counts[t] = sample_count(g, sp.limit);
  consumed[t] = g.consumed();
}

// calculate_force on a freshly sampled trajectory (simulation.c:334, :158-236)
__device__ inline void initial_forces(const Consts &c, const float *__restrict__ table, heroam_particle *p, uint32_t n) {
  for (uint32_t j = 0; j < n; ++j) {
    if (p[j].success) continue;  // checked once per j: a freshly flagged j keeps scanning (Q3)
    for (uint32_t k = j + 1; k < n; ++k) {
      if (p[k].success) continue;
      if (pair_dist_sq<Exact>(p[j].position, p[k].position) < c.icd2) { p[j].success = 1; p[k].success = 1; }
    }
  }
  uint32_t unused = 0;
  for (uint32_t j = 0; j + 1 < n; ++j) {
    if (p[j].success) continue;
    for (uint32_t k = j + 1; k < n; ++k) {
      if (p[k].success) continue;
      float fj[3], fk[3];
      ld3(p[j].force, fj);
      ld3(p[k].force, fk);
      pair_force<Exact>(c, table, p[j].position, p[k].position, fj, fk, unused);
      st3(p[j].force, fj);
      st3(p[k].force, fk);
    }
    p[j].force_mag = vec_norm(p[j].force);
  }
}

// The n particles of one trajectory from its stream (positioned after the count draws),
// then the initial calculate_force (simulation.c:311-334).
__device__ inline void sample_trajectory(Philox &g, const SampleParams &sp, const Consts &c,
                                         const float *__restrict__ table, heroam_particle *p, uint32_t n) {
  const float radius_f = (float)sp.radius_init;
  for (uint32_t j = 0; j < n; ++j) {
    heroam_particle out;
    out.index = j;
    out.success = 0;
    out.time = 0.0f;  // Q1
    out.force_mag = 0.0f;
    sample_orientation(g, out.orientation);
    float unit[3];
    pole_image_exact(out.orientation, unit);
    out.position[0] = __fmul_rn(unit[0], radius_f);
    out.position[1] = __fmul_rn(unit[1], radius_f);
    out.position[2] = __fmul_rn(unit[2], radius_f);
    const float vel = (float)((double)sample_speed(g, sp.sigma) * 1E-5);  // m/s -> A/fs (:323)
    const float ang_vel = (float)((double)vel / sp.radius_init);           // :324
    out.velocity_sq = __fmul_rn(vel, vel);
    // create_perpend_vector (vector.c:74-92): a x r, a = (-y, x, 0) away from the poles
    float a[3];
    if (out.position[0] != 0.0f || out.position[1] != 0.0f) { a[0] = -out.position[1]; a[1] = out.position[0]; a[2] = 0.0f; }
    else { a[0] = 0.0f; a[1] = -out.position[2]; a[2] = out.position[1]; }
    float tan[3], tan2;
    stored_velocity<Exact>(a, out.position, tan, tan2);  // plain cross product
    const float k = __fdiv_rn(ang_vel, __fsqrt_rn(tan2));
    out.ang_velocity[0] = __fmul_rn(tan[0], k);
    out.ang_velocity[1] = __fmul_rn(tan[1], k);
    out.ang_velocity[2] = __fmul_rn(tan[2], k);
    float v2;
    stored_velocity<Exact>(out.position, out.ang_velocity, out.velocity, v2);  // r x w (:329)
    out.force[0] = 0.0f; out.force[1] = 0.0f; out.force[2] = 0.0f;
    p[j] = out;
  }
  initial_forces(c, table, p, n);
}

// stage 2: the particles of every trajectory (serial within a trajectory: one stream)
__global__ void k_sample_particles(SampleParams sp, Consts c, const float *__restrict__ table, uint32_t n_traj,
                                   const uint32_t *__restrict__ counts, const uint32_t *__restrict__ first,
                                   const uint32_t *__restrict__ consumed, heroam_particle *__restrict__ particles) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_traj) return;
  Philox g;
  g.begin(sp.seed, sp.he_number, sp.first_index + t);
  g.skip_to(consumed[t]);
  sample_trajectory(g, sp, c, table, particles + first[t], counts[t]);
}

// ---------------------------------------------------------------------------------
// statistics of a finished batch
// ---------------------------------------------------------------------------------
struct DevStats {
  unsigned long long hist_n[HEROAM_HIST_N];
  unsigned long long hist_traj_time[HEROAM_HIST_T];
  unsigned long long hist_meet_time[HEROAM_HIST_T];
  unsigned long long done[8];
  unsigned long long unmet_particles;
  unsigned long long particles;
  unsigned long long trajectories;
  unsigned long long sum_traj_time_q10;  // sum of the loop times at exit in units of 2^-10 fs: integer, so that the sum
                                         // does not depend on the order in which the atomics land (a double sum did,
                                         // in its last bits, from run to run)
};

__device__ __forceinline__ int time_bin(float t, float max_time) {
  int b = (int)(t / max_time * (float)HEROAM_HIST_T);
  return b < 0 ? 0 : (b >= HEROAM_HIST_T ? HEROAM_HIST_T - 1 : b);
}

__device__ inline void accumulate_stats(const DevView &v, const TrajMeta &m, DevStats *__restrict__ st) {
  atomicAdd(&st->hist_n[m.no < HEROAM_HIST_N ? m.no : HEROAM_HIST_N - 1], 1ull);
  atomicAdd(&st->hist_traj_time[time_bin(m.time, v.c.max_time)], 1ull);
  const uint32_t code = m.status >= ST_DONE ? m.status - ST_DONE : 7u;
  atomicAdd(&st->done[code < 8u ? code : 7u], 1ull);
  atomicAdd(&st->sum_traj_time_q10, (unsigned long long)__float2ll_rn(m.time * 1024.0f));
  unsigned long long unmet = 0;
  for (uint32_t j = 0; j < m.no; ++j) {
    const heroam_particle *p = v.particles + m.first + j;
    if (p->success) atomicAdd(&st->hist_meet_time[time_bin(p->time, v.c.max_time)], 1ull);
    else ++unmet;
  }
  if (unmet) atomicAdd(&st->unmet_particles, unmet);
  atomicAdd(&st->particles, (unsigned long long)m.no);
  atomicAdd(&st->trajectories, 1ull);
}

__global__ void k_stats(DevView v, DevStats *__restrict__ st) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= v.n_traj) return;
  accumulate_stats(v, v.meta[t], st);
}

// ---------------------------------------------------------------------------------
// Streaming pool (retire and refill).  The device holds `n_traj` SLOTS of `cap` records
// each.  Once per pass every slot is visited by one thread of k_pool_resolve:
//   * a live trajectory gets its pending meeting step finished and its exit evaluated;
//   * a finished one is folded into the histograms and its slot is freed;
//   * a free slot claims the next trajectory index from a global cursor, samples that
//     trajectory in place (Philox: count, particles, initial forces) and goes live;
//   * live slots are binned by active count for the integrate kernels.
// The heavy tail of trajectory lengths (1 .. 1e7 steps) therefore costs one drain at the
// end of a run instead of one per batch.
// ---------------------------------------------------------------------------------
constexpr uint32_t ST_FREE = 0xfffffff0u;

// A run of the pool covers one or more SWEEP POINTS (laser settings: different mean excitation
// numbers on the same droplet).  The points share one index space -- point 0's trajectories,
// then point 1's, ... -- so that one point's drain overlaps the next points' bulk; each point
// keeps its own Philox indices (first_index + i, exactly as if it were run on its own) and its
// own histogram block.
constexpr int MAX_SWEEP_POINTS = 32;

// One part of a partitioned pool: a droplet size of a size sweep (reference main.c:158-164) run as one pool run.
// The slots are divided evenly among the parts, slot t belongs to part t / slots_per_part and only ever holds
// that part's trajectories, so a (part, bin) list can be launched with the part's constants.
struct PartParams {
  Consts c;                       // radius, scales ... of this droplet size
  const float *table_k;           // the table scaled by this part's kappa (Turbo)
  unsigned long long he_number;   // Philox key / lambda of this size
  unsigned long long first_index; // first Philox trajectory index of this part (sharding)
  unsigned long long count;       // trajectories of this part on this GPU
  double limit;                   // exp(-lambda)
  double radius_init;             // 2.22 * pow(N, 1/3) in double (simulation.c:297)
};

struct PoolParams {
  SampleParams sp;              // sp.limit / sp.first_index are per point, see below
  unsigned long long *cursor;   // next position in the concatenated index space
  unsigned long long count;     // trajectories in the run, all points together
  uint32_t cap;                 // records per slot
  uint32_t n_points;
  DevStats *stats;              // n_points blocks
  uint32_t refill;              // bit p = 0: the index space of part p is used up (the host has seen free slots stay
                                // free); an ordinary run is part 0
  uint32_t n_parts;             // > 1: partitioned pool, `parts` and `slots_per_part` apply, `cursor` is an array
  uint32_t slots_per_part;
  const PartParams *parts;      // device memory
  unsigned long long point_end[MAX_SWEEP_POINTS];    // exclusive end of each point in the index space
  unsigned long long point_first[MAX_SWEEP_POINTS];  // its first Philox trajectory index
  double point_limit[MAX_SWEEP_POINTS];              // its exp(-lambda)
};

__global__ void k_pool_init(DevView v, uint32_t cap) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= v.n_traj) return;
  TrajMeta m;
  m.first = t * cap;
  m.no = 0; m.steps = 0; m.time = 0.0f; m.n_flagged = 0; m.n_active = 0; m.step_stop = 0;
  m.n_loners = 0; m.loner_from = 0; m.loner_until = 0; m.loner_time0 = 0.0f; m.cross_from = 0; m.core_slots = 0;
  m.point = 0;
  m.status = ST_FREE;
  v.meta[t] = m;
}

template <class P, class H>
__global__ void k_pool_resolve(DevView v, PoolParams pool, Segments sg, uint32_t step_cap,
                               uint32_t *__restrict__ bin_count, uint32_t *__restrict__ bin_list,
                               uint2 *__restrict__ free_list, LiveLists live) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t t = idx;
  Counters work = {};
  Schedule sc;
  uint32_t steps_done = 0;
  PartSlice ps = {bin_count, bin_count, 0u, 0u, 0u};
  if (idx < live.count) {
    if (live.in) t = live.in[idx];
    // a partitioned pool: this slot's part decides constants, sampling parameters, counters and list stretches
    SampleParams sp = pool.sp;
    unsigned long long *cursor = pool.cursor;
    unsigned long long part_count = pool.count;
    if (pool.n_parts > 1u) {
      const uint32_t part = t / pool.slots_per_part;
      const PartParams &pp = pool.parts[part];
      v.c = pp.c;
      v.table_k = pp.table_k;
      sp.he_number = pp.he_number;
      sp.radius_init = pp.radius_init;
      cursor = pool.cursor + part;
      part_count = pp.count;
      ps.bin_count = bin_count + part * N_BIN_COUNTERS;
      ps.part = part;
      ps.list_offset = part * pool.slots_per_part;
      ps.free_offset = part * pool.slots_per_part * pool.cap;
    }
    TrajMeta m = v.meta[t];
    // a freshly spawned trajectory can be over at once (nobody moving), hence the loop
    for (;;) {
      if (m.status < ST_DONE) {
        sc = resolve_live<P, H>(v, m, sg, step_cap, work);
        if (m.status < ST_DONE) break;  // still live: scheduled
      }
      if (m.status != ST_FREE) {
        accumulate_stats(v, m, pool.stats + m.point);
        m.status = ST_FREE;
      }
      // Once the index space is used up there is nothing to claim.  A plain load first: the free slots
      // (millions, late in a run) would otherwise all serialise on the cursor's atomic in every pass,
      // ~1 ns each -- 4 ms per pass with 4e6 free slots.  The cursor only grows, so a stale value
      // can only send a thread on to the atomic, never past a trajectory.
      if (!((pool.refill >> ps.part) & 1u)) break;
      if (*reinterpret_cast<volatile unsigned long long *>(cursor) >= part_count) break;
      const unsigned long long idx_claimed = atomicAdd(cursor, 1ull);
      if (idx_claimed >= part_count) break;
      uint32_t pt = 0;
      unsigned long long stream_index;
      double limit;
      if (pool.n_parts > 1u) {
        pt = ps.part;
        stream_index = pool.parts[pt].first_index + idx_claimed;
        limit = pool.parts[pt].limit;
      } else {
        while (pt + 1u < pool.n_points && idx_claimed >= pool.point_end[pt]) ++pt;
        stream_index = pool.point_first[pt] + (idx_claimed - (pt ? pool.point_end[pt - 1u] : 0ull));
        limit = pool.point_limit[pt];
      }
      Philox g;
      g.begin(sp.seed, sp.he_number, stream_index);
      const uint32_t n = sample_count(g, limit);
      m.point = pt;
      m.no = n; m.steps = 0; m.time = 0.0f; m.n_flagged = 0; m.step_stop = 0;
      m.n_loners = 0; m.loner_from = 0; m.loner_until = 0; m.loner_time0 = 0.0f; m.cross_from = 0; m.core_slots = 0;
      if (n > pool.cap) {  // does not fit a slot: counted, not integrated
        m.no = 0; m.n_active = 0;
        m.status = ST_DONE + HEROAM_DONE_UNSUPPORTED;
        work.unsupported += 1;
        continue;
      }
      heroam_particle *p = v.particles + m.first;
      sample_trajectory(g, sp, v.c, v.table, p, n);
      for (uint32_t j = 0; j < n; ++j) m.n_flagged += p[j].success ? 1u : 0u;
      m.n_active = n - m.n_flagged;
      m.core_slots = m.n_active;
      m.status = ST_LIVE;
    }
    v.meta[t] = m;
    steps_done = m.steps;
  }
  append_schedule(sc, t, v.n_traj, ps, bin_list, free_list, sg.free_stride, live.out, steps_done);
  flush_work(v, work);
}

}  // namespace heroam

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
static double host_lambda(const heroam_config &conf, unsigned long long he_number) {
  // simulation.c:289-293: float product and quotient first, then double
  double lam = (double)((conf.intensity * conf.pulse_width * conf.cross_section) / conf.hv);
  lam = (double)he_number * lam;
  lam = 6.241509074E-06 * lam;
  return lam;
}

// Sample `count` trajectories starting at stream index first_index into the context's
// device buffers (they become the resident batch).  Leaves counts on the host in *h_counts
// when requested.
static int sample_device(heroam_ctx *ctx, unsigned long long he_number, unsigned long long first_index, int count,
                         std::vector<uint32_t> *h_counts_out) {
  if (count < 0) return fail(HEROAM_E_ARG, "negative count");
  CU(cudaSetDevice(ctx->device));
  Consts c;
  int rc = make_consts(ctx->conf, he_number, &c);
  if (rc) return rc;
  const double lambda = host_lambda(ctx->conf, he_number);
  if (!(lambda > 0.0) || lambda > 700.0)
    return fail(HEROAM_E_ARG, "mean excitation number %g outside (0, 700]: Knuth's product of uniforms "
                "(simulation.c:257-275) is undefined there", lambda);
  SampleParams sp;
  sp.seed = (long long)ctx->conf.seed;
  sp.he_number = he_number;
  sp.first_index = first_index;
  sp.limit = exp(-1.0 * lambda);
  sp.radius_init = 2.22 * pow((double)he_number, 1.0 / 3.0);
  sp.sigma = ctx->conf.velocity;
  ctx->he_number = he_number;
  ctx->n_traj = (uint32_t)count;
  ctx->n_particles = 0;
  memset(&ctx->info, 0, sizeof ctx->info);
  if (count == 0) { ctx->uploaded = true; return HEROAM_OK; }
  rc = reserve(ctx, (size_t)count, ctx->cap_particles ? ctx->cap_particles : 1);
  if (rc) return rc;
  const unsigned grid = ((unsigned)count + 127u) / 128u;
  // the bin list is free at this point: borrow its first n words for the draw offsets
  uint32_t *d_consumed = ctx->d_bin_list;
  k_sample_counts<<<grid, 128, 0, ctx->main>>>(sp, (uint32_t)count, ctx->d_counts, d_consumed);
  CU(cudaGetLastError());
  std::vector<uint32_t> h_counts((size_t)count), h_first((size_t)count);
  CU(cudaMemcpyAsync(h_counts.data(), ctx->d_counts, (size_t)count * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->main));
  CU(cudaStreamSynchronize(ctx->main));
  size_t total = 0;
  for (int i = 0; i < count; ++i) {
    h_first[i] = (uint32_t)total;
    total += h_counts[i];
  }
  if (total > 0xfffffff0ull) return fail(HEROAM_E_ARG, "batch too large: more than 2^32 particle records");
  rc = reserve(ctx, (size_t)count, total);
  if (rc) return rc;
  CU(cudaMemcpyAsync(ctx->d_first, h_first.data(), (size_t)count * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->main));
  k_sample_particles<<<grid, 128, 0, ctx->main>>>(sp, c, ctx->d_table, (uint32_t)count, ctx->d_counts, ctx->d_first,
                                                   d_consumed, ctx->d_particles);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(ctx->main));
  ctx->info.kernel_launches += 2;
  ctx->n_particles = total;
  ctx->uploaded = true;
  if (h_counts_out) h_counts_out->swap(h_counts);
  return HEROAM_OK;
}

extern "C" int heroam_gpu_sample_flat(heroam_ctx *ctx, unsigned long long he_number, unsigned long long first_index,
                                      int count, uint32_t *counts, heroam_particle *flat, long capacity, long *total) {
  if (!ctx || !counts || !total) return fail(HEROAM_E_ARG, "heroam_gpu_sample_flat: null argument");
  std::vector<uint32_t> h_counts;
  int rc = sample_device(ctx, he_number, first_index, count, &h_counts);
  if (rc) return rc;
  if (count > 0) memcpy(counts, h_counts.data(), (size_t)count * sizeof(uint32_t));
  *total = (long)ctx->n_particles;
  if ((long)ctx->n_particles <= capacity && flat && ctx->n_particles)
    CU(cudaMemcpy(flat, ctx->d_particles, ctx->n_particles * sizeof(heroam_particle), cudaMemcpyDeviceToHost));
  return HEROAM_OK;
}

extern "C" int heroam_gpu_sample_batch(heroam_ctx *ctx, unsigned long long he_number, unsigned long long first_index,
                                       int count, heroam_particles *out) {
  if (!ctx || (!out && count > 0)) return fail(HEROAM_E_ARG, "heroam_gpu_sample_batch: null argument");
  std::vector<uint32_t> h_counts;
  int rc = sample_device(ctx, he_number, first_index, count, &h_counts);
  if (rc) return rc;
  std::vector<heroam_particle> flat(ctx->n_particles ? ctx->n_particles : 1);
  if (ctx->n_particles)
    CU(cudaMemcpy(flat.data(), ctx->d_particles, ctx->n_particles * sizeof(heroam_particle), cudaMemcpyDeviceToHost));
  size_t at = 0;
  for (int i = 0; i < count; ++i) {
    out[i].no = h_counts[i];
    out[i].index = (uint32_t)(first_index + (unsigned long long)i);
    // ownership as in the reference: the engine mallocs, the caller frees (simulation.c:310, :339-341)
    out[i].particle = (heroam_particle *)malloc((size_t)h_counts[i] * sizeof(heroam_particle));
    if (!out[i].particle) return fail(HEROAM_E_ARG, "out of host memory");
    memcpy(out[i].particle, flat.data() + at, (size_t)h_counts[i] * sizeof(heroam_particle));
    at += h_counts[i];
  }
  return HEROAM_OK;
}

// slot capacity of the streaming pool: room for lambda + 8 sigma (+8), at most MAX_ACTIVE
static uint32_t pool_slot_capacity(double lambda) {
  const double want = lambda + 8.0 * sqrt(lambda) + 8.0;
  uint32_t cap = 8;
  while (cap < want && cap < (uint32_t)MAX_ACTIVE) cap *= 2;
  return cap;
}

// host copy of the parts of a partitioned pool (empty for an ordinary run)
struct PoolParts {
  std::vector<PartParams> host;
};

template <class R, class H, class F = H>
static int run_pool(heroam_ctx *ctx, const Consts &c, const PoolParams &pool_in, uint32_t slots, const PoolParts &parts) {
  PoolParams pool = pool_in;
  const uint32_t n_parts = pool.n_parts > 1u ? pool.n_parts : 1u;
  pool.refill = n_parts >= 32u ? 0xffffffffu : (1u << n_parts) - 1u;
  LiveLists lists = {nullptr, slots, ctx->d_live_list};
  DevView v = make_view(ctx, c);
  if (H::turbo && n_parts == 1u) {
    int rc = refresh_scaled_table(ctx, c);
    if (rc) return rc;
  }
  const unsigned tgrid = (slots + 127) / 128;
  const uint32_t step_cap = ctx->opt.max_steps == 0 || ctx->opt.max_steps > 0xffffffffull ? 0xffffffffu
                                                                                         : (uint32_t)ctx->opt.max_steps;
  const size_t n_counters = (size_t)n_parts * N_BIN_COUNTERS;
  double integrate_ms = 0.0;
  bool pending_timing = false;
  const bool trace = getenv("HEROAM_TRACE") != nullptr;  // per-pass bin populations and times on stderr
  Regime rg;
  k_pool_init<<<tgrid, 128, 0, ctx->main>>>(v, pool.cap);
  CU(cudaGetLastError());
  ctx->info.kernel_launches++;
  for (;;) {
    CU(cudaMemsetAsync(ctx->d_bin_count, 0, n_counters * sizeof(uint32_t), ctx->main));
    CU(cudaMemsetAsync(ctx->d_nbr_used, 0, sizeof(unsigned long long), ctx->main));
    const Segments sg = pass_segments(ctx, rg);
    if (lists.count)
      k_pool_resolve<R, F><<<(lists.count + 127) / 128, 128, 0, ctx->main>>>(v, pool, sg, step_cap, ctx->d_bin_count,
                                                                           ctx->d_bin_list, ctx->d_free_list, lists);
    CU(cudaGetLastError());
    ctx->info.kernel_launches++;
    CU(cudaMemcpyAsync(ctx->h_bin_count, ctx->d_bin_count, n_counters * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->main));
    CU(cudaEventRecord(ctx->ev_r, ctx->main));
    CU(cudaStreamSynchronize(ctx->main));
    if (pending_timing) {
      float ms = 0;
      CU(cudaEventElapsedTime(&ms, ctx->ev_i0, ctx->ev_i1));
      integrate_ms += ms;
      const float pass_ms = ms;
      CU(cudaEventElapsedTime(&ms, ctx->ev_i1, ctx->ev_r));
      ctx->info.resolve_ms += ms;
      pending_timing = false;
      if (trace) {
        Counters now;
        cudaMemcpy(&now, ctx->d_counters, sizeof now, cudaMemcpyDeviceToHost);
        fprintf(stderr, " %.3f ms  psteps %llu free %llu resolve %.3f ms\n", (double)pass_ms, now.particle_steps, now.free_steps, ms);
      }
    }
    uint64_t live = 0, bin_total[N_BINS] = {};
    for (uint32_t p = 0; p < n_parts; ++p)
      for (int b = 0; b < N_BINS; ++b) {
        bin_total[b] += ctx->h_bin_count[(size_t)p * N_BIN_COUNTERS + b];
        live += ctx->h_bin_count[(size_t)p * N_BIN_COUNTERS + b];
      }
    if (live == 0) break;  // every slot is free and the index space is used up
    // a free slot claims a new trajectory in every pass while there are any: fewer live trajectories than slots
    // means none are left (per part, in a partitioned pool); once that holds everywhere only the slots listed
    // as live need a visit
    if (n_parts == 1u) {
      if (ctx->h_bin_count[LIVE_SLOT] < slots) pool.refill = 0;
    } else {
      for (uint32_t p = 0; p < n_parts; ++p)
        if (ctx->h_bin_count[(size_t)p * N_BIN_COUNTERS + PART_LIVE_SLOT] < pool.slots_per_part) pool.refill &= ~(1u << p);
    }
    update_regime(ctx, rg, bin_total);
    if (!pool.refill) {
      lists.in = lists.out;
      lists.count = ctx->h_bin_count[LIVE_SLOT];
      lists.out = lists.in == ctx->d_live_list ? ctx->d_live_list + ctx->cap_traj : ctx->d_live_list;
    }
    ctx->info.passes++;
    if (trace) {
      fprintf(stderr, "pass %llu live %llu behind %u bins", ctx->info.passes, (unsigned long long)live,
              0xffffffffu - ctx->h_bin_count[LAG_SLOT]);
      for (int b = 1; b < N_BINS; ++b) fprintf(stderr, " %llu", (unsigned long long)bin_total[b]);
    }
    CU(cudaEventRecord(ctx->ev_i0, ctx->main));
    for (int b = 1; b < N_BINS; ++b) {
      if (!bin_total[b]) continue;
      CU(cudaStreamWaitEvent(ctx->bin_stream[b], ctx->ev_i0, 0));
      for (uint32_t p = 0; p < n_parts; ++p) {  // the parts of a bin share its stream: one after the other
        const uint32_t cnt = ctx->h_bin_count[(size_t)p * N_BIN_COUNTERS + b];
        if (!cnt) continue;
        DevView vp = v;
        uint32_t list_offset = 0, free_offset = 0;
        if (n_parts > 1u) {
          vp.c = parts.host[p].c;
          vp.c.f_max = v.c.f_max;
          vp.table_k = parts.host[p].table_k;
          list_offset = p * pool.slots_per_part;
          free_offset = p * pool.slots_per_part * pool.cap;
        }
        int rc = launch_bin<H, F>(ctx, vp, b, cnt, rg.draining, list_offset, free_offset);
        if (rc) return rc;
        ctx->info.kernel_launches++;
      }
      CU(cudaEventRecord(ctx->bin_done[b], ctx->bin_stream[b]));
      CU(cudaStreamWaitEvent(ctx->main, ctx->bin_done[b], 0));
    }
    CU(cudaEventRecord(ctx->ev_i1, ctx->main));
    pending_timing = true;
  }
  ctx->info.integrate_ms = integrate_ms;
  return HEROAM_OK;
}

template <class T>
static int run_pool_mode(heroam_ctx *ctx, const Consts &c, const PoolParams &pool, uint32_t slots, const T &parts) {
  return ctx->opt.mode == HEROAM_MODE_EXACT          ? run_pool<Exact, Exact>(ctx, c, pool, slots, parts)
         : ctx->opt.mode == HEROAM_MODE_FAST_PLAIN   ? run_pool<Fast, Fast>(ctx, c, pool, slots, parts)
         : ctx->opt.mode == HEROAM_MODE_FAST_RELAXED ? run_pool<Fast, Turbo8>(ctx, c, pool, slots, parts)
         : ctx->opt.mode == HEROAM_MODE_FAST_CLOSED  ? run_pool<Fast, Turbo, TurboClosed>(ctx, c, pool, slots, parts)
         : ctx->opt.count_work                     ? run_pool<Fast, TurboCount>(ctx, c, pool, slots, parts)
                                                   : run_pool<Fast, Turbo>(ctx, c, pool, slots, parts);
}

static double point_lambda(const heroam_config &conf, unsigned long long he_number, float intensity, float pulse_width) {
  heroam_config c = conf;
  c.intensity = intensity;
  c.pulse_width = pulse_width;
  return host_lambda(c, he_number);
}

// the histograms and counters of a finished pool run, one block per point, onto the host
static int collect_pool_stats(heroam_ctx *ctx, int rc, const DevStats *d_stats, int n_points, heroam_stats *out) {
  std::vector<DevStats> h((size_t)n_points);
  if (rc == HEROAM_OK) {
    if (cudaMemcpyAsync(h.data(), d_stats, (size_t)n_points * sizeof(DevStats), cudaMemcpyDeviceToHost, ctx->main) != cudaSuccess ||
        cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->main) != cudaSuccess ||
        cudaEventRecord(ctx->ev_stop, ctx->main) != cudaSuccess || cudaStreamSynchronize(ctx->main) != cudaSuccess)
      rc = fail(HEROAM_E_CUDA, "statistics readback failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  if (rc == HEROAM_OK) {
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_stop);
    ctx->info.device_ms = ms;
    ctx->info.d2h_bytes = (size_t)n_points * sizeof(DevStats);
    ctx->info.particle_steps = ctx->h_counters->particle_steps;
    ctx->info.pair_steps = ctx->h_counters->pair_steps;
    ctx->info.inrange_steps = ctx->h_counters->inrange_steps;
    ctx->info.pair_skipped = ctx->h_counters->pair_skipped;
    ctx->info.free_steps = ctx->h_counters->free_steps;
    for (int p = 0; p < n_points; ++p) {
      ctx->info.trajectories += h[p].trajectories;
      ctx->info.particles += h[p].particles;
    }
    for (int p = 0; p < n_points; ++p) {
      memcpy(out[p].hist_n, h[p].hist_n, sizeof out[p].hist_n);
      memcpy(out[p].hist_traj_time, h[p].hist_traj_time, sizeof out[p].hist_traj_time);
      memcpy(out[p].hist_meet_time, h[p].hist_meet_time, sizeof out[p].hist_meet_time);
      memcpy(out[p].done, h[p].done, sizeof out[p].done);
      out[p].unmet_particles = h[p].unmet_particles;
      out[p].particles = h[p].particles;
      out[p].trajectories = h[p].trajectories;
      out[p].sum_traj_time = (double)h[p].sum_traj_time_q10 / 1024.0;
      out[p].info = ctx->info;  // work and timing of the whole sweep
    }
    if (ctx->h_counters->unsupported)
      rc = fail(HEROAM_E_UNSUPPORTED, "%llu trajectories have more than %d particles", ctx->h_counters->unsupported,
                MAX_ACTIVE);
  }
  return rc;
}


extern "C" int heroam_gpu_run_sweep(heroam_ctx *ctx, unsigned long long he_number, int n_points,
                                    const heroam_sweep_point *points, heroam_stats *out) {
  if (!ctx || !out || !points) return fail(HEROAM_E_ARG, "heroam_gpu_run_sweep: null argument");
  if (n_points < 1 || n_points > MAX_SWEEP_POINTS)
    return fail(HEROAM_E_ARG, "heroam_gpu_run_sweep: %d points (1..%d supported per call)", n_points, MAX_SWEEP_POINTS);
  CU(cudaSetDevice(ctx->device));
  memset(out, 0, (size_t)n_points * sizeof *out);
  memset(&ctx->info, 0, sizeof ctx->info);
  ctx->uploaded = false;
  Consts c;
  int rc = make_consts(ctx->conf, he_number, &c);
  if (rc) return rc;
  PoolParams pool;
  memset(&pool, 0, sizeof pool);
  double lambda_max = 0.0;
  unsigned long long count = 0;
  for (int p = 0; p < n_points; ++p) {
    const double lambda = point_lambda(ctx->conf, he_number, points[p].intensity, points[p].pulse_width);
    if (!(lambda > 0.0) || lambda > 700.0)
      return fail(HEROAM_E_ARG, "mean excitation number %g outside (0, 700]: Knuth's product of uniforms "
                  "(simulation.c:257-275) is undefined there", lambda);
    if (lambda > lambda_max) lambda_max = lambda;
    count += points[p].count;
    pool.point_end[p] = count;
    pool.point_first[p] = points[p].first_index;
    pool.point_limit[p] = exp(-1.0 * lambda);
  }
  if (count == 0) return HEROAM_OK;
  const uint32_t cap = pool_slot_capacity(lambda_max);
  const unsigned long long pool_max = ctx->opt.pool_slots > 0 ? (unsigned long long)ctx->opt.pool_slots : (1ull << 22);
  const uint32_t slots = (uint32_t)(count < pool_max ? count : pool_max);
  if ((unsigned long long)slots * cap > 0xfffffff0ull)  // record offsets are 32-bit (TrajMeta::first)
    return fail(HEROAM_E_ARG, "pool_slots = %u with %u records per slot exceeds 2^32 particle records", slots, cap);
  rc = reserve(ctx, slots, (size_t)slots * cap);
  if (rc) return rc;
  ctx->he_number = he_number;
  ctx->n_traj = slots;
  ctx->n_particles = (size_t)slots * cap;
  DevBuf<DevStats> stats_buf;
  DevBuf<unsigned long long> cursor_buf;
  CU(stats_buf.alloc((size_t)n_points));
  CU(cursor_buf.alloc(1));
  DevStats *d_stats = stats_buf.ptr;
  unsigned long long *d_cursor = cursor_buf.ptr;
  pool.sp.seed = (long long)ctx->conf.seed;
  pool.sp.he_number = he_number;
  pool.sp.first_index = 0;
  pool.sp.limit = pool.point_limit[0];
  pool.sp.radius_init = 2.22 * pow((double)he_number, 1.0 / 3.0);
  pool.sp.sigma = ctx->conf.velocity;
  pool.cursor = d_cursor;
  pool.count = count;
  pool.cap = cap;
  pool.n_points = (uint32_t)n_points;
  pool.stats = d_stats;
  CU(cudaEventRecord(ctx->ev_start, ctx->main));
  CU(cudaMemsetAsync(d_stats, 0, (size_t)n_points * sizeof(DevStats), ctx->main));
  CU(cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long), ctx->main));
  CU(cudaMemsetAsync(ctx->d_counters, 0, sizeof(Counters), ctx->main));
  rc = run_pool_mode(ctx, c, pool, slots, PoolParts());
  return collect_pool_stats(ctx, rc, d_stats, n_points, out);
}

extern "C" int heroam_gpu_run_stats(heroam_ctx *ctx, unsigned long long he_number, unsigned long long first_index,
                                    unsigned long long count, heroam_stats *out) {
  if (!ctx || !out) return fail(HEROAM_E_ARG, "heroam_gpu_run_stats: null argument");
  heroam_sweep_point point;
  point.intensity = ctx->conf.intensity;
  point.pulse_width = ctx->conf.pulse_width;
  point.first_index = first_index;
  point.count = count;
  return heroam_gpu_run_sweep(ctx, he_number, 1, &point, out);
}

// ---------------------------------------------------------------------------------
// The reference's own sweep, over droplet sizes (main.c:158-164), as ONE run of the pool: every size has its own
// sphere radius, hence its own constants and its own kappa-scaled table copy, so the pool is partitioned -- each
// size owns an equal share of the slots and its own bin lists (PartParams, PartSlice) -- and all sizes share the
// passes: one drain for the whole sweep instead of one per size.
// ---------------------------------------------------------------------------------
extern "C" int heroam_gpu_run_size_sweep(heroam_ctx *ctx, int n_points, const heroam_size_point *points, heroam_stats *out) {
  if (!ctx || !out || !points) return fail(HEROAM_E_ARG, "heroam_gpu_run_size_sweep: null argument");
  if (n_points < 1 || n_points > MAX_PARTS)
    return fail(HEROAM_E_ARG, "heroam_gpu_run_size_sweep: %d sizes (1..%d supported per call)", n_points, MAX_PARTS);
  if (n_points == 1) return heroam_gpu_run_stats(ctx, points[0].he_number, points[0].first_index, points[0].count, out);
  CU(cudaSetDevice(ctx->device));
  memset(out, 0, (size_t)n_points * sizeof *out);
  memset(&ctx->info, 0, sizeof ctx->info);
  ctx->uploaded = false;
  PoolParts parts;
  parts.host.resize((size_t)n_points);
  double lambda_max = 0.0;
  unsigned long long count_max = 0, total = 0;
  for (int p = 0; p < n_points; ++p) {
    PartParams &pp = parts.host[p];
    int rc = make_consts(ctx->conf, points[p].he_number, &pp.c);
    if (rc) return rc;
    pp.c.f_max = ctx->f_max;
    const double lambda = host_lambda(ctx->conf, points[p].he_number);
    if (!(lambda > 0.0) || lambda > 700.0)
      return fail(HEROAM_E_ARG, "mean excitation number %g outside (0, 700]: Knuth's product of uniforms "
                  "(simulation.c:257-275) is undefined there", lambda);
    if (lambda > lambda_max) lambda_max = lambda;
    pp.he_number = points[p].he_number;
    pp.first_index = points[p].first_index;
    pp.count = points[p].count;
    pp.limit = exp(-1.0 * lambda);
    pp.radius_init = 2.22 * pow((double)points[p].he_number, 1.0 / 3.0);
    pp.table_k = nullptr;
    if (pp.count > count_max) count_max = pp.count;
    total += pp.count;
  }
  if (total == 0) return HEROAM_OK;
  const uint32_t cap = pool_slot_capacity(lambda_max);
  const unsigned long long pool_max = ctx->opt.pool_slots > 0 ? (unsigned long long)ctx->opt.pool_slots : (1ull << 22);
  // equal shares of the pool, whole blocks of the resolve kernel, no more than the largest part needs
  unsigned long long per_part = pool_max / (unsigned long long)n_points;
  if (per_part > count_max) per_part = count_max;
  per_part = (per_part + 127ull) / 128ull * 128ull;
  const uint32_t slots = (uint32_t)(per_part * (unsigned long long)n_points);
  if ((unsigned long long)slots * cap > 0xfffffff0ull)
    return fail(HEROAM_E_ARG, "%u pool slots with %u records per slot exceed 2^32 particle records", slots, cap);
  int rc = reserve(ctx, slots, (size_t)slots * cap);
  if (rc) return rc;
  ctx->he_number = points[0].he_number;
  ctx->n_traj = slots;
  ctx->n_particles = (size_t)slots * cap;
  // one kappa-scaled copy of the table per size (Turbo kernels)
  const unsigned table_n = (unsigned)ctx->table_len + 1u;
  DevBuf<float> scaled;
  CU(scaled.alloc((size_t)n_points * table_n));
  for (int p = 0; p < n_points; ++p) {
    float *dst = scaled.ptr + (size_t)p * table_n;
    k_scale_table<<<(table_n + 255u) / 256u, 256, 0, ctx->main>>>(dst, ctx->d_table, table_n, parts.host[p].c.kappa);
    CU(cudaGetLastError());
    ctx->info.kernel_launches++;
    parts.host[p].table_k = dst;
  }
  DevBuf<PartParams> d_parts;
  DevBuf<DevStats> stats_buf;
  DevBuf<unsigned long long> cursor_buf;
  CU(d_parts.alloc((size_t)n_points));
  CU(stats_buf.alloc((size_t)n_points));
  CU(cursor_buf.alloc((size_t)n_points));
  CU(cudaMemcpyAsync(d_parts.ptr, parts.host.data(), (size_t)n_points * sizeof(PartParams), cudaMemcpyHostToDevice, ctx->main));
  PoolParams pool;
  memset(&pool, 0, sizeof pool);
  pool.sp.seed = (long long)ctx->conf.seed;
  pool.sp.he_number = points[0].he_number;
  pool.sp.first_index = 0;
  pool.sp.limit = parts.host[0].limit;
  pool.sp.radius_init = parts.host[0].radius_init;
  pool.sp.sigma = ctx->conf.velocity;
  pool.cursor = cursor_buf.ptr;
  pool.count = total;
  pool.cap = cap;
  pool.n_points = (uint32_t)n_points;
  pool.stats = stats_buf.ptr;
  pool.n_parts = (uint32_t)n_points;
  pool.slots_per_part = (uint32_t)per_part;
  pool.parts = d_parts.ptr;
  CU(cudaEventRecord(ctx->ev_start, ctx->main));
  CU(cudaMemsetAsync(stats_buf.ptr, 0, (size_t)n_points * sizeof(DevStats), ctx->main));
  CU(cudaMemsetAsync(cursor_buf.ptr, 0, (size_t)n_points * sizeof(unsigned long long), ctx->main));
  CU(cudaMemsetAsync(ctx->d_counters, 0, sizeof(Counters), ctx->main));
  rc = run_pool_mode(ctx, parts.host[0].c, pool, slots, parts);
  rc = collect_pool_stats(ctx, rc, stats_buf.ptr, n_points, out);
  CU(cudaStreamSynchronize(ctx->main));  // the scaled tables and part records go out of scope here
  return rc;
}
