/*
 * include/heroam_gpu.h -- C ABI of libheroam_gpu.so, the B200 (sm_100a) engine for
 * the Monte Carlo hot path of s-kesh/Sphere_roaming.
 *
 * The reference has no plugin registry; its seam is the C interface of
 * src/simulation.h, called from exactly one place, simulate_for_number
 * (src/main.c:50-97).  The entry points below are the batch forms of those
 * calls.  Every function returns 0 on success and a non-zero HEROAM_E_* code on
 * failure; heroam_gpu_last_error() then holds a message.  There is NO CPU
 * fallback: without a usable CUDA device heroam_gpu_create fails.
 *
 *   reference call (file:line)                         replaced by
 *   -------------------------------------------------  ------------------------------
 *   initialize_particles      simulation.h:76-79,      heroam_gpu_sample_batch /
 *                             simulation.c:277-337       heroam_gpu_sample_flat
 *   #pragma omp parallel for                           heroam_gpu_simulate_batch /
 *     simulate_particles      main.c:76-78,              heroam_gpu_simulate_flat
 *                             simulation.h:94-97,
 *                             simulation.c:438-526
 *   free_particles            simulation.h:85,         heroam_gpu_free_particles
 *                             simulation.c:339-341
 *   saveflag = 1: save_timestep_to_hdf5 per iteration  heroam_gpu_simulate_traced
 *                             simulation.c:460-476,      (+ host/heroam_h5.h for the file)
 *                             h5file.c:88-112
 *   (none: the reference leaves histogramming to       heroam_gpu_run_stats
 *    the user of particlefile_after_<N>)
 */
#ifndef HEROAM_GPU_H_
#define HEROAM_GPU_H_

#include "heroam_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#define HEROAM_ABI_VERSION 2

enum {
  HEROAM_OK = 0,
  HEROAM_E_CUDA = 1,      /* CUDA runtime / driver failure (no device, launch error, OOM)   */
  HEROAM_E_ARG = 2,       /* invalid argument or configuration                              */
  HEROAM_E_UNSUPPORTED = 3, /* a trajectory has more particles than the kernels support      */
  HEROAM_E_CLOCK = 4      /* dt/max_time such that the reference's float clock would stall
                             and its loop never end                          */
};

/* Arithmetic mode of the integrator. */
enum {
  /* Every float operation of the reference in the reference's order, no FMA
   * contraction, correctly rounded sqrt/divide, the one double-precision expression
   * (simulation.c:220) in double: agrees bit-for-bit with the reference compiled
   * `-O2 -ffp-contract=off`. */
  HEROAM_MODE_EXACT = 0,
  /* Same integrator and state, algebraically simplified (closed-form q (0,0,1) q^-1,
   * dead zero-terms dropped), FMA contraction, hardware rsqrt: the counterpart of the
   * reference's own `-O3 -ffast-math -march=native` release build.  Agrees with the
   * strict reference within the tolerance stated in tests/test_gpu_parity.py. */
  HEROAM_MODE_FAST = 1,
  /* The fast arithmetic without the folded constants of the hot kernels:
   * same tolerance class, ~30 % more instructions; kept as a cross-check of HEROAM_MODE_FAST. */
  HEROAM_MODE_FAST_PLAIN = 2,
  /* HEROAM_MODE_FAST with the orientation renormalised after every 8th step of a trajectory
   * instead of every step (the direction of q, which is all the position depends on, is the
   * same either way in exact arithmetic).  ~14 % faster, but the per-step roundings no longer
   * mirror the reference's one for one: positions agree with the strict reference to ~2e-3 A
   * (median, after 2e5 steps) instead of ~2e-5 A.  Exit times and success patterns meet the
   * same bounds as HEROAM_MODE_FAST.  Opt-in. */
  HEROAM_MODE_FAST_RELAXED = 3,
  /* HEROAM_MODE_FAST with every free flight taken as ONE rotation instead of step by step (without
   * force a particle turns rigidly about its fixed axis: the exact solution of the map the reference
   * iterates).  Three quarters of all particle-steps of the repository configuration are free flight,
   * so this is about twice as fast again -- but the per-step roundings of the reference are not
   * reproduced during flights, so positions agree with the strict reference only to the size of its own
   * rounding walk, and results depend (in the last bits, not statistically) on how a run is cut into
   * passes: a trajectory is reproducible for the same run, not across batch or pool sizes.  Opt-in;
   * its own tolerance row in tests/horizon_common.py. */
  HEROAM_MODE_FAST_CLOSED = 4
};

/* How a trajectory left the step loop (heroam_traj_result.status). */
enum {
  HEROAM_DONE_SUCCESS = 0,  /* find_success criterion met (simulation.c:33-46, :519)       */
  HEROAM_DONE_MAXTIME = 1,  /* time >= max_time (simulation.c:469)                         */
  HEROAM_DONE_DEADLOCK = 2, /* no particle left to move, criterion unmet: the reference
                               would never return; defined exit              */
  HEROAM_DONE_STEPCAP = 4,  /* heroam_options.max_steps reached                            */
  HEROAM_DONE_UNSUPPORTED = 6 /* more than 128 particles still moving: not integrated; the
                               call returns HEROAM_E_UNSUPPORTED                            */
};

typedef struct heroam_options {
  int mode;               /* HEROAM_MODE_*                                                 */
  unsigned int segment_steps; /* steps a trajectory runs between re-binning passes (0 = default) */
  unsigned long long max_steps; /* per-trajectory cap on loop iterations (0 = none)         */
  int wide_kernel;        /* kernels for > 8 moving particles.  0 (default): register/shared-memory
                             hybrid for 9..12, shared-memory kernel for 13..16, warp-per-trajectory
                             above; 1 = the slow record-resident kernel (cross-check only);
                             2 = warp-per-trajectory for everything above 8 (cross-check);
                             3 = shared-memory kernel for 9..16 (cross-check)                  */
  unsigned int pool_slots; /* trajectories kept in flight by heroam_gpu_run_stats (0 = 2^22)        */
  int count_work;         /* fast mode only: 1 = the integrate kernels also count in-range pair-steps
                             (heroam_run_info.inrange_steps) at ~3 % cost; exact mode always counts   */
  int no_flight;          /* cross-check only: 1 = no free flight and no splitting, every moving particle is
                             integrated in full every step (results must not change)                  */
} heroam_options;

typedef struct heroam_traj_result {
  uint32_t steps;         /* iterations of the step loop                                   */
  uint32_t status;        /* HEROAM_DONE_*                                                 */
  uint32_t n_flagged;     /* particles with success == 1 at exit                           */
  float time;             /* the loop's time variable at exit [fs]                         */
} heroam_traj_result;

/* Work and timing of the last simulate/run call on a context. */
typedef struct heroam_run_info {
  unsigned long long trajectories;
  unsigned long long particles;
  unsigned long long particle_steps; /* sum over steps of particles advanced               */
  unsigned long long pair_steps;     /* sum over steps of active pairs a(a-1)/2            */
  unsigned long long inrange_steps;  /* ... of which hit the force table                   */
  unsigned long long passes;         /* re-binning passes                                  */
  unsigned long long kernel_launches;/* kernels of this library launched                   */
  double device_ms;       /* CUDA-event time, first kernel to last kernel (state resident) */
  double integrate_ms;    /* sum of CUDA-event times of the integrate kernels alone        */
  double h2d_ms, d2h_ms;  /* CUDA-event times of the copies (0 for resident runs)          */
  unsigned long long h2d_bytes, d2h_bytes;
  unsigned long long pair_skipped;   /* pair-steps of free-flight segments: part of pair_steps, but
                                        provably force-free and never evaluated               */
  unsigned long long free_steps;     /* particle-steps taken in free flight (part of particle_steps) */
  double resolve_ms;      /* sum of CUDA-event times from the end of a pass's integrate kernels to the
                             bin counts of the next being on the host (resolve kernel + copy)          */
} heroam_run_info;

#define HEROAM_HIST_N 256
#define HEROAM_HIST_T 4096
typedef struct heroam_stats {
  unsigned long long trajectories;
  unsigned long long particles;
  unsigned long long hist_n[HEROAM_HIST_N];   /* excitation count per trajectory (clamped) */
  unsigned long long hist_traj_time[HEROAM_HIST_T]; /* loop time at exit, bins of max_time/4096 */
  unsigned long long hist_meet_time[HEROAM_HIST_T]; /* per frozen particle: its time        */
  unsigned long long done[8];                 /* trajectories by HEROAM_DONE_* code        */
  unsigned long long unmet_particles;         /* particles still unpaired at exit          */
  double sum_traj_time;                       /* sum of loop times at exit [fs], accumulated on the
                                                 device as an integer count of 2^-10 fs: the same
                                                 bits on every run                            */
  heroam_run_info info;
} heroam_stats;

typedef struct heroam_ctx heroam_ctx;

/* Library / device discovery.  heroam_gpu_device_count returns 0 without a GPU. */
int heroam_gpu_abi_version(void);
int heroam_gpu_device_count(void);
const char *heroam_gpu_last_error(void);

/* Context: one per GPU.  Uploads the force table (conf->force_grid_length floats,
 * main.c:130) once; conf and force_list are copied, the caller may free them.
 * conf->seed == 0 means "seed from the clock", as in the reference (simulation.c:283-286):
 * it is resolved to time(NULL) here, once per context -- a caller that drives several
 * contexts as one run should resolve it itself and pass the same non-zero seed to all. */
int heroam_gpu_create(const heroam_config *conf, const float *force_list, int device,
                      const heroam_options *opt /* NULL = defaults (exact mode) */,
                      heroam_ctx **out);
void heroam_gpu_destroy(heroam_ctx *ctx);
int heroam_gpu_set_mode(heroam_ctx *ctx, int mode);

/* Drop-in for the loop at main.c:76-78.  Consumes host initial states in the
 * reference's own Particles layout (from initialize_particles or from
 * heroam_gpu_sample_batch) and writes the final states back in place; afterwards
 * save_particles / particlefile_after_<N> work unchanged.  Particle.time is taken
 * as 0 on entry (the reference leaves it uninitialised).
 * results may be NULL. */
int heroam_gpu_simulate_batch(heroam_ctx *ctx, unsigned long long he_number,
                              heroam_particles *trajectories, int count,
                              heroam_traj_result *results);

/* Optional: size the device buffers and the page-locked staging of heroam_gpu_simulate_batch for batches of up
 * to n_trajectories trajectories / n_records particle records ahead of time, so that the first large call does
 * not pay for the allocations (they otherwise grow on demand and are kept).  The reference allocates per call
 * (main.c:55, simulation.c:310). */
int heroam_gpu_reserve(heroam_ctx *ctx, unsigned long long n_trajectories, unsigned long long n_records);

/* Same on a flat batch: counts[count] particles per trajectory, records stored
 * trajectory after trajectory in flat[]. */
int heroam_gpu_simulate_flat(heroam_ctx *ctx, unsigned long long he_number,
                             const uint32_t *counts, heroam_particle *flat, int count,
                             heroam_traj_result *results);

/* The saveflag = 1 path of simulate_particles (simulation.c:460-476, :522-523): as
 * heroam_gpu_simulate_flat, and in addition the sink is called once per trajectory and loop
 * iteration with the records the reference hands to save_timestep_to_hdf5 at the top of that
 * iteration (all `no` particles, frozen ones included) -- what ends up in
 * ./results/trajectory_%06d.h5 as /Step_<step>/particles (h5file.c:88-112; host/heroam_h5.h
 * writes that file).  Calls for one trajectory come in increasing step order; `trajectory` is
 * the index within the batch; `records` is only valid during the call.  A non-zero return
 * from the sink aborts the run.  The free-flight and splitting shortcuts are off on this
 * path (every record must be current at every step), so it runs at the speed of the I/O it
 * feeds, not of the histogram path. */
typedef int (*heroam_step_sink)(void *user, uint32_t trajectory, uint32_t step,
                                const heroam_particle *records, uint32_t no);
int heroam_gpu_simulate_traced(heroam_ctx *ctx, unsigned long long he_number,
                               const uint32_t *counts, heroam_particle *flat, int count,
                               heroam_traj_result *results, heroam_step_sink sink, void *user);

/* The same work split in three so that a benchmark can time the resident part:
 * upload copies the batch to the device, run integrates it there (may be called
 * once per upload), download copies the final states (and results) back. */
int heroam_gpu_upload(heroam_ctx *ctx, unsigned long long he_number, const uint32_t *counts,
                      const heroam_particle *flat, int count);
int heroam_gpu_run(heroam_ctx *ctx);
int heroam_gpu_download(heroam_ctx *ctx, heroam_particle *flat, heroam_traj_result *results);

/* Philox replacement for initialize_particles (simulation.c:277-337): trajectory i of
 * the call is trajectory (first_index + i) of the stream keyed by (conf->seed,
 * he_number), so results do not depend on how a run is sharded.
 *   _batch: fills out[i].no/.index and mallocs out[i].particle (release with
 *           heroam_gpu_free_particles, or the reference's free_particles).
 *   _flat : counts[count] and up to capacity records; *total receives the number of
 *           records needed (call again with a larger buffer if it exceeds capacity). */
int heroam_gpu_sample_batch(heroam_ctx *ctx, unsigned long long he_number,
                            unsigned long long first_index, int count, heroam_particles *out);
int heroam_gpu_sample_flat(heroam_ctx *ctx, unsigned long long he_number,
                           unsigned long long first_index, int count, uint32_t *counts,
                           heroam_particle *flat, long capacity, long *total);
void heroam_gpu_free_particles(heroam_particles *p);

/* Fused sample -> integrate -> histogram for large runs: nothing but counters and
 * histograms comes back.  Works through [first_index, first_index + count) in
 * device-resident batches. */
int heroam_gpu_run_stats(heroam_ctx *ctx, unsigned long long he_number,
                         unsigned long long first_index, unsigned long long count,
                         heroam_stats *out);

/* A laser sweep (BASELINE config 4: intensity x pulse width at one droplet size) as ONE run of
 * the streaming pool: the points' trajectories are concatenated into one index space, so the
 * drain of one point (its few trajectories that run to max_time) overlaps the bulk of the
 * next ones instead of idling the GPU once per point.  Point p uses conf with intensity and
 * pulse_width replaced (they only enter the mean excitation number, simulation.c:289-293) and
 * the Philox indices [first_index, first_index + count): out[p] is exactly what
 * heroam_gpu_run_stats would return for that point alone, except out[p].info, which describes
 * the whole sweep.  At most 32 points per call.  (The reference's own sweep is over droplet
 * sizes, main.c:158-164: heroam_gpu_run_size_sweep below.) */
typedef struct heroam_sweep_point {
  float intensity;                /* [GW/cm^2] */
  float pulse_width;              /* [fs]      */
  unsigned long long first_index; /* first Philox trajectory index of this point (sharding) */
  unsigned long long count;       /* trajectories of this point on this GPU */
} heroam_sweep_point;
int heroam_gpu_run_sweep(heroam_ctx *ctx, unsigned long long he_number, int n_points,
                         const heroam_sweep_point *points, heroam_stats *out /* [n_points] */);

/* The reference's own sweep -- over droplet sizes, main.c:158-164: `for (N = start; N < stop; N += step)
 * simulate_for_number(N, ...)` -- as ONE run of the streaming pool.  Every size has its own sphere radius
 * (simulation.c:297, :448), so the pool is partitioned: each size owns an equal share of the slots, its own
 * constants and its own scaled copy of the force table, and all sizes share the passes -- one latency-bound
 * end of run for the whole sweep instead of one per size.  out[p] is exactly what heroam_gpu_run_stats would
 * return for size p alone (same Philox indices), except out[p].info, which describes the whole sweep.
 * At most 32 sizes per call. */
typedef struct heroam_size_point {
  unsigned long long he_number;   /* helium atoms in the droplet */
  unsigned long long first_index; /* first Philox trajectory index of this size (sharding) */
  unsigned long long count;       /* trajectories of this size on this GPU */
} heroam_size_point;
int heroam_gpu_run_size_sweep(heroam_ctx *ctx, int n_points, const heroam_size_point *points,
                              heroam_stats *out /* [n_points] */);

int heroam_gpu_last_run_info(const heroam_ctx *ctx, heroam_run_info *out);

/* Register-resident FFMA microbenchmark on the context's device: the measured FP32
 * peak used as the roofline denominator.  TFLOP/s. */
int heroam_gpu_fp32_peak(heroam_ctx *ctx, double *tflops, double *sm_clock_mhz);
/* The same with both the multiplier and the addend in registers (three register operands per
 * FFMA, as in the integrate kernels) instead of the constant bank. */
int heroam_gpu_fp32_peak_rrr(heroam_ctx *ctx, double *tflops);
/* The same chains as packed FFMA2 (two FP32 lanes per 64-bit register pair, three register operands). */
int heroam_gpu_fp32_peak_x2(heroam_ctx *ctx, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* HEROAM_GPU_H_ */
