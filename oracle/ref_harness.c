/*
 * oracle/ref_harness.c  --  TEST INFRASTRUCTURE, not product code.
 *
 * Thin C harness around the UNMODIFIED reference sources.  The reference's
 * simulation.c is pulled into this translation unit by path (the Makefile
 * passes -DREF_SIMULATION_C='"/root/reference/src/simulation.c"'), so that its
 * `static` helpers (calculate_force, update_orientation, ... reference
 * simulation.c:33,:158,:238,:257,:343,:354,:394) are reachable for
 * known-answer tests; vector.c is compiled by path as a separate object.
 * Nothing from the reference is copied into this repository.
 *
 * The three HDF5 entry points the trajectory loop calls when saveflag=1
 * (reference simulation.c:463,:476,:523) are defined here as no-ops: libhdf5
 * does not exist in this image.  They are defined here: no-ops when saveflag=0, and with
 * saveflag=1 (ref_simulate_traced) save_timestep_to_hdf5 copies what it is given into a
 * caller-supplied buffer, which is how the per-step golden states are obtained.
 *
 * Everything exported is prefixed ref_ and uses flat arrays so that Python
 * (ctypes, tests/ only) and bench.py's cpu_baseline leg can drive it.
 *
 * Deviation from the reference that the harness applies on purpose:
 *   Particle.time is never initialised by initialize_particles
 *   (simulation.c:310-331, first use is `+=` at :507).  ref_initialize zeroes
 *   it so that results are defined.
 */
#include <setjmp.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include REF_SIMULATION_C

/* ---- HDF5 no-ops (declared in the reference's h5file.h) ---- */
/* When a capture buffer is armed (ref_simulate_traced below) save_timestep_to_hdf5 -- called by
 * the reference itself at the top of every loop iteration when saveflag=1, simulation.c:475-476
 * -- appends the records it is handed: exactly what the reference would put into
 * /Step_<step>/particles. */
static _Thread_local Particle *g_trace_out = NULL;
static _Thread_local unsigned long g_trace_cap = 0, g_trace_steps = 0;
/* Guard against the reference's one non-terminating case: once every particle is frozen
 * but find_success's `==` criterion is unmet, no clock advances any more and `while (!sucess && time <
 * max_time)` (simulation.c:469) spins forever.  ref_simulate_guarded runs the reference with saveflag = 1,
 * so that it calls save_timestep_to_hdf5 (below) at the top of every iteration; there the state is
 * inspected, and on that condition the harness steps out of simulate_particles with longjmp.  The
 * records are then exactly what every further iteration would leave them as. */
static _Thread_local int g_guard = 0;
static _Thread_local jmp_buf g_guard_jmp;
hid_t initialize_file(const char *filename) { (void)filename; return 0; }
hid_t open_file(const char *filename) { (void)filename; return 0; }
herr_t save_timestep_to_hdf5(const hid_t file_id, const unsigned int step,
                             const Particles *particles) {
  (void)file_id;
  if (g_guard && step > 0) {
    unsigned int i = 0;
    while (i < particles->no && particles->particle[i].success) ++i;
    if (i == particles->no && !find_success(particles)) longjmp(g_guard_jmp, 1);
  }
  if (g_trace_out) {
    if (step != g_trace_steps) return -1; /* the reference numbers the groups 0, 1, 2, ... */
    if (g_trace_steps < g_trace_cap)
      memcpy(g_trace_out + (size_t)g_trace_steps * particles->no, particles->particle,
             (size_t)particles->no * sizeof(Particle));
    g_trace_steps++;
  }
  return 0;
}
herr_t read_timestep_from_hdf5(const hid_t file_id, const unsigned int step,
                               Particles *particles) {
  (void)file_id; (void)step; (void)particles; return -1;
}
void close_file(const hid_t file_id) { (void)file_id; }

/* ---- layout probes: the repo's own headers are checked against these ---- */
int ref_sizeof_particle(void) { return (int)sizeof(Particle); }
int ref_sizeof_particles(void) { return (int)sizeof(Particles); }
int ref_sizeof_config(void) { return (int)sizeof(struct config); }
int ref_particle_offsets(int *out) {
  out[0] = (int)offsetof(Particle, index);
  out[1] = (int)offsetof(Particle, success);
  out[2] = (int)offsetof(Particle, time);
  out[3] = (int)offsetof(Particle, force_mag);
  out[4] = (int)offsetof(Particle, velocity_sq);
  out[5] = (int)offsetof(Particle, orientation);
  out[6] = (int)offsetof(Particle, position);
  out[7] = (int)offsetof(Particle, velocity);
  out[8] = (int)offsetof(Particle, ang_velocity);
  out[9] = (int)offsetof(Particle, force);
  return 10;
}
int ref_config_offsets(int *out) {
  out[0] = (int)offsetof(struct config, print);
  out[1] = (int)offsetof(struct config, hv);
  out[2] = (int)offsetof(struct config, intensity);
  out[3] = (int)offsetof(struct config, cross_section);
  out[4] = (int)offsetof(struct config, pulse_width);
  out[5] = (int)offsetof(struct config, number);
  out[6] = (int)offsetof(struct config, saveflag);
  out[7] = (int)offsetof(struct config, start);
  out[8] = (int)offsetof(struct config, step);
  out[9] = (int)offsetof(struct config, stop);
  out[10] = (int)offsetof(struct config, helium_number);
  out[11] = (int)offsetof(struct config, seed);
  out[12] = (int)offsetof(struct config, mass);
  out[13] = (int)offsetof(struct config, velocity);
  out[14] = (int)offsetof(struct config, dt);
  out[15] = (int)offsetof(struct config, max_time);
  out[16] = (int)offsetof(struct config, icd_dist);
  out[17] = (int)offsetof(struct config, force_grid);
  out[18] = (int)offsetof(struct config, force_grid_start);
  out[19] = (int)offsetof(struct config, force_grid_length);
  out[20] = (int)offsetof(struct config, force_file);
  return 21;
}
int ref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---- config file parser of the reference (simulation.c:48-156) ---- */
int ref_read_config(const char *filename, struct config *conf) {
  return read_config(filename, conf);
}

/* ---- sampling stage: initialize_particles (simulation.c:277-337) ----
 * Runs the reference sampler for conf->number trajectories, zeroes
 * Particle.time (Q1), and flattens the result.
 *   counts[number]      <- particles per trajectory
 *   flat[capacity]      <- Particle records, trajectory after trajectory
 * Returns the total number of particles, or -(needed) if capacity is short. */
long ref_initialize(unsigned long long he_number, const struct config *conf,
                    const float *force_list, unsigned int *counts,
                    Particle *flat, long capacity) {
  const int number = conf->number;
  Particles *all = (Particles *)malloc((size_t)number * sizeof(Particles));
  initialize_particles(he_number, conf, force_list, all);
  long total = 0;
  for (int i = 0; i < number; ++i) total += all[i].no;
  int ok = (total <= capacity);
  long at = 0;
  for (int i = 0; i < number; ++i) {
    counts[i] = all[i].no;
    if (ok) {
      for (unsigned int j = 0; j < all[i].no; ++j) {
        all[i].particle[j].time = 0.0f;
        flat[at + j] = all[i].particle[j];
      }
    }
    at += all[i].no;
    free_particles(all + i);
  }
  free(all);
  return ok ? total : -total;
}

/* ---- integration stage: simulate_particles (simulation.c:438-526) under the
 * same OpenMP loop shape as the reference driver (main.c:76-78).
 *   schedule_dynamic = 0 : the reference's as-written static schedule
 *   schedule_dynamic = 1 : schedule(dynamic,1)
 * `flat` is advanced in place.  nthreads <= 0 keeps the OpenMP default. */
void ref_simulate(unsigned long long he_number, const struct config *conf,
                  const float *force_list, int number,
                  const unsigned int *counts, Particle *flat, int nthreads,
                  int schedule_dynamic) {
  Particles *all = (Particles *)malloc((size_t)number * sizeof(Particles));
  long at = 0;
  for (int i = 0; i < number; ++i) {
    all[i].no = counts[i];
    all[i].index = (unsigned int)i;
    all[i].particle = flat + at;
    at += counts[i];
  }
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
  if (schedule_dynamic) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < number; ++i)
      simulate_particles(conf, he_number, force_list, all + i);
  } else {
#pragma omp parallel for
    for (int i = 0; i < number; ++i)
      simulate_particles(conf, he_number, force_list, all + i);
  }
  free(all);
}

/* The same loop with the guard armed (see g_guard): deadlocked[i] = 1 for the trajectories the
 * reference itself would never return from.  The only difference from ref_simulate is saveflag = 1
 * (one call of the no-op save_timestep_to_hdf5 per iteration). */
void ref_simulate_guarded(unsigned long long he_number, const struct config *conf,
                          const float *force_list, int number,
                          const unsigned int *counts, Particle *flat, int nthreads,
                          int schedule_dynamic, unsigned char *deadlocked) {
  Particles *all = (Particles *)malloc((size_t)number * sizeof(Particles));
  long at = 0;
  for (int i = 0; i < number; ++i) {
    all[i].no = counts[i];
    all[i].index = (unsigned int)i;
    all[i].particle = flat + at;
    at += counts[i];
    deadlocked[i] = 0;
  }
  struct config c = *conf;
  c.saveflag = 1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
  if (schedule_dynamic) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < number; ++i) {
      g_guard = 1;
      if (setjmp(g_guard_jmp) == 0) simulate_particles(&c, he_number, force_list, all + i);
      else deadlocked[i] = 1;
      g_guard = 0;
    }
  } else {
#pragma omp parallel for
    for (int i = 0; i < number; ++i) {
      g_guard = 1;
      if (setjmp(g_guard_jmp) == 0) simulate_particles(&c, he_number, force_list, all + i);
      else deadlocked[i] = 1;
      g_guard = 0;
    }
  }
  free(all);
}

/* ---- one trajectory with saveflag=1: every state the reference would write to its HDF5 file.
 * out receives up to cap_steps snapshots of `no` records each; returns the number of
 * save_timestep_to_hdf5 calls (= loop iterations).  `p` is advanced in place. */
unsigned long ref_simulate_traced(unsigned long long he_number, const struct config *conf,
                                  const float *force_list, unsigned int no, Particle *p,
                                  unsigned long cap_steps, Particle *out) {
  struct config c = *conf;
  c.saveflag = 1;
  Particles ps;
  ps.no = no;
  ps.index = 0;
  ps.particle = p;
  g_trace_out = out;
  g_trace_cap = cap_steps;
  g_trace_steps = 0;
  simulate_particles(&c, he_number, force_list, &ps);
  g_trace_out = NULL;
  return g_trace_steps;
}

/* ---- known-answer wrappers for the file-static helpers ---- */
void ref_calculate_force(unsigned int no, Particle *p, const struct config *conf,
                         const float *force_list) {
  Particles ps;
  ps.no = no;
  ps.index = 0;
  ps.particle = p;
  calculate_force(&ps, conf, force_list);
}
int ref_find_success(unsigned int no, Particle *p) {
  Particles ps;
  ps.no = no;
  ps.index = 0;
  ps.particle = p;
  return find_success(&ps);
}
void ref_find_accelration(float mass, float radius, const float *pos,
                          const float *force, float *acc) {
  find_accelration(mass, radius, (const Vector3D *)pos, (const Vector3D *)force,
                   (Vector3D *)acc);
}
void ref_update_angvel(float timestep, const float *oldw, const float *acc,
                       float *neww) {
  update_angvel(timestep, (const Vector3D *)oldw, (const Vector3D *)acc,
                (Vector3D *)neww);
}
void ref_update_orientation(float radius, float timestep, float *quat,
                            const float *angvel, float *pos) {
  update_orientation(radius, timestep, (Quat *)quat, (const Vector3D *)angvel,
                     (Vector3D *)pos);
}
/* libc rand() driven samplers; the caller seeds with ref_srand first. */
void ref_srand(unsigned int seed) { srand(seed); }
int ref_rand(void) { return rand(); }
float ref_generate_velocity(const struct config *conf) {
  return generate_velocity(conf);
}
int ref_generate_no_of_excitation(double lambda) {
  return generate_no_of_excitation(lambda);
}
