/*
 * oracle/heroam_oracle.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Plain-C restatement of the reference's Monte Carlo hot path (sampling stage
 * + trajectory integration).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this; the shipped
 * CUDA path never calls into it.
 *
 * Parity status: PINNED.  The reference has no tests or golden vectors of its
 * own, so the restatement is pinned against the
 * reference's own object code: oracle/_ref (reference simulation.c + vector.c
 * compiled unmodified, `-O2 -ffp-contract=off`) must agree bit-for-bit with
 * this file on every KAT and trajectory in tests/test_oracle_vs_ref.py, and
 * against tests/golden/ fixtures generated from oracle/_ref by
 * tests/golden/make_golden.py.
 */
#ifndef HEROAM_ORACLE_H_
#define HEROAM_ORACLE_H_

#include "../include/heroam_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* How a trajectory left the step loop. */
enum {
  ORC_DONE_SUCCESS = 0,  /* find_success criterion met (simulation.c:33-46, :519)          */
  ORC_DONE_MAXTIME = 1,  /* time >= max_time (simulation.c:469)                            */
  ORC_DONE_DEADLOCK = 2, /* nobody left to move but criterion unmet: the reference would
                            spin forever here (time stops advancing); defined exit (Q4)     */
  ORC_DONE_STALLED = 3,  /* float time no longer advances by dt (Q10); defined exit          */
  ORC_DONE_STEPCAP = 4   /* caller-imposed step cap hit (used for k-step state KATs)         */
};

typedef struct orc_result {
  uint32_t steps;     /* iterations of the step loop that were executed */
  uint32_t status;    /* ORC_DONE_*                                     */
  uint32_t n_flagged; /* particles with success == 1 at exit            */
  float time;         /* the loop's `time` variable at exit             */
} orc_result;

/* Work counters for the roofline model of SURVEY.md section 8(d). */
typedef struct orc_counters {
  uint64_t particle_steps; /* sum over steps of active particles advanced          */
  uint64_t pair_steps;     /* sum over steps of active pairs examined in pass 2    */
  uint64_t inrange_steps;  /* ... of which fell inside the force table             */
} orc_counters;

/* calculate_force (simulation.c:158-236) on one trajectory. */
void orc_calculate_force(uint32_t no, heroam_particle *p, const heroam_config *conf,
                         const float *table);

/* simulate_particles (simulation.c:438-526) on one trajectory, in place.
 * max_steps == 0 means "no cap" (run until the reference's own loop would exit). */
orc_result orc_simulate_one(const heroam_config *conf, unsigned long long he_number,
                            const float *table, uint32_t no, heroam_particle *p,
                            uint64_t max_steps, orc_counters *ctr);

/* The driver loop main.c:76-78 over a flat batch. results/ctr may be NULL. */
void orc_simulate(const heroam_config *conf, unsigned long long he_number, const float *table,
                  int number, const uint32_t *counts, heroam_particle *flat,
                  uint64_t max_steps, int nthreads, orc_result *results, orc_counters *ctr);

/* initialize_particles (simulation.c:277-337).  source: 0 = libc rand() seeded with
 * conf->seed exactly as the reference does (bit-comparable with oracle/_ref);
 * 1 = Philox4x32-10 keyed on (seed, he_number, first_index + i) -- the stream the GPU
 * sampler uses.  Writes counts[number] and the flat records; returns the number of
 * particles, or -(needed) if capacity is too small.  Particle.time is set to 0 (Q1). */
long orc_sample(const heroam_config *conf, unsigned long long he_number, const float *table,
                int source, uint64_t first_index, int number, uint32_t *counts,
                heroam_particle *flat, long capacity);

/* Philox4x32-10 block function, exposed for known-answer tests. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* Derived constants, exposed so tests can check the GPU side uses the same ones. */
double orc_lambda(const heroam_config *conf, unsigned long long he_number);
float orc_radius_sim(unsigned long long he_number);   /* simulation.c:448 */
double orc_radius_init(unsigned long long he_number); /* simulation.c:297 */

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
