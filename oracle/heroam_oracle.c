/*
 * oracle/heroam_oracle.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 * See heroam_oracle.h for the role and the pinning of this file.
 *
 * The arithmetic below follows the reference operation by operation (same
 * float/double types, same evaluation order) so that a strict-IEEE build
 * (`-O2 -ffp-contract=off`, see Makefile) agrees bit-for-bit with the
 * reference compiled the same way.  It is organised differently: value-type
 * helpers, one flat record array per trajectory, no per-step heap traffic, and
 * the three phases of a step (kick+drift, forces, kick) as separate functions.
 *
 * Where the reference's behaviour is undefined the oracle DEFINES it
 * (SURVEY.md Q-list); each such place is marked "DEFINED:".
 */
#include "heroam_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* small value types                                                          */
/* ------------------------------------------------------------------------- */
typedef struct { float x, y, z; } v3;
typedef struct { float a, x, y, z; } qt; /* (scalar, vector) as vector.h:15-19 */

/* vector.c:16-22 */
static inline v3 v3_cross(v3 u, v3 w) {
  v3 r;
  r.x = u.y * w.z - u.z * w.y;
  r.y = u.z * w.x - u.x * w.z;
  r.z = u.x * w.y - u.y * w.x;
  return r;
}
/* vector.c:47-53 : component * scalar */
static inline v3 v3_scale(v3 u, float s) {
  v3 r = {u.x * s, u.y * s, u.z * s};
  return r;
}
/* vector.c:39-45 */
static inline v3 v3_add(v3 u, v3 w) {
  v3 r = {u.x + w.x, u.y + w.y, u.z + w.z};
  return r;
}
/* vector.c:28-30 */
static inline float v3_norm_sq(v3 u) { return u.x * u.x + u.y * u.y + u.z * u.z; }
/* vector.c:24-26 : the float sum goes through the double sqrt and back to float */
static inline float v3_norm(v3 u) { return (float)sqrt((double)v3_norm_sq(u)); }

/* vector.c:166-171 */
static inline qt qt_mul(qt p, qt q) {
  qt r;
  r.a = p.a * q.a - p.x * q.x - p.y * q.y - p.z * q.z;
  r.x = p.a * q.x + p.x * q.a + p.y * q.z - p.z * q.y;
  r.y = p.a * q.y - p.x * q.z + p.y * q.a + p.z * q.x;
  r.z = p.a * q.z + p.x * q.y - p.y * q.x + p.z * q.a;
  return r;
}
/* vector.c:145-150 */
static inline qt qt_scale(qt q, float s) {
  qt r = {q.a * s, q.x * s, q.y * s, q.z * s};
  return r;
}
/* vector.c:152-154 */
static inline float qt_norm(qt q) {
  return (float)sqrt((double)(q.a * q.a + q.x * q.x + q.y * q.y + q.z * q.z));
}
/* vector.c:182-188 : conjugate over squared norm; `-1 * c / n2` is (-c)/n2 */
static inline qt qt_inverse(qt q) {
  float n2 = q.a * q.a + q.x * q.x + q.y * q.y + q.z * q.z;
  qt r = {q.a / n2, (-1 * q.x) / n2, (-1 * q.y) / n2, (-1 * q.z) / n2};
  return r;
}
/* vector.c:94-110 with :191-205 : image of the pole (0,0,1) under q . q^-1 */
static inline v3 qt_pole_image(qt q) {
  qt pole = {0.0f, 0.0f, 0.0f, 1.0f};
  qt inv = qt_inverse(q);
  qt r = qt_mul(qt_mul(q, pole), inv);
  v3 out = {r.x, r.y, r.z};
  return out;
}

static inline v3 ld3(const float *f) { v3 r = {f[0], f[1], f[2]}; return r; }
static inline void st3(float *f, v3 u) { f[0] = u.x; f[1] = u.y; f[2] = u.z; }
static inline qt ld4(const float *f) { qt r = {f[0], f[1], f[2], f[3]}; return r; }
static inline void st4(float *f, qt q) { f[0] = q.a; f[1] = q.x; f[2] = q.y; f[3] = q.z; }

/* ------------------------------------------------------------------------- */
/* loop-invariant scalars, rounded exactly where the reference rounds them    */
/* ------------------------------------------------------------------------- */
typedef struct {
  float dt;        /* simulation.c:443                                              */
  float half_dt;   /* dt/2.0 evaluated in double, narrowed at the call (:486, :406)  */
  float radius;    /* :448  (float)(2.22 * pow((float)N, 1.0/3.0))                   */
  float acc_scale; /* :351  (float)(1.0 / mass / radius / radius), double chain      */
  float inv_grid;  /* :162  (float)(1.0 / force_grid)                                */
  float grid_start;/* :163                                                           */
  float icd2;      /* :164  float product                                            */
  long grid_len;   /* :219                                                           */
  float max_time;  /* :469                                                           */
} consts;

float orc_radius_sim(unsigned long long he_number) {
  return (float)(2.22 * pow((double)(float)he_number, 1.0 / 3.0));
}
double orc_radius_init(unsigned long long he_number) {
  return 2.22 * pow((double)he_number, 1.0 / 3.0);
}
double orc_lambda(const heroam_config *conf, unsigned long long he_number) {
  /* simulation.c:289-293 : float product and quotient first, then double */
  double lam = (double)((conf->intensity * conf->pulse_width * conf->cross_section) / conf->hv);
  lam = (double)he_number * lam;
  lam = 6.241509074E-06 * lam;
  return lam;
}

static consts make_consts(const heroam_config *conf, unsigned long long he_number) {
  consts c;
  c.dt = conf->dt;
  c.half_dt = (float)((double)conf->dt / 2.0);
  c.radius = orc_radius_sim(he_number);
  c.acc_scale = (float)(1.0 / (double)conf->mass / (double)c.radius / (double)c.radius);
  c.inv_grid = (float)(1.0 / (double)conf->force_grid);
  c.grid_start = conf->force_grid_start;
  c.icd2 = conf->icd_dist * conf->icd_dist;
  c.grid_len = conf->force_grid_length;
  c.max_time = conf->max_time;
  return c;
}

/* ------------------------------------------------------------------------- */
/* pair interaction (simulation.c:158-236)                                    */
/* ------------------------------------------------------------------------- */
static inline float pair_dist_sq(const heroam_particle *pj, const heroam_particle *pk) {
  const float dx = pj->position[0] - pk->position[0];
  const float dy = pj->position[1] - pk->position[1];
  const float dz = pj->position[2] - pk->position[2];
  return dx * dx + dy * dy + dz * dz;
}

/* Table bin of a squared distance (:218).  DEFINED (Q5): the reference has no
 * lower bound check and converts out-of-range floats to int (undefined); here
 * anything that does not truncate to a bin in [0, len) means "no force".     */
static inline long table_bin(const consts *c, float d) {
  const float pos = (d - c->grid_start) * c->inv_grid;
  if (!(pos > -1.0f) || !((double)pos < (double)c->grid_len)) return -1;
  return (long)(int)pos;
}

static void forces_and_meetings(const consts *c, uint32_t n, heroam_particle *p,
                                const float *table, orc_counters *ctr) {
  for (uint32_t j = 0; j < n; ++j) { /* :170-176 -- frozen particles included */
    p[j].force[0] = 0.0f;
    p[j].force[1] = 0.0f;
    p[j].force[2] = 0.0f;
    p[j].force_mag = 0.0f;
  }
  /* pass 1 (:178-202): greedy, order dependent.  j keeps scanning after a hit,
   * so one sweep can freeze three or more particles (Q3).                     */
  for (uint32_t j = 0; j < n; ++j) {
    if (p[j].success) continue;
    for (uint32_t k = j + 1; k < n; ++k) {
      if (p[k].success) continue;
      if (pair_dist_sq(p + j, p + k) < c->icd2) {
        p[j].success = 1;
        p[k].success = 1;
      }
    }
  }
  /* pass 2 (:205-233): central force from the table, Newton's third law.
   * The reference re-reads the distances stored in pass 1; recomputing them
   * from the unchanged positions gives the same bits.                         */
  for (uint32_t j = 0; j + 1 < n; ++j) {
    if (p[j].success) continue;
    for (uint32_t k = j + 1; k < n; ++k) {
      if (p[k].success) continue;
      const float d = pair_dist_sq(p + j, p + k);
      const long bin = table_bin(c, d);
      if (ctr) ctr->pair_steps++;
      if (bin < 0) continue;
      if (ctr) ctr->inrange_steps++;
      const float scale = (float)((double)table[bin] / sqrt((double)d)); /* :220 */
      const float fx = scale * (p[j].position[0] - p[k].position[0]);
      const float fy = scale * (p[j].position[1] - p[k].position[1]);
      const float fz = scale * (p[j].position[2] - p[k].position[2]);
      p[j].force[0] += fx; p[j].force[1] += fy; p[j].force[2] += fz;
      p[k].force[0] -= fx; p[k].force[1] -= fy; p[k].force[2] -= fz;
    }
    p[j].force_mag = v3_norm(ld3(p[j].force)); /* :232 -- never for the last index (Q9) */
  }
}

void orc_calculate_force(uint32_t no, heroam_particle *p, const heroam_config *conf,
                         const float *table) {
  consts c = make_consts(conf, 1);
  forces_and_meetings(&c, no, p, table, NULL);
}

/* ------------------------------------------------------------------------- */
/* one velocity-Verlet step (simulation.c:478-509)                            */
/* ------------------------------------------------------------------------- */
/* angular acceleration from the torque (:343-352) */
static inline v3 ang_accel(const consts *c, const heroam_particle *p) {
  return v3_scale(v3_cross(ld3(p->position), ld3(p->force)), c->acc_scale);
}

/* first half kick + drift of the orientation + re-projection (:483-488, :394-413) */
static void kick_and_drift(const consts *c, heroam_particle *p, v3 *w_half) {
  /* w(t+dt/2) = w(t) + (dt/2) a(t)   (:354-361) */
  *w_half = v3_add(ld3(p->ang_velocity), v3_scale(ang_accel(c, p), c->half_dt));
  /* q += (dt/2) (0,w) q ; q /= |q| ; r = R * q(0,0,1)q^-1 */
  qt omega = {0.0f, w_half->x, w_half->y, w_half->z};
  qt q = ld4(p->orientation);
  qt dq = qt_scale(qt_mul(omega, q), c->half_dt);
  q.a = q.a + dq.a; q.x = q.x + dq.x; q.y = q.y + dq.y; q.z = q.z + dq.z;
  q = qt_scale(q, (float)(1.0 / (double)qt_norm(q)));
  st4(p->orientation, q);
  st3(p->position, v3_scale(qt_pole_image(q), c->radius));
}

/* second half kick, stored velocity r x w, clock (:499-507) */
static void kick_and_clock(const consts *c, heroam_particle *p, v3 w_half) {
  v3 w = v3_add(w_half, v3_scale(ang_accel(c, p), c->half_dt));
  st3(p->ang_velocity, w);
  v3 v = v3_cross(ld3(p->position), w);
  st3(p->velocity, v);
  p->velocity_sq = v3_norm_sq(v);
  p->time += c->dt;
}

/* :33-46 -- equality, not >= (Q4) */
static int all_paired(uint32_t n, const heroam_particle *p, uint32_t *flagged_out) {
  uint32_t flagged = 0;
  for (uint32_t i = 0; i < n; ++i) flagged += p[i].success;
  if (flagged_out) *flagged_out = flagged;
  return (n % 2 == 0) ? (flagged == n) : (flagged == n - 1);
}

orc_result orc_simulate_one(const heroam_config *conf, unsigned long long he_number,
                            const float *table, uint32_t no, heroam_particle *p,
                            uint64_t max_steps, orc_counters *ctr) {
  const consts c = make_consts(conf, he_number);
  v3 *w_half = (v3 *)calloc(no ? no : 1, sizeof(v3)); /* :451-456, once per trajectory */
  orc_result res = {0, ORC_DONE_SUCCESS, 0, 0.0f};
  int done = 0;
  float now = 0.0f; /* :467 */
  while (!done && now < c.max_time) { /* :469 */
    if (max_steps && res.steps >= max_steps) { res.status = ORC_DONE_STEPCAP; break; }
    uint32_t moved = 0; /* particles that complete this step (their clock advances) */
    for (uint32_t i = 0; i < no; ++i)
      if (!p[i].success) kick_and_drift(&c, p + i, w_half + i);
    forces_and_meetings(&c, no, p, table, ctr);
    for (uint32_t i = 0; i < no; ++i)
      if (!p[i].success) { kick_and_clock(&c, p + i, w_half[i]); ++moved; }
    if (ctr) ctr->particle_steps += moved;
    res.steps++;
    float latest = p[0].time; /* :512-517 */
    for (uint32_t i = 1; i < no; ++i)
      if (latest < p[i].time) latest = p[i].time;
    const int advanced = latest > now;
    now = latest;
    done = all_paired(no, p, &res.n_flagged);
    if (!done && now < c.max_time) { /* the reference's own exits take precedence */
      uint32_t active = 0;
      for (uint32_t i = 0; i < no; ++i) active += !p[i].success;
      /* DEFINED: nobody left to move and the == criterion is unmet -> the
       * reference loops forever (time frozen below max_time).  Stop. */
      if (active == 0) { res.status = ORC_DONE_DEADLOCK; break; }
      /* DEFINED (Q10): float clock stuck -> the reference loops forever. */
      if (!advanced) { res.status = ORC_DONE_STALLED; break; }
    }
  }
  if (res.status == ORC_DONE_SUCCESS && !done) res.status = ORC_DONE_MAXTIME;
  if (res.steps == 0) all_paired(no, p, &res.n_flagged);
  res.time = now;
  free(w_half);
  return res;
}

void orc_simulate(const heroam_config *conf, unsigned long long he_number, const float *table,
                  int number, const uint32_t *counts, heroam_particle *flat,
                  uint64_t max_steps, int nthreads, orc_result *results, orc_counters *ctr) {
  size_t *first = (size_t *)malloc(((size_t)number + 1) * sizeof(size_t));
  first[0] = 0;
  for (int i = 0; i < number; ++i) first[i + 1] = first[i] + counts[i];
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
  uint64_t ps = 0, prs = 0, irs = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : ps, prs, irs)
  for (int i = 0; i < number; ++i) {
    orc_counters local = {0, 0, 0};
    orc_result r = orc_simulate_one(conf, he_number, table, counts[i], flat + first[i],
                                    max_steps, ctr ? &local : NULL);
    if (results) results[i] = r;
    ps += local.particle_steps; prs += local.pair_steps; irs += local.inrange_steps;
  }
  if (ctr) { ctr->particle_steps = ps; ctr->pair_steps = prs; ctr->inrange_steps = irs; }
  free(first);
}

/* ------------------------------------------------------------------------- */
/* Philox4x32-10 (Salmon et al., SC'11) -- counter-based stream                */
/* ------------------------------------------------------------------------- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int round = 0; round < 10; ++round) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* ------------------------------------------------------------------------- */
/* sampling stage (simulation.c:238-337, vector.c:74-134)                      */
/* ------------------------------------------------------------------------- */
/* A source of 31-bit integers with the range of glibc rand(): [0, 2^31-1]. */
typedef struct {
  int kind;          /* 0 = libc rand(), 1 = Philox */
  uint32_t key[2];
  uint32_t ctr[4];   /* ctr[0] = block number within this trajectory's stream */
  uint32_t buf[4];
  int have;          /* unread words in buf */
} draw_src;

static void src_begin_philox(draw_src *s, long seed, unsigned long long he_number,
                             uint64_t trajectory) {
  s->kind = 1;
  s->key[0] = (uint32_t)((uint64_t)seed & 0xffffffffu);
  s->key[1] = (uint32_t)((uint64_t)seed >> 32) ^ (uint32_t)(he_number >> 32);
  s->ctr[0] = 0;
  s->ctr[1] = (uint32_t)(trajectory & 0xffffffffu);
  s->ctr[2] = (uint32_t)(trajectory >> 32);
  s->ctr[3] = (uint32_t)(he_number & 0xffffffffu);
  s->have = 0;
}
static inline int src_next(draw_src *s) {
  if (s->kind == 0) return rand();
  if (s->have == 0) {
    orc_philox4x32_10(s->ctr, s->key, s->buf);
    s->ctr[0]++;
    s->have = 4;
  }
  const uint32_t word = s->buf[4 - s->have];
  s->have--;
  return (int)(word >> 1);
}
/* (float)rand() / RAND_MAX : int -> float, RAND_MAX -> float (2^31), float divide */
static inline float src_unit_f(draw_src *s) { return (float)src_next(s) / (float)2147483647; }
/* DEFINED: the reference feeds a uniform that can be exactly 0 into log()
 * (simulation.c:249-251, probability 2^-31 per draw) and gets an infinite speed.
 * Draws that enter a logarithm are taken from (0,1]: a zero is replaced by 1/2^31. */
static inline float src_unit_f_pos(draw_src *s) {
  int r = src_next(s);
  if (r == 0) r = 1;
  return (float)r / (float)2147483647;
}

/* simulation.c:257-275 (Knuth), retried until >= 1 by the caller (:302-304) */
static int poisson_knuth(draw_src *s, double limit) {
  int n = 0;
  double p = 1.0, u;
  do {
    n = n + 1;
    u = (double)src_next(s) / (double)2147483647;
    p = p * u;
  } while (p > limit);
  return n - 1;
}

/* vector.c:112-134 (Marsaglia 1972), uniform on S^3 */
static qt random_orientation(draw_src *s) {
  float x, y, z, u, v, w;
  do {
    x = 2 * src_unit_f(s) - 1;
    y = 2 * src_unit_f(s) - 1;
    z = x * x + y * y;
  } while (z > 1);
  do {
    u = 2 * src_unit_f(s) - 1;
    v = 2 * src_unit_f(s) - 1;
    w = u * u + v * v;
  } while (w > 1);
  const float sc = (float)sqrt((double)((1 - z) / w));
  qt q = {x, y, sc * u, sc * v};
  const float mag = qt_norm(q);
  q.a /= mag; q.x /= mag; q.y /= mag; q.z /= mag;
  return q;
}

/* simulation.c:238-255 : sigma * chi_3 via Box-Muller, three of four normals */
static float thermal_speed(draw_src *s, float sigma) {
  const float pi = (float)(4 * atan(1));
  const float u1 = src_unit_f_pos(s);
  const float u2 = src_unit_f(s);
  const float u3 = src_unit_f_pos(s);
  const float u4 = src_unit_f(s);
  const float z1 = (float)(sqrt(-2.0 * log((double)u1)) * cos(2.0 * (double)pi * (double)u2));
  const float z2 = (float)(sqrt(-2.0 * log((double)u1)) * sin(2.0 * (double)pi * (double)u2));
  const float z3 = (float)(sqrt(-2.0 * log((double)u3)) * cos(2.0 * (double)pi * (double)u4));
  return (float)(sqrt((double)(z1 * z1 + z2 * z2 + z3 * z3)) * (double)sigma);
}

/* vector.c:74-92 : a tangent vector, a x r with a = (-y, x, 0) away from the poles */
static v3 tangent_of(v3 r) {
  v3 a = {0, 0, 0};
  if (r.x != 0 || r.y != 0) { a.x = -1 * r.y; a.y = r.x; a.z = 0; }
  else { a.x = 0; a.y = -1 * r.z; a.z = r.y; }
  return v3_cross(a, r);
}

long orc_sample(const heroam_config *conf, unsigned long long he_number, const float *table,
                int source, uint64_t first_index, int number, uint32_t *counts,
                heroam_particle *flat, long capacity) {
  const consts c = make_consts(conf, he_number);
  const double lambda = orc_lambda(conf, he_number);
  const double limit = exp(-1.0 * lambda);          /* :263 */
  const double radius = orc_radius_init(he_number); /* :297 (double here, float in the loop) */
  draw_src src;
  memset(&src, 0, sizeof src);
  if (source == 0) srand((unsigned int)conf->seed); /* :283-286; seed 0 (time) not modelled */
  long at = 0;
  for (int i = 0; i < number; ++i) {
    if (source != 0) src_begin_philox(&src, conf->seed, he_number, first_index + (uint64_t)i);
    int n;
    do { n = poisson_knuth(&src, limit); } while (n < 1); /* zero-truncated */
    counts[i] = (uint32_t)n;
    if (at + n > capacity) { /* keep consuming the stream so the count stays right */
      for (int j = 0; j < n; ++j) { (void)random_orientation(&src); (void)thermal_speed(&src, conf->velocity); }
      at += n;
      continue;
    }
    heroam_particle *p = flat + at;
    for (int j = 0; j < n; ++j) {
      memset(p + j, 0, sizeof *p); /* DEFINED (Q1): time starts at 0 */
      p[j].index = (uint32_t)j;
      const qt q = random_orientation(&src);
      st4(p[j].orientation, q);
      const v3 r = v3_scale(qt_pole_image(q), (float)radius); /* :319-320 */
      st3(p[j].position, r);
      const float vel = (float)((double)thermal_speed(&src, conf->velocity) * 1E-5); /* :323 */
      const float ang_vel = (float)((double)vel / radius);                            /* :324 */
      p[j].velocity_sq = vel * vel;
      const v3 tan = tangent_of(r);
      const v3 w = v3_scale(tan, ang_vel / v3_norm(tan)); /* :328 */
      st3(p[j].ang_velocity, w);
      st3(p[j].velocity, v3_cross(r, w)); /* :329 -- stored as r x w */
    }
    forces_and_meetings(&c, (uint32_t)n, p, table, NULL); /* :334 */
    at += n;
  }
  return at <= capacity ? at : -at;
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------- */
/* known-answer entry points: expose the internal pieces so tests can pin    */
/* them against the reference's vector.c / simulation.c one function at a    */
/* time (tests/golden/vector_kats.npz, sampler_streams.npz)                  */
/* ------------------------------------------------------------------------- */
void orc_kat_cross(const float *a, const float *b, float *out) { st3(out, v3_cross(ld3(a), ld3(b))); }
void orc_kat_add(const float *a, const float *b, float *out) { st3(out, v3_add(ld3(a), ld3(b))); }
void orc_kat_scale(float s, const float *a, float *out) { st3(out, v3_scale(ld3(a), s)); }
float orc_kat_norm(const float *a) { return v3_norm(ld3(a)); }
float orc_kat_norm_sq(const float *a) { return v3_norm_sq(ld3(a)); }
void orc_kat_tangent(const float *a, float *out) { st3(out, tangent_of(ld3(a))); }
void orc_kat_qmul(const float *p, const float *q, float *out) { st4(out, qt_mul(ld4(p), ld4(q))); }
void orc_kat_qinv(const float *q, float *out) { st4(out, qt_inverse(ld4(q))); }
float orc_kat_qnorm(const float *q) { return qt_norm(ld4(q)); }
void orc_kat_pole(const float *q, float *out) { st3(out, qt_pole_image(ld4(q))); }
/* find_accelration + update_angvel + update_orientation on one particle */
void orc_kat_accel(float mass, float radius, const float *pos, const float *force, float *acc) {
  consts c;
  memset(&c, 0, sizeof c);
  c.acc_scale = (float)(1.0 / (double)mass / (double)radius / (double)radius);
  heroam_particle p;
  memset(&p, 0, sizeof p);
  memcpy(p.position, pos, sizeof p.position);
  memcpy(p.force, force, sizeof p.force);
  st3(acc, ang_accel(&c, &p));
}
void orc_kat_orientation(float radius, float dt, float *quat, const float *w_half, float *pos) {
  consts c;
  memset(&c, 0, sizeof c);
  c.radius = radius;
  c.half_dt = (float)((double)dt / 2.0);
  heroam_particle p;
  memset(&p, 0, sizeof p);
  memcpy(p.orientation, quat, sizeof p.orientation);
  /* kick_and_drift with zero force leaves w_half == ang_velocity */
  memcpy(p.ang_velocity, w_half, sizeof p.ang_velocity);
  v3 wh;
  kick_and_drift(&c, &p, &wh);
  memcpy(quat, p.orientation, sizeof p.orientation);
  memcpy(pos, p.position, sizeof p.position);
}
void orc_kat_poisson(unsigned int seed, double lambda, int n, int *out) {
  draw_src s;
  memset(&s, 0, sizeof s);
  srand(seed);
  const double limit = exp(-1.0 * lambda);
  for (int i = 0; i < n; ++i) out[i] = poisson_knuth(&s, limit);
}
void orc_kat_speed(unsigned int seed, float sigma, int n, float *out) {
  draw_src s;
  memset(&s, 0, sizeof s);
  srand(seed);
  for (int i = 0; i < n; ++i) out[i] = thermal_speed(&s, sigma);
}
void orc_kat_quat(unsigned int seed, int n, float *out) {
  draw_src s;
  memset(&s, 0, sizeof s);
  srand(seed);
  for (int i = 0; i < n; ++i) st4(out + 4 * i, random_orientation(&s));
}
int orc_find_success(uint32_t no, const heroam_particle *p) { return all_paired(no, p, NULL); }
