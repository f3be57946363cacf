/*
 * host/heroam_main.c -- the `HeRoam <config.conf>` driver (reference main.c:102-169) on top
 * of the C ABI: same CLI, same config keys, same ./results files.
 *
 *   simulate_for_number (main.c:50-97)       here
 *   -------------------------------------    -------------------------------------------
 *   initialize_particles                     heroam_gpu_sample_batch   (Philox stream)
 *   particlefile_<N>                         heroam_write_particlefile
 *   #pragma omp parallel for simulate_...    heroam_gpu_simulate_batch, index blocks over GPUs
 *   particlefile_after_<N>                   heroam_write_particlefile
 *   free_particles                           heroam_gpu_free_particles
 *   saveflag = 1: trajectory_%06d.h5         heroam_gpu_simulate_traced -> heroam_h5_save_timestep
 *     (simulation.c:460-476, h5file.c)         (host/heroam_h5.c, no libhdf5 needed)
 *
 * Extra, optional environment knobs (the config file format is unchanged):
 *   HEROAM_MODE=exact|fast   arithmetic mode (default exact)
 *   HEROAM_GPUS=n            number of GPUs to use (default: all visible)
 *   HEROAM_STATS=1           histogram-only run through heroam_gpu_run_stats (no TSV files)
 *   HEROAM_STREAM_ABOVE=n    runs of more than n trajectories (default 2^21) write their result files block by
 *                            block from flat buffers instead of holding every trajectory to the end
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <time.h>

#include "../include/heroam_gpu.h"
#include "heroam_h5.h"
#include "heroam_host.h"

#ifdef _OPENMP
#include <omp.h>
#endif

static int n_gpus_wanted(void) {
  int have = heroam_gpu_device_count();
  const char *env = getenv("HEROAM_GPUS");
  if (env && atoi(env) > 0 && atoi(env) < have) have = atoi(env);
  return have;
}

/* saveflag = 1 (simulation.c:460-476, :522-523): every iteration of every trajectory goes to
 * ./results/trajectory_%06d.h5 as /Step_<k>/particles.  Trajectories are traced in groups so
 * that the number of files open at once stays bounded. */
#define TRACE_GROUP 64
typedef struct {
  heroam_h5_writer *writer[TRACE_GROUP];
} trace_files;

static int trace_sink(void *user, uint32_t traj, uint32_t step, const heroam_particle *records, uint32_t no) {
  trace_files *tf = (trace_files *)user;
  return heroam_h5_save_timestep(tf->writer[traj], step, records, no);
}

static int simulate_traced_block(heroam_ctx *ctx, int helium_number, heroam_particles *block, int count) {
  int failed = 0;
  for (int g0 = 0; g0 < count && !failed; g0 += TRACE_GROUP) {
    const int n = count - g0 < TRACE_GROUP ? count - g0 : TRACE_GROUP;
    trace_files tf;
    uint32_t counts[TRACE_GROUP];
    size_t total = 0;
    memset(&tf, 0, sizeof tf);
    for (int i = 0; i < n; ++i) {
      char name[100];
      snprintf(name, sizeof name, "./results/trajectory_%06d.h5", (int)block[g0 + i].index); /* simulation.c:462 */
      tf.writer[i] = heroam_h5_create(name);
      if (!tf.writer[i]) {
        fprintf(stderr, "%s\n", heroam_h5_last_error());
        failed = 1;
      }
      counts[i] = block[g0 + i].no;
      total += counts[i];
    }
    heroam_particle *flat = (heroam_particle *)malloc((total ? total : 1) * sizeof *flat);
    size_t at = 0;
    for (int i = 0; i < n; ++i) {
      memcpy(flat + at, block[g0 + i].particle, counts[i] * sizeof *flat);
      at += counts[i];
    }
    if (!failed && heroam_gpu_simulate_traced(ctx, (unsigned long long)helium_number, counts, flat, n, NULL, trace_sink, &tf)) {
      fprintf(stderr, "%s\n", heroam_gpu_last_error());
      failed = 1;
    }
    at = 0;
    for (int i = 0; i < n; ++i) {
      memcpy(block[g0 + i].particle, flat + at, counts[i] * sizeof *flat);
      at += counts[i];
      if (tf.writer[i] && heroam_h5_close(tf.writer[i])) failed = 1;
    }
    free(flat);
  }
  return failed;
}

#define STREAM_ABOVE (1 << 21)
#define STREAM_BLOCK (1 << 20)
static int stream_for_number(int helium_number, const heroam_config *conf, heroam_ctx **ctx, int n_gpus);

static int simulate_for_number(int helium_number, const heroam_config *conf, const float *table,
                               heroam_ctx **ctx, int n_gpus) {
  (void)table; /* already resident on every GPU */
  const int number = conf->number;
  const char *stream = getenv("HEROAM_STREAM_ABOVE"); /* tests lower the threshold */
  if (!conf->saveflag && number > (stream ? atoi(stream) : STREAM_ABOVE)) return stream_for_number(helium_number, conf, ctx, n_gpus);
  heroam_particles *all = (heroam_particles *)calloc((size_t)(number > 0 ? number : 1), sizeof *all);
  if (!all) return 1;
  int failed = 0;
  /* sampling stage: each GPU draws its own block of the global index space */
#pragma omp parallel for num_threads(n_gpus) schedule(static, 1) reduction(| : failed)
  for (int g = 0; g < n_gpus; ++g) {
    const long first = (long)g * number / n_gpus, last = (long)(g + 1) * number / n_gpus;
    if (heroam_gpu_sample_batch(ctx[g], (unsigned long long)helium_number, (unsigned long long)first,
                                (int)(last - first), all + first)) {
      fprintf(stderr, "GPU %d: %s\n", g, heroam_gpu_last_error());
      failed |= 1;
    }
  }
  char path[96];
  snprintf(path, sizeof path, "./results/particlefile_%d", helium_number);
  if (!failed && !heroam_write_particlefile(path, all, number)) {
    perror(path);
    failed = 1;
  }
  if (failed) { /* whatever the GPUs that did succeed allocated is released */
    for (int i = 0; i < number; ++i) heroam_gpu_free_particles(all + i);
    free(all);
    return 1;
  }
  /* integration stage: the loop of main.c:76-78 */
#pragma omp parallel for num_threads(n_gpus) schedule(static, 1) reduction(| : failed)
  for (int g = 0; g < n_gpus; ++g) {
    const long first = (long)g * number / n_gpus, last = (long)(g + 1) * number / n_gpus;
    if (conf->saveflag) {
      failed |= simulate_traced_block(ctx[g], helium_number, all + first, (int)(last - first));
    } else if (heroam_gpu_simulate_batch(ctx[g], (unsigned long long)helium_number, all + first, (int)(last - first), NULL)) {
      fprintf(stderr, "GPU %d: %s\n", g, heroam_gpu_last_error());
      failed |= 1;
    }
  }
  snprintf(path, sizeof path, "./results/particlefile_after_%d", helium_number);
  if (!failed && !heroam_write_particlefile(path, all, number)) {
    perror(path);
    failed = 1;
  }
  for (int i = 0; i < number; ++i) heroam_gpu_free_particles(all + i);
  free(all);
  return failed;
}

/* Large runs (number > STREAM_ABOVE, no per-step files): the same two result files, written in trajectory order
 * from flat page-sized blocks instead of holding `number` separately malloc'ed arrays to the end (main.c:55,
 * simulation.c:310).  Each GPU takes STREAM_BLOCK consecutive trajectories of a round, samples them
 * (heroam_gpu_sample_flat), keeps a copy of the initial records, integrates them (heroam_gpu_simulate_flat); the
 * rows of the round then go to particlefile_<N> and particlefile_after_<N>.  Memory: 2 x STREAM_BLOCK x ~4 records
 * per GPU, whatever `number` is. */
static void write_rows(FILE *f, long first_traj, const uint32_t *counts, const heroam_particle *flat, long n) {
  long at = 0;
  for (long i = 0; i < n; ++i) {
    heroam_particles one;
    one.no = counts[i];
    one.index = (uint32_t)(first_traj + i);
    one.particle = (heroam_particle *)(flat + at);
    heroam_save_particles(f, &one);
    at += counts[i];
  }
}

static int stream_for_number(int helium_number, const heroam_config *conf, heroam_ctx **ctx, int n_gpus) {
  const long number = conf->number;
  const long cap = (long)STREAM_BLOCK * 6 + 4096; /* records per block: mean 4 per trajectory at the repository's lambda */
  char path[96];
  snprintf(path, sizeof path, "./results/particlefile_%d", helium_number);
  FILE *before = fopen(path, "w");
  snprintf(path, sizeof path, "./results/particlefile_after_%d", helium_number);
  FILE *after = fopen(path, "w");
  uint32_t **counts = (uint32_t **)calloc((size_t)n_gpus, sizeof *counts);
  heroam_particle **init = (heroam_particle **)calloc((size_t)n_gpus, sizeof *init);
  heroam_particle **fin = (heroam_particle **)calloc((size_t)n_gpus, sizeof *fin);
  long *got = (long *)calloc((size_t)n_gpus, sizeof *got);
  int failed = !before || !after || !counts || !init || !fin || !got;
  for (int g = 0; g < n_gpus && !failed; ++g) {
    counts[g] = (uint32_t *)malloc((size_t)STREAM_BLOCK * sizeof(uint32_t));
    init[g] = (heroam_particle *)malloc((size_t)cap * sizeof(heroam_particle));
    fin[g] = (heroam_particle *)malloc((size_t)cap * sizeof(heroam_particle));
    failed = !counts[g] || !init[g] || !fin[g];
  }
  if (failed) fprintf(stderr, "cannot open the result files or allocate the streaming blocks\n");
  static const char *header = "Index\tNumber\tSuccess\tTime\tX\tY\tZ\tVX\tVY\tVZ\tFX\tFY\tFZ\n"; /* main.c:65-69 */
  if (!failed) { fputs(header, before); fputs(header, after); }
  for (long round0 = 0; round0 < number && !failed; round0 += (long)n_gpus * STREAM_BLOCK) {
#pragma omp parallel for num_threads(n_gpus) schedule(static, 1) reduction(| : failed)
    for (int g = 0; g < n_gpus; ++g) {
      const long first = round0 + (long)g * STREAM_BLOCK;
      long n = number - first;
      if (n > STREAM_BLOCK) n = STREAM_BLOCK;
      got[g] = 0;
      if (n <= 0) continue;
      long total = 0;
      if (heroam_gpu_sample_flat(ctx[g], (unsigned long long)helium_number, (unsigned long long)first, (int)n, counts[g],
                                 init[g], cap, &total) || total > cap) {
        fprintf(stderr, "GPU %d: %s\n", g, total > cap ? "a block needs more records than reserved" : heroam_gpu_last_error());
        failed |= 1;
        continue;
      }
      memcpy(fin[g], init[g], (size_t)total * sizeof(heroam_particle));
      if (heroam_gpu_simulate_flat(ctx[g], (unsigned long long)helium_number, counts[g], fin[g], (int)n, NULL)) {
        fprintf(stderr, "GPU %d: %s\n", g, heroam_gpu_last_error());
        failed |= 1;
        continue;
      }
      got[g] = n;
    }
    for (int g = 0; g < n_gpus && !failed; ++g) {
      write_rows(before, round0 + (long)g * STREAM_BLOCK, counts[g], init[g], got[g]);
      write_rows(after, round0 + (long)g * STREAM_BLOCK, counts[g], fin[g], got[g]);
    }
    if (!failed && (ferror(before) || ferror(after))) {
      perror("./results/particlefile");
      failed = 1;
    }
  }
  if (before && fclose(before)) failed = 1;
  if (after && fclose(after)) failed = 1;
  for (int g = 0; g < n_gpus; ++g) {
    if (counts) free(counts[g]);
    if (init) free(init[g]);
    if (fin) free(fin[g]);
  }
  free(counts); free(init); free(fin); free(got);
  return failed;
}

static int stats_for_number(int helium_number, const heroam_config *conf, heroam_ctx **ctx, int n_gpus) {
  heroam_stats *st = (heroam_stats *)calloc((size_t)n_gpus, sizeof *st);
  int failed = 0;
#pragma omp parallel for num_threads(n_gpus) schedule(static, 1) reduction(| : failed)
  for (int g = 0; g < n_gpus; ++g) {
    const long first = (long)g * conf->number / n_gpus, last = (long)(g + 1) * conf->number / n_gpus;
    if (heroam_gpu_run_stats(ctx[g], (unsigned long long)helium_number, (unsigned long long)first,
                             (unsigned long long)(last - first), st + g)) {
      fprintf(stderr, "GPU %d: %s\n", g, heroam_gpu_last_error());
      failed |= 1;
    }
  }
  if (!failed) {
    char path[96];
    snprintf(path, sizeof path, "./results/histograms_%d.tsv", helium_number);
    FILE *f = fopen(path, "w");
    if (f) {
      unsigned long long traj = 0, psteps = 0;
      double ms = 0;
      fprintf(f, "Bin\tTrajectoryExitTime\tParticleMeetingTime\n");
      for (int b = 0; b < HEROAM_HIST_T; ++b) {
        unsigned long long a = 0, m = 0;
        for (int g = 0; g < n_gpus; ++g) { a += st[g].hist_traj_time[b]; m += st[g].hist_meet_time[b]; }
        fprintf(f, "%d\t%llu\t%llu\n", b, a, m);
      }
      fclose(f);
      for (int g = 0; g < n_gpus; ++g) {
        traj += st[g].trajectories; psteps += st[g].info.particle_steps;
        if (st[g].info.device_ms > ms) ms = st[g].info.device_ms;
      }
      printf("N=%d: %llu trajectories, %llu particle-steps in %.1f ms on %d GPU(s): %.3e particle-steps/s\n",
             helium_number, traj, psteps, ms, n_gpus, psteps / (ms * 1e-3));
    }
  }
  free(st);
  return failed;
}

/* The droplet-size sweep of main.c:158-164 in histogram mode: ONE pool run per GPU for all sizes
 * (heroam_gpu_run_size_sweep), every size sharded by trajectory index over the GPUs. */
static int stats_for_sweep(const heroam_config *conf, heroam_ctx **ctx, int n_gpus) {
  int n_sizes = 0;
  for (int n = conf->start; n < conf->stop; n += conf->step) ++n_sizes;
  if (n_sizes < 1) return 0;
  if (n_sizes > 32 || conf->step <= 0) return -1; /* the caller falls back to the loop of runs */
  heroam_stats *st = (heroam_stats *)calloc((size_t)n_gpus * (size_t)n_sizes, sizeof *st);
  heroam_size_point *pts = (heroam_size_point *)calloc((size_t)n_gpus * (size_t)n_sizes, sizeof *pts);
  if (!st || !pts) { free(st); free(pts); return 1; }
  int failed = 0;
#pragma omp parallel for num_threads(n_gpus) schedule(static, 1) reduction(| : failed)
  for (int g = 0; g < n_gpus; ++g) {
    heroam_size_point *mine = pts + (size_t)g * n_sizes;
    for (int i = 0; i < n_sizes; ++i) {
      const long first = (long)g * conf->number / n_gpus, last = (long)(g + 1) * conf->number / n_gpus;
      mine[i].he_number = (unsigned long long)(conf->start + i * conf->step);
      mine[i].first_index = (unsigned long long)first;
      mine[i].count = (unsigned long long)(last - first);
    }
    if (heroam_gpu_run_size_sweep(ctx[g], n_sizes, mine, st + (size_t)g * n_sizes)) {
      fprintf(stderr, "GPU %d: %s\n", g, heroam_gpu_last_error());
      failed |= 1;
    }
  }
  for (int i = 0; i < n_sizes && !failed; ++i) {
    char path[96];
    const int helium_number = conf->start + i * conf->step;
    snprintf(path, sizeof path, "./results/histograms_%d.tsv", helium_number);
    FILE *f = fopen(path, "w");
    if (!f) { perror(path); failed = 1; break; }
    fprintf(f, "Bin\tTrajectoryExitTime\tParticleMeetingTime\n");
    unsigned long long traj = 0;
    for (int b = 0; b < HEROAM_HIST_T; ++b) {
      unsigned long long a = 0, m = 0;
      for (int g = 0; g < n_gpus; ++g) {
        a += st[(size_t)g * n_sizes + i].hist_traj_time[b];
        m += st[(size_t)g * n_sizes + i].hist_meet_time[b];
      }
      fprintf(f, "%d\t%llu\t%llu\n", b, a, m);
    }
    if (fclose(f)) failed = 1;
    for (int g = 0; g < n_gpus; ++g) traj += st[(size_t)g * n_sizes + i].trajectories;
    printf("N=%d: %llu trajectories\n", helium_number, traj);
  }
  if (!failed) {
    unsigned long long psteps = 0;
    double ms = 0;
    for (int g = 0; g < n_gpus; ++g) {
      psteps += st[(size_t)g * n_sizes].info.particle_steps;
      if (st[(size_t)g * n_sizes].info.device_ms > ms) ms = st[(size_t)g * n_sizes].info.device_ms;
    }
    printf("%d droplet sizes as one run: %llu particle-steps in %.1f ms on %d GPU(s): %.3e particle-steps/s\n", n_sizes, psteps,
           ms, n_gpus, psteps / (ms * 1e-3));
  }
  free(st);
  free(pts);
  return failed;
}

int main(int argc, char *argv[]) {
  if (argc < 2) {
    fprintf(stderr, "Usage: %s <config_file>\n", argv[0]);
    return 1;
  }
  heroam_config conf;
  if (!heroam_read_config(argv[1], &conf)) {
    perror("Error reading configuration file");
    return 1;
  }
  if (conf.print) heroam_print_config(stdout, &conf);
  if (conf.force_grid_length <= 0) {
    fprintf(stderr, "force_grid_length must be positive\n");
    return 1;
  }
  float *table = (float *)malloc((size_t)conf.force_grid_length * sizeof(float));
  const long count = heroam_load_force_file(conf.force_file, table, conf.force_grid_length);
  if (count < 0) {
    perror("Error opening file");
    return 1;
  }
  struct stat st;
  if (stat("./results", &st) == -1) mkdir("./results", 0700);
  if (!heroam_write_forcelist("./results/forcelist.txt", &conf, table, count)) {
    perror("./results/forcelist.txt");
    free(table);
    return 1;
  }
  /* seed = 0 means "seed from the clock" (simulation.c:283-286).  Resolved ONCE here, so that every GPU
   * samples from the same Philox stream, and reported, so that the run can be repeated. */
  if (conf.seed == 0) {
    conf.seed = (long)time(NULL);
    printf("seed = 0: seeding from the clock, seed = %ld\n", conf.seed);
  }

  const int n_gpus = n_gpus_wanted();
  if (n_gpus < 1) {
    fprintf(stderr, "no CUDA device visible: HeRoam-B200 has no CPU path\n");
    return 1;
  }
  heroam_options opt;
  memset(&opt, 0, sizeof opt);
  const char *mode = getenv("HEROAM_MODE");
  opt.mode = (mode && strcmp(mode, "fast") == 0) ? HEROAM_MODE_FAST : HEROAM_MODE_EXACT;
  heroam_ctx **ctx = (heroam_ctx **)calloc((size_t)n_gpus, sizeof *ctx);
  for (int g = 0; g < n_gpus; ++g)
    if (heroam_gpu_create(&conf, table, g, &opt, &ctx[g])) {
      fprintf(stderr, "GPU %d: %s\n", g, heroam_gpu_last_error());
      for (int h = 0; h < g; ++h) heroam_gpu_destroy(ctx[h]);
      free(ctx);
      free(table);
      return 1;
    }
  const int stats_only = getenv("HEROAM_STATS") && atoi(getenv("HEROAM_STATS")) != 0;
  int failed = 0;
  /* droplet-size sweep exactly as main.c:158-164; in histogram mode all sizes share one pool run */
  int swept = conf.start && stats_only ? stats_for_sweep(&conf, ctx, n_gpus) : -1;
  if (swept >= 0) {
    failed = swept;
  } else if (conf.start) {
    for (int n = conf.start; n < conf.stop && !failed; n += conf.step)
      failed = stats_only ? stats_for_number(n, &conf, ctx, n_gpus) : simulate_for_number(n, &conf, table, ctx, n_gpus);
  } else {
    failed = stats_only ? stats_for_number(conf.helium_number, &conf, ctx, n_gpus)
                        : simulate_for_number(conf.helium_number, &conf, table, ctx, n_gpus);
  }
  for (int g = 0; g < n_gpus; ++g) heroam_gpu_destroy(ctx[g]);
  free(ctx);
  free(table);
  return failed;
}
