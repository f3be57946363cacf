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
 *
 * Extra, optional environment knobs (the config file format is unchanged):
 *   HEROAM_MODE=exact|fast   arithmetic mode (default exact)
 *   HEROAM_GPUS=n            number of GPUs to use (default: all visible)
 *   HEROAM_STATS=1           histogram-only run through heroam_gpu_run_stats (no TSV files)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/types.h>

#include "../include/heroam_gpu.h"
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

static int simulate_for_number(int helium_number, const heroam_config *conf, const float *table,
                               heroam_ctx **ctx, int n_gpus) {
  (void)table; /* already resident on every GPU */
  const int number = conf->number;
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
  if (failed) return 1;
  char path[96];
  snprintf(path, sizeof path, "./results/particlefile_%d", helium_number);
  heroam_write_particlefile(path, all, number);
  /* integration stage: the loop of main.c:76-78 */
#pragma omp parallel for num_threads(n_gpus) schedule(static, 1) reduction(| : failed)
  for (int g = 0; g < n_gpus; ++g) {
    const long first = (long)g * number / n_gpus, last = (long)(g + 1) * number / n_gpus;
    if (heroam_gpu_simulate_batch(ctx[g], (unsigned long long)helium_number, all + first, (int)(last - first), NULL)) {
      fprintf(stderr, "GPU %d: %s\n", g, heroam_gpu_last_error());
      failed |= 1;
    }
  }
  snprintf(path, sizeof path, "./results/particlefile_after_%d", helium_number);
  if (!failed) heroam_write_particlefile(path, all, number);
  for (int i = 0; i < number; ++i) heroam_gpu_free_particles(all + i);
  free(all);
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
  if (conf.saveflag) {
    fprintf(stderr, "saveflag = 1 (per-step HDF5 trajectory dump, h5file.c) is not available in this build: "
                    "libhdf5 is not installed; set saveflag = 0\n");
    return 1;
  }
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
  heroam_write_forcelist("./results/forcelist.txt", &conf, table, count);

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
      return 1;
    }
  const int stats_only = getenv("HEROAM_STATS") && atoi(getenv("HEROAM_STATS")) != 0;
  int failed = 0;
  /* droplet-size sweep exactly as main.c:158-164 */
  if (conf.start) {
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
