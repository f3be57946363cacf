/*
 * oracle/ref_patched_driver.c -- TEST INFRASTRUCTURE: INTEGRATION.md section 2, compiled and run.
 *
 * The reference's driver with the ONE change a maintainer would make: the OpenMP loop over
 * simulate_particles (reference src/main.c:76-78) replaced by heroam_gpu_simulate_batch.  Everything
 * around it is the reference's own object code, compiled from /root/reference by oracle/Makefile:
 * read_config, initialize_particles (libc rand), save_particles, free_particles.
 *
 * It also is the compile-level check of the boundary: the reference's simulation.h is included
 * FIRST, so the static assertions at the bottom of include/heroam_types.h compare heroam_config /
 * heroam_particle / heroam_particles with struct config / Particle / Particles, and the casts below
 * are between layout-identical types.
 *
 * tests/test_gpu_reference_patch.py runs this next to _ref/HeRoam_ref (the reference's unmodified
 * main.c) on the same config: ./results/particlefile_<N> and particlefile_after_<N> must be equal
 * byte for byte (exact mode).
 */
#include <stdio.h>
#include <stdlib.h>
#include <sys/stat.h>
#include <sys/types.h>

#include "simulation.h" /* the reference's (src/simulation.h) */

#include "heroam_gpu.h" /* after simulation.h: layouts are static-asserted equal */

static const char *const kHeader = "Index\tNumber\tSuccess\tTime\tX\tY\tZ\tVX\tVY\tVZ\tFX\tFY\tFZ\n";

static int dump(const char *stem, int helium_number, const Particles *particles, int number) {
  char name[96];
  snprintf(name, sizeof name, "./results/%s_%d", stem, helium_number);
  FILE *f = fopen(name, "w");
  if (!f) { perror(name); return 1; }
  fputs(kHeader, f);
  for (int i = 0; i < number; ++i) save_particles(f, particles + i);
  return fclose(f) != 0;
}

/* INTEGRATION.md section 2 */
static int simulate_for_number(const int helium_number, const struct config *conf, const float *Flist) {
  Particles *particles = (Particles *)malloc(conf->number * sizeof(Particles));
  initialize_particles(helium_number, conf, Flist, particles);
  int failed = dump("particlefile", helium_number, particles, conf->number);

  heroam_ctx *gpu = NULL;
  if (heroam_gpu_create((const heroam_config *)conf, Flist, /*device*/ 0, /*opt: exact mode*/ NULL, &gpu) ||
      heroam_gpu_simulate_batch(gpu, helium_number, (heroam_particles *)particles, conf->number, NULL)) {
    fprintf(stderr, "%s\n", heroam_gpu_last_error());
    exit(1);
  }
  heroam_gpu_destroy(gpu);

  failed |= dump("particlefile_after", helium_number, particles, conf->number);
  for (int i = 0; i < conf->number; ++i) free_particles(particles + i);
  free(particles);
  return failed;
}

int main(int argc, char **argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s <config_file>\n", argv[0]);
    return 1;
  }
  struct config con;
  if (!read_config(argv[1], &con)) {
    perror("config");
    return 1;
  }
  FILE *in = fopen(con.force_file, "r");
  if (!in) { perror(con.force_file); return 1; }
  float *table = (float *)malloc(con.force_grid_length * sizeof(float));
  long rows = 0;
  float r2, singlet, unused_a, unused_b;  /* column 2 is the table (main.c:136) */
  while (rows < con.force_grid_length && fscanf(in, "%f\t%f\t%f\t%f\n", &r2, &singlet, &unused_a, &unused_b) == 4)
    table[rows++] = singlet;
  fclose(in);
  mkdir("./results", 0700);
  int failed = 0;
  if (con.start)
    for (int n = con.start; n < con.stop; n += con.step) failed |= simulate_for_number(n, &con, table);
  else
    failed = simulate_for_number(con.helium_number, &con, table);
  free(table);
  return failed;
}
