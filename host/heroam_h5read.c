/*
 * host/heroam_h5read.c -- `h5read <filename> <time_step>`: the post-processing tool of the
 * reference (reader.c:65-228) on top of host/heroam_h5.c instead of libhdf5.
 *
 *   reader.c                                   here
 *   ---------------------------------------    ----------------------------------------------
 *   open_file, H5Gget_num_objs (:74-85)        heroam_h5_open, heroam_h5_num_steps
 *   read_timestep_from_hdf5 per step (:88-100) heroam_h5_read_timestep
 *   radius = |position of particle 0, step 0|  same (:104-109)
 *   trajectory_<p>.bin: float32 x,y,z / radius same bytes (:123-147)
 *   one gnuplot process per plotted step       images/step_%010llu.gp, the same commands written
 *   piping the splot command (:185-222)        to a script; run through `gnuplot` when it is on PATH
 *
 * Like the reference it assumes every step holds the same number of particles as step 0.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>

#include "heroam_h5.h"

#define PI 3.14159265358979323846
#define N_LAT 20
#define N_LON 20

static void setup_gnuplot(FILE *g) { /* reader.c:14-30 */
  fprintf(g, "set terminal pngcairo enhanced size 800,800\n");
  fprintf(g, "unset border\nunset key\nset xzeroaxis\nset yzeroaxis\nset zzeroaxis\nset xyplane at 0\nunset tics\n");
  fprintf(g, "set parametric\nset xrange [-1.2:1.2]\nset yrange [-1.2:1.2]\nset zrange [-1.2:1.2]\n");
  fprintf(g, "set view equal xyz\nset size ratio -1\nset autoscale fix\n");
}

static void sphere_grid(FILE *g) { /* the inline data of the two '-' plots, reader.c:32-63 */
  for (int k = 1; k < N_LAT; k++) {
    const double theta = (PI * k) / N_LAT, z = cos(theta), r = sin(theta);
    for (int j = 0; j <= N_LON; j++) {
      const double phi = (2 * PI * j) / N_LON;
      fprintf(g, "%f %f %f\n", r * cos(phi), r * sin(phi), z);
    }
    fprintf(g, "\n");
  }
  fprintf(g, "e\n");
  for (int j = 0; j < N_LON; j++) {
    const double phi = (2 * PI * j) / N_LON;
    for (int k = 0; k <= N_LAT; k++) {
      const double theta = (PI * k) / N_LAT;
      fprintf(g, "%f %f %f\n", sin(theta) * cos(phi), sin(theta) * sin(phi), cos(theta));
    }
    fprintf(g, "\n");
  }
  fprintf(g, "e\n");
}

static int gnuplot_on_path(void) { return system("command -v gnuplot > /dev/null 2>&1") == 0; }

int main(int argc, char *argv[]) {
  if (argc != 3) {
    fprintf(stderr, "Usage: %s <filename> <time_step>\n", argv[0]);
    return 1;
  }
  const char *filename = argv[1];
  const int plot_time_step = atoi(argv[2]);
  heroam_h5_reader *file = heroam_h5_open(filename);
  if (!file) {
    fprintf(stderr, "Error opening HDF5 file '%s'.\n", filename);
    return 1;
  }
  clock_t start = clock();
  const long num_time = heroam_h5_num_steps(file);
  printf("Time steps in file: %llu\n", (unsigned long long)num_time);
  heroam_particles *particles = (heroam_particles *)calloc((size_t)(num_time > 0 ? num_time : 1), sizeof *particles);
  for (long i = 0; i < num_time; ++i) {
    if (heroam_h5_read_timestep(file, (unsigned)i, &particles[i]) < 0) {
      fprintf(stderr, "Error reading timestep %llu\n", (unsigned long long)i);
      for (long j = 0; j < i; j++) free(particles[j].particle);
      free(particles);
      heroam_h5_reader_close(file);
      return 1;
    }
  }
  heroam_h5_reader_close(file);
  printf("Reading time: %.2fs\n", (double)(clock() - start) / CLOCKS_PER_SEC);
  if (num_time <= 0 || particles[0].no == 0) {
    fprintf(stderr, "No particle data in '%s'.\n", filename);
    free(particles);
    return 1;
  }
  const unsigned int num_particles = particles[0].no;
  const float *first_pos = particles[0].particle[0].position;
  const double radius = sqrt(first_pos[0] * first_pos[0] + first_pos[1] * first_pos[1] + first_pos[2] * first_pos[2]);

  start = clock();
#pragma omp parallel for
  for (unsigned int p = 0; p < num_particles; p++) {
    char name[256];
    snprintf(name, sizeof name, "trajectory_%u.bin", p);
    FILE *out = fopen(name, "wb");
    if (!out) {
      fprintf(stderr, "Error creating file for particle %u\n", p);
      continue;
    }
    for (long t = 0; t < num_time; t++) {
      if (p >= particles[t].no) break;
      const float *pos = particles[t].particle[p].position;
      const float normalized[3] = {(float)(pos[0] / radius), (float)(pos[1] / radius), (float)(pos[2] / radius)};
      fwrite(normalized, sizeof(float), 3, out);
    }
    fclose(out);
  }
  printf("Trajectory write time: %.2fs\n", (double)(clock() - start) / CLOCKS_PER_SEC);
  for (long i = 0; i < num_time; i++) free(particles[i].particle);
  free(particles);

  start = clock();
  static const char *colors[] = {"#007ACC", "#FF4500", "#32CD32", "#FFD700", "#8A2BE2",
                                 "#FF69B4", "#40E0D0", "#DC143C", "#6495ED", "#FF8C00"};
  const int num_colors = (int)(sizeof colors / sizeof colors[0]);
  struct stat st;
  if (stat("images", &st) == -1) mkdir("images", 0700);
  const int have_gnuplot = gnuplot_on_path();
  if (plot_time_step > 0) {
#pragma omp parallel for schedule(dynamic)
    for (long t = 0; t < num_time; t += plot_time_step) {
      char script[64];
      snprintf(script, sizeof script, "images/step_%010llu.gp", (unsigned long long)t);
      FILE *g = fopen(script, "w");
      if (!g) continue;
      setup_gnuplot(g);
      fprintf(g, "set output 'images/step_%010llu.png'\n", (unsigned long long)t);
      fprintf(g, "set title 'Time %llu fs'\n", (unsigned long long)t);
      fprintf(g, "splot '-' w l lc 'gray', '-' w l lc 'gray'");
      for (unsigned int p = 0; p < num_particles; p++) {
        const char *color = colors[p % num_colors];
        if (t > 0)
          fprintf(g, ", 'trajectory_%u.bin' binary format='%%float%%float%%float' every ::0::%llu using 1:2:3 w l lc rgb '%s'",
                  p, (unsigned long long)(t - 1), color);
        fprintf(g, ", 'trajectory_%u.bin' binary format='%%float%%float%%float' every ::%llu::%llu using 1:2:3 w p pt 7 lc rgb '%s'",
                p, (unsigned long long)t, (unsigned long long)t, color);
      }
      fprintf(g, "\n");
      sphere_grid(g);
      fprintf(g, "unset output\n");
      fclose(g);
      if (have_gnuplot) {
        char cmd[128];
        snprintf(cmd, sizeof cmd, "gnuplot %s", script);
        if (system(cmd) != 0) fprintf(stderr, "gnuplot failed for timestep %llu\n", (unsigned long long)t);
      }
    }
  }
  if (!have_gnuplot) printf("gnuplot is not on PATH: wrote images/step_*.gp, render them with `gnuplot <script>`\n");
  printf("Total plotting time: %.2fs\n", (double)(clock() - start) / CLOCKS_PER_SEC);
  return 0;
}
