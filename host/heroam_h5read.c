/*
 * host/heroam_h5read.c -- `h5read <filename> <time_step>`: post-processing of a per-step
 * trajectory file, the job of the reference's second executable (reader.c:65-228), on top of
 * host/heroam_h5.c instead of libhdf5.
 *
 * Same command line, same products:
 *   trajectory_<p>.bin        per particle, one float32 triple per time step: position divided by
 *                             |position of particle 0 at step 0|   (reader.c:104-147, byte-identical)
 *   images/step_%010llu.png   one frame every <time_step> steps: unit sphere with a 20 x 20 graticule,
 *                             each particle's path so far as a line and its current place as a dot
 *                             (reader.c:150-222)
 *
 * Built differently from the reference: the file is streamed step by step (memory is one step, not the
 * whole file), and the frames are drawn by ONE gnuplot script with a loop (images/frames.gp) instead of
 * one gnuplot process per frame.  gnuplot is not part of this image: the script is always written and
 * only run when a gnuplot binary is on PATH.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>

#include "heroam_h5.h"

static double seconds_since(clock_t t0) { return (double)(clock() - t0) / CLOCKS_PER_SEC; }

/* Pass over the file: normalised positions of every particle, appended step by step. */
static int export_positions(heroam_h5_reader *file, long n_steps, unsigned *n_particles_out) {
  FILE **out = NULL;
  unsigned n_particles = 0;
  double radius = 1.0;
  int rc = 0;
  for (long t = 0; t < n_steps && rc == 0; ++t) {
    heroam_particles step;
    if (heroam_h5_read_timestep(file, (unsigned)t, &step) < 0) {
      fprintf(stderr, "Error reading timestep %llu\n", (unsigned long long)t);
      rc = 1;
      break;
    }
    if (t == 0) {
      if (step.no == 0) {
        fprintf(stderr, "No particle data in the first step.\n");
        free(step.particle);
        return 1;
      }
      const float *p = step.particle[0].position; /* float sum, then a double square root: reader.c:106-109 */
      radius = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
      n_particles = step.no;
      out = (FILE **)calloc(n_particles, sizeof *out);
      for (unsigned k = 0; out && k < n_particles; ++k) {
        char name[64];
        snprintf(name, sizeof name, "trajectory_%u.bin", k);
        out[k] = fopen(name, "wb");
        if (!out[k]) fprintf(stderr, "Error creating file for particle %u\n", k);
      }
      if (!out) rc = 1;
    }
    /* like the reference, every step is expected to hold the particles of step 0 */
    for (unsigned k = 0; rc == 0 && k < n_particles && k < step.no; ++k) {
      if (!out[k]) continue;
      const float *p = step.particle[k].position;
      const float unit[3] = {(float)(p[0] / radius), (float)(p[1] / radius), (float)(p[2] / radius)};
      if (fwrite(unit, sizeof unit, 1, out[k]) != 1) rc = 1;
    }
    free(step.particle);
  }
  for (unsigned k = 0; out && k < n_particles; ++k)
    if (out[k]) fclose(out[k]);
  free(out);
  *n_particles_out = n_particles;
  return rc;
}

/* The graticule as inline data: parallels every 9 degrees, meridians every 18. */
static void graticule(FILE *g) {
  const double pi = acos(-1.0);
  const int n = 20;
  fprintf(g, "$parallels << EOD\n");
  for (int k = 1; k < n; ++k) {
    const double polar = pi * k / n;
    for (int j = 0; j <= n; ++j) fprintf(g, "%f %f %f\n", sin(polar) * cos(2 * pi * j / n), sin(polar) * sin(2 * pi * j / n), cos(polar));
    fprintf(g, "\n");
  }
  fprintf(g, "EOD\n$meridians << EOD\n");
  for (int j = 0; j < n; ++j) {
    for (int k = 0; k <= n; ++k) fprintf(g, "%f %f %f\n", sin(pi * k / n) * cos(2 * pi * j / n), sin(pi * k / n) * sin(2 * pi * j / n), cos(pi * k / n));
    fprintf(g, "\n");
  }
  fprintf(g, "EOD\n");
}

static int write_frames_script(const char *path, long n_steps, int every, unsigned n_particles) {
  FILE *g = fopen(path, "w");
  if (!g) return 1;
  fprintf(g, "# frames of h5read: gnuplot %s\n", path);
  fprintf(g, "set terminal pngcairo enhanced size 800,800\n");
  fprintf(g, "unset border; unset key; unset tics\n");
  fprintf(g, "set xzeroaxis; set yzeroaxis; set zzeroaxis; set xyplane at 0\n");
  fprintf(g, "set xrange [-1.2:1.2]; set yrange [-1.2:1.2]; set zrange [-1.2:1.2]\n");
  fprintf(g, "set view equal xyz; set size ratio -1\n");
  fprintf(g, "colour(p) = word('#007ACC #FF4500 #32CD32 #FFD700 #8A2BE2 #FF69B4 #40E0D0 #DC143C #6495ED #FF8C00', p %% 10 + 1)\n");
  fprintf(g, "path(p) = sprintf('trajectory_%%d.bin', p)\n");
  graticule(g);
  fprintf(g, "do for [t = 0:%ld:%d] {\n", n_steps - 1, every);
  fprintf(g, "  set output sprintf('images/step_%%010d.png', t)\n");
  fprintf(g, "  set title sprintf('Time %%d fs', t)\n");
  fprintf(g, "  splot $parallels w l lc 'gray', $meridians w l lc 'gray', \\\n");
  fprintf(g, "    for [p = 0:%u] path(p) binary format='%%float%%float%%float' every ::0::(t > 0 ? t - 1 : 0) u 1:2:3 w l lc rgb colour(p), \\\n",
          n_particles - 1);
  fprintf(g, "    for [p = 0:%u] path(p) binary format='%%float%%float%%float' every ::t::t u 1:2:3 w p pt 7 lc rgb colour(p)\n",
          n_particles - 1);
  fprintf(g, "  unset output\n}\n");
  fclose(g);
  return 0;
}

int main(int argc, char *argv[]) {
  if (argc != 3) {
    fprintf(stderr, "Usage: %s <filename> <time_step>\n", argv[0]);
    return 1;
  }
  const int every = atoi(argv[2]);
  heroam_h5_reader *file = heroam_h5_open(argv[1]);
  if (!file) {
    fprintf(stderr, "Error opening HDF5 file '%s'.\n", argv[1]);
    return 1;
  }
  const long n_steps = heroam_h5_num_steps(file);
  printf("Time steps in file: %llu\n", (unsigned long long)n_steps);
  clock_t t0 = clock();
  unsigned n_particles = 0;
  const int failed = n_steps <= 0 || export_positions(file, n_steps, &n_particles);
  heroam_h5_reader_close(file);
  if (failed) return 1;
  printf("Trajectory write time: %.2fs\n", seconds_since(t0));

  t0 = clock();
  struct stat st;
  if (stat("images", &st) == -1) mkdir("images", 0700);
  if (every > 0 && write_frames_script("images/frames.gp", n_steps, every, n_particles) == 0) {
    if (system("command -v gnuplot > /dev/null 2>&1") == 0) {
      if (system("gnuplot images/frames.gp") != 0) fprintf(stderr, "gnuplot failed on images/frames.gp\n");
    } else {
      printf("gnuplot is not on PATH: run `gnuplot images/frames.gp` where it is\n");
    }
  }
  printf("Total plotting time: %.2fs\n", seconds_since(t0));
  return 0;
}
