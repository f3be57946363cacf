/* host/heroam_table.c -- force table: text loader and the synthetic T-ref generator. */
#include "heroam_host.h"

#include <string.h>

void heroam_synthetic_table(float *table, long length, double start, double grid, double amplitude,
                            int half_power) {
  for (long i = 0; i < length; ++i) {
    const double d = start + grid * (double)i;
    const double ratio = 36.0 / d;
    double v = amplitude;
    for (int k = 0; k < half_power; ++k) v = v * ratio;
    table[i] = (float)(-v);
  }
}

long heroam_load_force_file(const char *filename, float *table, long length) {
  memset(table, 0, (size_t)length * sizeof(float));
  if (strncmp(filename, "synthetic", 9) == 0) { /* not a file: said out loud, the reference would try to open it */
    fprintf(stderr, "force_file = %s: using the built-in synthetic T-ref table (%ld entries), not a file\n", filename, length);
    heroam_synthetic_table(table, length, 4.0, 0.001, 1.0e-6, 2);
    return length;
  }
  FILE *f = fopen(filename, "r");
  if (!f) return -1;
  float r2, value, singlet_triplet, triplet_triplet; /* columns of main.c:131-136 */
  long count = 0;
  while (count < length &&
         fscanf(f, "%f\t%f\t%f\t%f\n", &r2, &value, &singlet_triplet, &triplet_triplet) == 4)
    table[count++] = value;
  fclose(f);
  return count;
}

int heroam_write_forcelist(const char *path, const heroam_config *conf, const float *table, long count) {
  FILE *f = fopen(path, "w");
  if (!f) return 0;
  fprintf(f, "Distance_Squared\tForce\n");
  for (int i = 0; i < count; ++i) /* float arithmetic, int index: as main.c:151-152 */
    fprintf(f, "%f\t%.10e\n", conf->force_grid_start + conf->force_grid * i, table[i]);
  const int bad = ferror(f);
  return fclose(f) == 0 && !bad;
}
