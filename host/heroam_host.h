/*
 * host/heroam_host.h -- the C host side that stays C (BASELINE.json north_star): config
 * file parsing, force-file loading, result files, and the HeRoam <config.conf> driver.
 * None of this is accelerated; it feeds and drains the C ABI of include/heroam_gpu.h.
 *
 *   heroam_read_config      <->  read_config            reference simulation.c:48-156
 *   heroam_print_config     <->  print_configuration    reference main.c:20-42
 *   heroam_load_force_file  <->  the fscanf loop        reference main.c:122-139
 *   heroam_write_forcelist  <->  forcelist.txt echo     reference main.c:147-153
 *   heroam_save_particles   <->  save_particles         reference simulation.c:415-436
 *   heroam_write_particlefile <-> particlefile_<N> / particlefile_after_<N>, main.c:60-89
 */
#ifndef HEROAM_HOST_H_
#define HEROAM_HOST_H_

#include <stdio.h>

#include "../include/heroam_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Same file format, same 21 keys, same conversions (strtol / strtof), same diagnostics
 * on stderr, same return convention: 1 = ok, 0 = could not open.  Deviation (SURVEY.md
 * Q13): *conf is zero-initialised first, the reference parses into uninitialised stack. */
int heroam_read_config(const char *filename, heroam_config *conf);
void heroam_print_config(FILE *out, const heroam_config *conf);

/* Reads column 2 of the 4-column text file into table[0..length).  Returns the number of
 * rows read (the reference's `count`), or -1 if the file cannot be opened.  Deviation
 * (Q12): rows beyond `length` are not stored (the reference overflows by one float); a
 * short file leaves the tail 0 instead of uninitialised.
 * A file name beginning with "synthetic" selects the built-in T-ref table instead. */
long heroam_load_force_file(const char *filename, float *table, long length);

/* T-ref generator: table[i] = -(amplitude * (36/d_i)^half_power), d_i = start + grid*i,
 * evaluated in double with multiplications only (bit-identical to
 * sphere_roaming_b200/forcetable.py). */
void heroam_synthetic_table(float *table, long length, double start, double grid, double amplitude,
                            int half_power);

int heroam_write_forcelist(const char *path, const heroam_config *conf, const float *table, long count);

/* One TSV row per particle: trajectory index, particle number, success, time, r, v, F;
 * "%d\t%d\t%d\t%.10e\t..." exactly as the reference. */
void heroam_save_particles(FILE *f, const heroam_particles *p);
int heroam_write_particlefile(const char *path, const heroam_particles *all, int number);

#ifdef __cplusplus
}
#endif
#endif
