/*
 * oracle/hdf5_stub/h5_noop.c -- TEST INFRASTRUCTURE.
 *
 * The five entry points of the reference's h5file.h as no-ops (libhdf5 is not in this image), for
 * the two executables oracle/Makefile links from the reference's sources with saveflag = 0:
 * _ref/HeRoam_ref (the reference's own main.c, unmodified) and _ref/HeRoam_patched (the same driver
 * with INTEGRATION.md section 2's patch).  Also __wrap_malloc: both are linked with
 * -Wl,--wrap=malloc so that malloc hands out zeroed memory -- the reference reads Particle.time
 * before ever writing it (simulation.c:310-331 never sets it, :507 is `+=`), and the
 * two programs must start from the same, defined, value.
 */
#include <stdlib.h>

#include "h5file.h"

hid_t initialize_file(const char *filename) { (void)filename; return 0; }
hid_t open_file(const char *filename) { (void)filename; return 0; }
herr_t save_timestep_to_hdf5(const hid_t file_id, const unsigned int step, const Particles *particles) {
  (void)file_id; (void)step; (void)particles; return 0;
}
herr_t read_timestep_from_hdf5(const hid_t file_id, const unsigned int step, Particles *particles) {
  (void)file_id; (void)step; (void)particles; return -1;
}
void close_file(const hid_t file_id) { (void)file_id; }

void *__wrap_malloc(size_t size) { return calloc(1, size); }
