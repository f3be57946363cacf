/*
 * host/heroam_h5.h -- per-step trajectory files in the reference's HDF5 layout, written and
 * read WITHOUT libhdf5 (the library is not part of this image; the on-disk format is the
 * published HDF5 File Format Specification, version 0 superblock / "old style" groups, which is
 * what H5Fcreate + H5Gcreate + H5Dcreate with default property lists produce).
 *
 *   heroam_h5_create         <->  initialize_file           reference h5file.c:69-76
 *   heroam_h5_save_timestep  <->  save_timestep_to_hdf5     reference h5file.c:88-112
 *   heroam_h5_close          <->  close_file                reference h5file.c:146-149
 *   heroam_h5_open           <->  open_file                 reference h5file.c:78-85
 *   heroam_h5_num_steps      <->  H5Gget_num_objs on "/"    reference reader.c:83-85
 *   heroam_h5_read_timestep  <->  read_timestep_from_hdf5   reference h5file.c:115-143
 *
 * Layout of a file (reference simulation.c:462, h5file.c:21-49, :94-107):
 *   ./results/trajectory_%06d.h5
 *     /Step_<k>                group, one per iteration k of the step loop
 *       particles              dataset, shape (no,), compound of 84 bytes with the members
 *                              index, success (u32), time, force_mag, velocity_sq (f32),
 *                              orientation f32[4], position, velocity, ang_velocity, force f32[3]
 *                              at the offsets of `Particle` (simulation.h:37-48)
 */
#ifndef HEROAM_H5_H_
#define HEROAM_H5_H_

#include "../include/heroam_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct heroam_h5_writer heroam_h5_writer;
typedef struct heroam_h5_reader heroam_h5_reader;

/* Creates (truncates) the file.  NULL if it cannot be opened. */
heroam_h5_writer *heroam_h5_create(const char *filename);
/* Appends group /Step_<step> with its `particles` dataset.  0 = ok, -1 = I/O error or a step
 * number that was already written. */
int heroam_h5_save_timestep(heroam_h5_writer *w, unsigned int step, const heroam_particle *records, unsigned int no);
/* Writes the root group's index (B-tree, name heap), fixes up the superblock, closes.  0 = ok. */
int heroam_h5_close(heroam_h5_writer *w);

/* Opens a file written by heroam_h5_create or by libhdf5 with default property lists
 * (superblock 0/1, symbol-table groups, contiguous datasets).  NULL on failure;
 * heroam_h5_last_error() says why. */
heroam_h5_reader *heroam_h5_open(const char *filename);
/* Number of objects in the root group (= time steps in the file). */
long heroam_h5_num_steps(const heroam_h5_reader *r);
/* Reads /Step_<step>/particles; mallocs out->particle (release with free), sets out->no and
 * out->index = step as the reference does.  0 = ok, -1 = no such step / malformed. */
int heroam_h5_read_timestep(heroam_h5_reader *r, unsigned int step, heroam_particles *out);
void heroam_h5_reader_close(heroam_h5_reader *r);
const char *heroam_h5_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
