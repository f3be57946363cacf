/*
 * include/heroam_types.h -- plain-C data types at the drop-in boundary.
 *
 * These are layout-identical restatements of the reference's boundary types so
 * that a buffer produced by either side can be handed to the other unchanged:
 *
 *   heroam_config     <->  struct config   (reference src/simulation.h:10-32)
 *   heroam_particle   <->  Particle        (reference src/simulation.h:37-48, 84 bytes;
 *                                           the HDF5 compound of src/h5file.c:31-47 has
 *                                           the same member order and offsets)
 *   heroam_particles  <->  Particles       (reference src/simulation.h:53-57)
 *
 * If the reference's simulation.h was included first (SIMULATION_H_ defined) the
 * static assertions at the bottom check the two sets of types against each other,
 * so a reference-side translation unit may cast between them freely.
 */
#ifndef HEROAM_TYPES_H_
#define HEROAM_TYPES_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct heroam_config {
  int print;               /* echo the parsed configuration                     */
  float hv;                /* photon energy [eV]                                */
  float intensity;         /* laser intensity [GW/cm^2]                         */
  float cross_section;     /* absorption cross-section [Mb]                     */
  float pulse_width;       /* pulse duration [fs]                               */
  int number;              /* trajectories per droplet size                     */
  int saveflag;            /* per-step trajectory dump on/off                   */
  int start;               /* droplet-size sweep: first size (0 = no sweep)     */
  int step;                /* droplet-size sweep: increment                     */
  int stop;                /* droplet-size sweep: exclusive upper bound         */
  int helium_number;       /* droplet size when start == 0                      */
  long seed;               /* RNG seed (0 = time based)                         */
  float mass;              /* particle mass [amu]                               */
  float velocity;          /* sqrt(kT/m) [m/s]                                  */
  float dt;                /* time step [fs]                                    */
  float max_time;          /* integration horizon [fs]                          */
  float icd_dist;          /* meeting distance [A]                              */
  float force_grid;        /* spacing of the force table in r^2 [A^2]           */
  float force_grid_start;  /* r^2 of table entry 0 [A^2]                        */
  long force_grid_length;  /* number of table entries                           */
  char force_file[50];     /* path of the 4-column text force file              */
} heroam_config;

typedef struct heroam_particle {
  uint32_t index;          /* particle number inside its trajectory             */
  uint32_t success;        /* 1 once the particle has met a partner (frozen)    */
  float time;              /* time this particle has been propagated [fs]       */
  float force_mag;         /* |force|, as maintained by the reference (Q9)      */
  float velocity_sq;       /* |velocity|^2                                      */
  float orientation[4];    /* quaternion (a, x, y, z)                           */
  float position[3];       /* [A]                                               */
  float velocity[3];       /* stored as position x ang_velocity                 */
  float ang_velocity[3];   /* [1/fs]                                            */
  float force[3];          /* table units                                       */
} heroam_particle;

typedef struct heroam_particles {
  uint32_t no;             /* particles in this trajectory                      */
  uint32_t index;          /* trajectory number                                 */
  heroam_particle *particle;
} heroam_particles;

#if defined(__cplusplus)
#define HEROAM_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define HEROAM_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif

HEROAM_STATIC_ASSERT(sizeof(heroam_particle) == 84, "Particle must be 84 bytes");
HEROAM_STATIC_ASSERT(offsetof(heroam_particle, orientation) == 20, "orientation@20");
HEROAM_STATIC_ASSERT(offsetof(heroam_particle, position) == 36, "position@36");
HEROAM_STATIC_ASSERT(offsetof(heroam_particle, velocity) == 48, "velocity@48");
HEROAM_STATIC_ASSERT(offsetof(heroam_particle, ang_velocity) == 60, "ang_velocity@60");
HEROAM_STATIC_ASSERT(offsetof(heroam_particle, force) == 72, "force@72");
HEROAM_STATIC_ASSERT(sizeof(heroam_particles) == 16, "Particles must be 16 bytes");
HEROAM_STATIC_ASSERT(sizeof(heroam_config) == 152, "struct config must be 152 bytes");
HEROAM_STATIC_ASSERT(offsetof(heroam_config, seed) == 48, "seed@48");
HEROAM_STATIC_ASSERT(offsetof(heroam_config, force_grid_length) == 88, "force_grid_length@88");
HEROAM_STATIC_ASSERT(offsetof(heroam_config, force_file) == 96, "force_file@96");

#ifdef SIMULATION_H_
/* Compiled next to the reference's own header: prove the layouts agree. */
HEROAM_STATIC_ASSERT(sizeof(heroam_particle) == sizeof(Particle), "Particle layout");
HEROAM_STATIC_ASSERT(sizeof(heroam_particles) == sizeof(Particles), "Particles layout");
HEROAM_STATIC_ASSERT(sizeof(heroam_config) == sizeof(struct config), "config layout");
HEROAM_STATIC_ASSERT(offsetof(heroam_particle, position) == offsetof(Particle, position), "position");
HEROAM_STATIC_ASSERT(offsetof(heroam_particle, force) == offsetof(Particle, force), "force");
HEROAM_STATIC_ASSERT(offsetof(heroam_config, max_time) == offsetof(struct config, max_time), "max_time");
HEROAM_STATIC_ASSERT(offsetof(heroam_config, force_file) == offsetof(struct config, force_file), "force_file");
#endif

#ifdef __cplusplus
}
#endif
#endif /* HEROAM_TYPES_H_ */
