/* host/heroam_output.c -- particlefile_<N> / particlefile_after_<N> in the reference's format. */
#include "heroam_host.h"

void heroam_save_particles(FILE *f, const heroam_particles *p) {
  for (unsigned int i = 0; i < p->no; ++i) {
    const heroam_particle *q = p->particle + i;
    fprintf(f, "%d\t%d\t%d\t%.10e\t%.10e\t%.10e\t%.10e\t%.10e\t%.10e\t%.10e\t%.10e\t%.10e\t%.10e\n",
            p->index, i, q->success, q->time, q->position[0], q->position[1], q->position[2],
            q->velocity[0], q->velocity[1], q->velocity[2], q->force[0], q->force[1], q->force[2]);
  }
}

int heroam_write_particlefile(const char *path, const heroam_particles *all, int number) {
  FILE *f = fopen(path, "w");
  if (!f) return 0;
  fprintf(f, "Index\tNumber\tSuccess\tTime\tX\tY\tZ\tVX\tVY\tVZ\tFX\tFY\tFZ\n"); /* main.c:65-69 */
  for (int i = 0; i < number; ++i) heroam_save_particles(f, all + i);
  const int bad = ferror(f);
  return fclose(f) == 0 && !bad; /* 1 = written in full (a full disk shows up here) */
}
