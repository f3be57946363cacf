/* host/heroam_config.c -- config.conf parser and echo (see heroam_host.h). */
#include "heroam_host.h"

#include <ctype.h>
#include <stdlib.h>
#include <string.h>

/* strip leading and trailing white space in place */
static char *strip(char *s) {
  while (*s && isspace((unsigned char)*s)) ++s;
  size_t n = strlen(s);
  while (n > 0 && isspace((unsigned char)s[n - 1])) s[--n] = '\0';
  return s;
}

enum kind { K_INT, K_LONG, K_FLOAT, K_STR };
struct key {
  const char *name;
  enum kind kind;
  size_t offset;
};
#define KEY(n, k) {#n, k, offsetof(heroam_config, n)}
/* the 21 keys of the reference (simulation.c:85-148) */
static const struct key KEYS[] = {
    KEY(print, K_INT),          KEY(hv, K_FLOAT),
    KEY(intensity, K_FLOAT),    KEY(cross_section, K_FLOAT),
    KEY(pulse_width, K_FLOAT),  KEY(number, K_INT),
    KEY(saveflag, K_INT),       KEY(start, K_INT),
    KEY(step, K_INT),           KEY(stop, K_INT),
    KEY(helium_number, K_INT),  KEY(seed, K_LONG),
    KEY(mass, K_FLOAT),         KEY(velocity, K_FLOAT),
    KEY(dt, K_FLOAT),           KEY(max_time, K_FLOAT),
    KEY(icd_dist, K_FLOAT),     KEY(force_grid, K_FLOAT),
    KEY(force_grid_start, K_FLOAT), KEY(force_grid_length, K_LONG),
    KEY(force_file, K_STR),
};

int heroam_read_config(const char *filename, heroam_config *conf) {
  memset(conf, 0, sizeof *conf);
  FILE *f = fopen(filename, "r");
  if (!f) {
    fprintf(stderr, "Error opening config file: %s\n", filename);
    return 0;
  }
  char line[256];
  int lineno = 0;
  while (fgets(line, sizeof line, f)) {
    ++lineno;
    char *hash = strchr(line, '#'); /* inline comments */
    if (hash) *hash = '\0';
    char *text = strip(line);
    if (!*text) continue;
    char *eq = strchr(text, '=');
    if (!eq) {
      fprintf(stderr, "Invalid syntax at line %d: %s\n", lineno, text);
      continue;
    }
    *eq = '\0';
    char *name = strip(text);
    char *value = strip(eq + 1);
    const struct key *hit = NULL;
    for (size_t i = 0; i < sizeof KEYS / sizeof KEYS[0]; ++i)
      if (strcmp(name, KEYS[i].name) == 0) hit = &KEYS[i];
    if (!hit) {
      fprintf(stderr, "Unknown configuration key '%s' at line %d\n", name, lineno);
      continue;
    }
    char *field = (char *)conf + hit->offset;
    switch (hit->kind) {
      case K_INT: *(int *)field = (int)strtol(value, NULL, 10); break;
      case K_LONG: *(long *)field = strtol(value, NULL, 10); break;
      case K_FLOAT: *(float *)field = strtof(value, NULL); break;
      case K_STR:
        strncpy(field, value, sizeof conf->force_file - 1);
        field[sizeof conf->force_file - 1] = '\0';
        break;
    }
  }
  fclose(f);
  return 1;
}

void heroam_print_config(FILE *out, const heroam_config *c) {
  fprintf(out, "Configuration File Loaded\n");
  fprintf(out, "hv = %f\n", c->hv);
  fprintf(out, "intensity = %f\n", c->intensity);
  fprintf(out, "cross_section = %f\n", c->cross_section);
  fprintf(out, "pulse_width = %f\n", c->pulse_width);
  fprintf(out, "number = %d\n", c->number);
  fprintf(out, "saveflag = %d\n", c->saveflag);
  fprintf(out, "start = %d\n", c->start);
  fprintf(out, "step = %d\n", c->step);
  fprintf(out, "stop = %d\n", c->stop);
  fprintf(out, "helium_number = %d\n", c->helium_number);
  fprintf(out, "seed = %ld\n", c->seed);
  fprintf(out, "mass = %f\n", c->mass);
  fprintf(out, "velocity = %f\n", c->velocity);
  fprintf(out, "dt = %f\n", c->dt);
  fprintf(out, "max_time = %f\n", c->max_time);
  fprintf(out, "icd_dist = %f\n", c->icd_dist);
  fprintf(out, "force_grid = %f\n", c->force_grid);
  fprintf(out, "force_grid_start = %f\n", c->force_grid_start);
  fprintf(out, "force_grid_length = %ld\n", c->force_grid_length);
  fprintf(out, "force_file = %s\n", c->force_file);
}
