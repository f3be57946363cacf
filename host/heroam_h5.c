/*
 * host/heroam_h5.c -- the reference's per-step HDF5 trajectory layout (h5file.c:21-112) without
 * libhdf5: a writer that emits exactly the objects H5Fcreate / H5Gcreate / H5Dcreate / H5Dwrite
 * with default property lists create (superblock version 0, symbol-table groups indexed by a
 * version-1 B-tree over a local heap, version-1 object headers, contiguous dataset storage),
 * and a reader for files of that kind (ours or libhdf5's).
 *
 * On-disk structures follow the HDF5 File Format Specification:
 *   superblock v0      sig, versions, sizes (8/8), group K (leaf 4, internal 16), addresses, root entry
 *   object header v1   16-byte prefix, then messages (type u16, size u16, flags u8, 3 reserved, data padded to 8)
 *   messages used      0x0001 dataspace v1, 0x0003 datatype, 0x0005 fill value v2, 0x0008 layout v3,
 *                      0x0011 symbol table, 0x0012 modification time
 *   local heap         "HEAP": data segment of NUL-terminated names, 8-byte aligned
 *   B-tree v1, type 0  "TREE": keys are heap offsets of names, ordered by strcmp; leaves point to
 *   symbol nodes       "SNOD": up to 2*leafK entries of 40 bytes (name offset, header address, cache)
 *   datatype           compound v2 of 84 bytes; members u32 / f32 / array v2 of f32 (an array member
 *                      makes libhdf5 encode the enclosing compound as version 2)
 */
#include "heroam_h5.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define H5_UNDEF 0xffffffffffffffffull
#define LEAF_K 4u
#define INTERNAL_K 16u
#define SNOD_BYTES (8u + 2u * LEAF_K * 40u)                              /* 328 */
#define TREE_BYTES (24u + 2u * INTERNAL_K * 8u + (2u * INTERNAL_K + 1u) * 8u) /* 544 */
#define SUPERBLOCK_BYTES 96u
#define HEAP_FREE_NULL 1ull /* "no free block" marker of a local heap */

static _Thread_local char g_h5err[256] = "";  /* written from the per-GPU host threads */
const char *heroam_h5_last_error(void) { return g_h5err; }
static int h5fail(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_h5err, sizeof g_h5err, fmt, ap);
  va_end(ap);
  return -1;
}

/* ------------------------------------------------------------------------------------------
 * little-endian byte buffer
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  unsigned char *p;
  size_t n, cap;
  int oom; /* an append could not get memory: the buffer is incomplete and must not be written */
} bytes;

static int reserve_bytes(bytes *b, size_t extra) {
  if (b->n + extra <= b->cap) return 0;
  size_t cap = b->cap ? b->cap * 2 : 1024;
  while (cap < b->n + extra) cap *= 2;
  unsigned char *q = (unsigned char *)realloc(b->p, cap);
  if (!q) {
    b->oom = 1;
    return -1;
  }
  b->p = q;
  b->cap = cap;
  return 0;
}
static void put(bytes *b, const void *src, size_t n) {
  if (reserve_bytes(b, n)) return;
  memcpy(b->p + b->n, src, n);
  b->n += n;
}
static void put_zero(bytes *b, size_t n) {
  if (reserve_bytes(b, n)) return;
  memset(b->p + b->n, 0, n);
  b->n += n;
}
static void put_u8(bytes *b, unsigned v) {
  unsigned char c = (unsigned char)v;
  put(b, &c, 1);
}
static void put_le(bytes *b, unsigned long long v, int width) {
  unsigned char c[8];
  for (int i = 0; i < width; ++i) c[i] = (unsigned char)(v >> (8 * i));
  put(b, c, (size_t)width);
}
static void put_u16(bytes *b, unsigned v) { put_le(b, v, 2); }
static void put_u32(bytes *b, unsigned long v) { put_le(b, v, 4); }
static void put_u64(bytes *b, unsigned long long v) { put_le(b, v, 8); }
static void pad8(bytes *b) { put_zero(b, (8 - b->n % 8) % 8); }
static void poke_u64(unsigned char *at, unsigned long long v) {
  for (int i = 0; i < 8; ++i) at[i] = (unsigned char)(v >> (8 * i));
}
static unsigned long long peek_le(const unsigned char *at, int width) {
  unsigned long long v = 0;
  for (int i = 0; i < width; ++i) v |= (unsigned long long)at[i] << (8 * i);
  return v;
}

/* ------------------------------------------------------------------------------------------
 * datatype message of `Particle` (h5file.c:21-49)
 * ------------------------------------------------------------------------------------------ */
static void dt_u32(bytes *b) { /* H5T_NATIVE_UINT: class 0 v1, little endian, unsigned */
  put_u8(b, 0x10); put_u8(b, 0x00); put_u8(b, 0); put_u8(b, 0);
  put_u32(b, 4);
  put_u16(b, 0);  /* bit offset */
  put_u16(b, 32); /* precision  */
}
static void dt_f32(bytes *b) { /* H5T_NATIVE_FLOAT: class 1 v1, IEEE binary32 little endian */
  put_u8(b, 0x11); put_u8(b, 0x20); put_u8(b, 0x1f); put_u8(b, 0x00);
  put_u32(b, 4);
  put_u16(b, 0);  /* bit offset */
  put_u16(b, 32); /* precision */
  put_u8(b, 23);  /* exponent location */
  put_u8(b, 8);   /* exponent size */
  put_u8(b, 0);   /* mantissa location */
  put_u8(b, 23);  /* mantissa size */
  put_u32(b, 127);
}
static void dt_f32_array(bytes *b, unsigned n) { /* H5Tarray_create(H5T_NATIVE_FLOAT, 1, {n}): class 10 v2 */
  put_u8(b, 0x2a); put_u8(b, 0); put_u8(b, 0); put_u8(b, 0);
  put_u32(b, 4 * n);
  put_u8(b, 1); put_zero(b, 3); /* rank, reserved */
  put_u32(b, n);                /* dimension size */
  put_u32(b, 0);                /* permutation index */
  dt_f32(b);
}
static void dt_member(bytes *b, const char *name, unsigned offset) { /* compound v2 member prefix */
  const size_t len = strlen(name) + 1;
  put(b, name, len);
  put_zero(b, (8 - len % 8) % 8);
  put_u32(b, offset);
}
static void dt_particle(bytes *b) {
  put_u8(b, 0x26); put_u16(b, 10); put_u8(b, 0); /* class 6 v2, 10 members */
  put_u32(b, (unsigned long)sizeof(heroam_particle));
  dt_member(b, "index", offsetof(heroam_particle, index));               dt_u32(b);
  dt_member(b, "success", offsetof(heroam_particle, success));           dt_u32(b);
  dt_member(b, "time", offsetof(heroam_particle, time));                 dt_f32(b);
  dt_member(b, "force_mag", offsetof(heroam_particle, force_mag));       dt_f32(b);
  dt_member(b, "velocity_sq", offsetof(heroam_particle, velocity_sq));   dt_f32(b);
  dt_member(b, "orientation", offsetof(heroam_particle, orientation));   dt_f32_array(b, 4);
  dt_member(b, "position", offsetof(heroam_particle, position));         dt_f32_array(b, 3);
  dt_member(b, "velocity", offsetof(heroam_particle, velocity));         dt_f32_array(b, 3);
  dt_member(b, "ang_velocity", offsetof(heroam_particle, ang_velocity)); dt_f32_array(b, 3);
  dt_member(b, "force", offsetof(heroam_particle, force));               dt_f32_array(b, 3);
}

/* ------------------------------------------------------------------------------------------
 * object header v1
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  bytes *b;
  size_t start;    /* offset of the header prefix in b */
  size_t msg_head; /* offset of the message being written */
  unsigned nmsg;
} ohdr;

static void ohdr_begin(ohdr *h, bytes *b) {
  h->b = b;
  h->start = b->n;
  h->nmsg = 0;
  put_u8(b, 1); put_u8(b, 0); /* version, reserved */
  put_u16(b, 0);              /* number of messages (patched) */
  put_u32(b, 1);              /* object reference count */
  put_u32(b, 0);              /* header data size (patched) */
  put_zero(b, 4);             /* pad the prefix to 16 bytes */
}
static void msg_begin(ohdr *h, unsigned type, unsigned flags) {
  h->msg_head = h->b->n;
  put_u16(h->b, type);
  put_u16(h->b, 0); /* size (patched) */
  put_u8(h->b, flags);
  put_zero(h->b, 3);
}
static void msg_end(ohdr *h) {
  pad8(h->b);
  if (h->b->oom) return;
  const size_t size = h->b->n - h->msg_head - 8;
  h->b->p[h->msg_head + 2] = (unsigned char)size;
  h->b->p[h->msg_head + 3] = (unsigned char)(size >> 8);
  h->nmsg++;
}
static void ohdr_end(ohdr *h) {
  if (h->b->oom) return;
  const size_t size = h->b->n - h->start - 16;
  unsigned char *p = h->b->p + h->start;
  p[2] = (unsigned char)h->nmsg;
  p[3] = (unsigned char)(h->nmsg >> 8);
  for (int i = 0; i < 4; ++i) p[8 + i] = (unsigned char)(size >> (8 * i));
}
/* group header: one symbol-table message; returns the offsets of its two addresses */
static void group_header(bytes *b, size_t *btree_at, size_t *heap_at) {
  ohdr h;
  ohdr_begin(&h, b);
  msg_begin(&h, 0x0011, 0);
  *btree_at = b->n; put_u64(b, 0);
  *heap_at = b->n;  put_u64(b, 0);
  msg_end(&h);
  ohdr_end(&h);
}
static void local_heap_header(bytes *b, unsigned long long data_size, size_t *data_addr_at) {
  put(b, "HEAP", 4);
  put_u8(b, 0); put_zero(b, 3);
  put_u64(b, data_size);
  put_u64(b, HEAP_FREE_NULL);
  *data_addr_at = b->n; put_u64(b, 0);
}
static void symbol_entry(bytes *b, unsigned long long name_off, unsigned long long header, int is_group,
                         unsigned long long btree, unsigned long long heap) {
  put_u64(b, name_off);
  put_u64(b, header);
  put_u32(b, is_group ? 1 : 0); /* cache type: 1 = group, scratch holds its B-tree and heap */
  put_u32(b, 0);
  if (is_group) { put_u64(b, btree); put_u64(b, heap); }
  else put_zero(b, 16);
}

/* ------------------------------------------------------------------------------------------
 * writer
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  unsigned step;
  unsigned long long header, btree, heap, name_off;
  char name[20];
} root_entry;

struct heroam_h5_writer {
  FILE *f;
  unsigned long long eof;
  root_entry *ent;
  size_t n_ent, cap_ent;
  unsigned last_step;
  int sorted_input; /* steps arrived in increasing order so far (duplicate check is then O(1)) */
  /* per-step block for the current particle count: addresses relative to the block start */
  bytes tmpl;
  unsigned tmpl_no;
  size_t fix[8];
  int n_fix;
  size_t off_btree, off_heap, off_data;
  unsigned char *scratch;
  size_t scratch_cap;
  unsigned long mtime;
  int failed;
};

static void build_step_template(heroam_h5_writer *w, unsigned no) {
  bytes *b = &w->tmpl;
  b->n = 0;
  w->n_fix = 0;
  size_t g_btree_at, g_heap_at, heap_data_at;
  /* /Step_k: header, name heap ("" and "particles"), B-tree node, symbol node */
  group_header(b, &g_btree_at, &g_heap_at);
  w->off_heap = b->n;
  local_heap_header(b, 24, &heap_data_at);
  const size_t heap_data = b->n;
  put_zero(b, 8);
  put(b, "particles", 10);
  put_zero(b, 6);
  w->off_btree = b->n;
  put(b, "TREE", 4);
  put_u8(b, 0); put_u8(b, 0); put_u16(b, 1); /* group node, leaf level, one entry */
  put_u64(b, H5_UNDEF); put_u64(b, H5_UNDEF);
  put_u64(b, 0);                              /* key 0: "" */
  const size_t child_at = b->n; put_u64(b, 0);
  put_u64(b, 8);                              /* key 1: "particles" */
  put_zero(b, w->off_btree + TREE_BYTES - b->n);
  const size_t snod = b->n;
  put(b, "SNOD", 4);
  put_u8(b, 1); put_u8(b, 0); put_u16(b, 1);
  const size_t entry_at = b->n;
  symbol_entry(b, 8, 0, 0, 0, 0);
  put_zero(b, snod + SNOD_BYTES - b->n);
  /* the dataset */
  const size_t dset = b->n;
  ohdr h;
  ohdr_begin(&h, b);
  msg_begin(&h, 0x0001, 0); /* dataspace v1: rank 1, no maximum */
  put_u8(b, 1); put_u8(b, 1); put_u8(b, 0); put_zero(b, 5);
  put_u64(b, no);
  msg_end(&h);
  msg_begin(&h, 0x0003, 1); /* datatype (constant) */
  dt_particle(b);
  msg_end(&h);
  msg_begin(&h, 0x0005, 1); /* fill value v2: late allocation, write if set, default value */
  put_u8(b, 2); put_u8(b, 2); put_u8(b, 2); put_u8(b, 1);
  put_u32(b, 0);
  msg_end(&h);
  msg_begin(&h, 0x0008, 0); /* layout v3, contiguous */
  put_u8(b, 3); put_u8(b, 1);
  const size_t data_addr_at = b->n; put_u64(b, 0);
  put_u64(b, (unsigned long long)no * sizeof(heroam_particle));
  msg_end(&h);
  msg_begin(&h, 0x0012, 0); /* modification time v1 */
  put_u8(b, 1); put_zero(b, 3);
  put_u32(b, w->mtime);
  msg_end(&h);
  ohdr_end(&h);
  w->off_data = b->n;
  if (b->oom) return;
  /* relative addresses, to be rebased per step */
  poke_u64(b->p + g_btree_at, w->off_btree);  w->fix[w->n_fix++] = g_btree_at;
  poke_u64(b->p + g_heap_at, w->off_heap);    w->fix[w->n_fix++] = g_heap_at;
  poke_u64(b->p + heap_data_at, heap_data);   w->fix[w->n_fix++] = heap_data_at;
  poke_u64(b->p + child_at, snod);            w->fix[w->n_fix++] = child_at;
  poke_u64(b->p + entry_at + 8, dset);        w->fix[w->n_fix++] = entry_at + 8;
  poke_u64(b->p + data_addr_at, w->off_data); w->fix[w->n_fix++] = data_addr_at;
  w->tmpl_no = no;
}

heroam_h5_writer *heroam_h5_create(const char *filename) {
  FILE *f = fopen(filename, "wb");
  if (!f) {
    h5fail("cannot create %s", filename);
    return NULL;
  }
  heroam_h5_writer *w = (heroam_h5_writer *)calloc(1, sizeof *w);
  if (!w) {
    fclose(f);
    return NULL;
  }
  w->f = f;
  setvbuf(f, NULL, _IOFBF, 1 << 20);
  w->mtime = (unsigned long)time(NULL);
  w->tmpl_no = 0xffffffffu;
  w->sorted_input = 1;
  /* superblock and root header are rewritten at close; reserve their space */
  unsigned char zero[SUPERBLOCK_BYTES + 40] = {0};
  if (fwrite(zero, 1, sizeof zero, f) != sizeof zero) w->failed = 1;
  w->eof = sizeof zero;
  return w;
}

int heroam_h5_save_timestep(heroam_h5_writer *w, unsigned int step, const heroam_particle *records, unsigned int no) {
  if (!w || w->failed) return -1;
  if (w->n_ent) {
    if (w->sorted_input && step > w->last_step) {
      /* the usual case: steps arrive in order */
    } else {
      w->sorted_input = 0;
      for (size_t i = 0; i < w->n_ent; ++i)
        if (w->ent[i].step == step) return h5fail("step %u written twice", step);
    }
  }
  if (no != w->tmpl_no) {
    build_step_template(w, no);
    if (w->tmpl.oom) {
      w->failed = 1;
      return h5fail("out of memory");
    }
  }
  const size_t data_bytes = (size_t)no * sizeof(heroam_particle);
  const size_t block = w->tmpl.n + ((data_bytes + 7) & ~(size_t)7);
  if (block > w->scratch_cap) {
    unsigned char *q = (unsigned char *)realloc(w->scratch, block);
    if (!q) return h5fail("out of memory");
    w->scratch = q;
    w->scratch_cap = block;
  }
  memcpy(w->scratch, w->tmpl.p, w->tmpl.n);
  for (int i = 0; i < w->n_fix; ++i)
    poke_u64(w->scratch + w->fix[i], peek_le(w->scratch + w->fix[i], 8) + w->eof);
  if (data_bytes) memcpy(w->scratch + w->tmpl.n, records, data_bytes);
  memset(w->scratch + w->tmpl.n + data_bytes, 0, block - w->tmpl.n - data_bytes);
  if (fwrite(w->scratch, 1, block, w->f) != block) {
    w->failed = 1;
    return h5fail("write failed");
  }
  if (w->n_ent == w->cap_ent) {
    const size_t cap = w->cap_ent ? w->cap_ent * 2 : 1024;
    root_entry *q = (root_entry *)realloc(w->ent, cap * sizeof *q);
    if (!q) return h5fail("out of memory");
    w->ent = q;
    w->cap_ent = cap;
  }
  root_entry *e = &w->ent[w->n_ent++];
  e->step = step;
  e->header = w->eof;
  e->btree = w->eof + w->off_btree;
  e->heap = w->eof + w->off_heap;
  e->name_off = 0;
  snprintf(e->name, sizeof e->name, "Step_%u", step); /* h5file.c:94-95 */
  w->last_step = step;
  w->eof += block;
  return 0;
}

static int by_name(const void *a, const void *b) {
  return strcmp(((const root_entry *)a)->name, ((const root_entry *)b)->name);
}

typedef struct {
  unsigned long long addr, max_key; /* a B-tree child and the heap offset of the largest name below it */
} child_ref;

int heroam_h5_close(heroam_h5_writer *w) {
  if (!w) return -1;
  int rc = w->failed ? -1 : 0;
  bytes b = {0, 0, 0, 0};
  /* root name heap: "" at offset 0, then the group names in B-tree (strcmp) order */
  qsort(w->ent, w->n_ent, sizeof *w->ent, by_name);
  bytes names = {0, 0, 0, 0};
  put_zero(&names, 8);
  for (size_t i = 0; i < w->n_ent; ++i) {
    w->ent[i].name_off = names.n;
    put(&names, w->ent[i].name, strlen(w->ent[i].name) + 1);
    pad8(&names);
  }
  size_t heap_data_at;
  const unsigned long long heap_addr = w->eof;
  local_heap_header(&b, names.n, &heap_data_at);
  if (!b.oom) poke_u64(b.p + heap_data_at, heap_addr + 32);
  put(&b, names.p, names.n);
  /* symbol nodes, 2*LEAF_K entries each */
  const size_t per_snod = 2 * LEAF_K, per_node = 2 * INTERNAL_K;
  size_t n_child = (w->n_ent + per_snod - 1) / per_snod;
  child_ref *child = (child_ref *)malloc((n_child ? n_child : 1) * sizeof *child);
  child_ref *parent = (child_ref *)malloc((n_child ? n_child : 1) * sizeof *parent);
  if (!child || !parent) rc = -1;
  for (size_t s = 0; s < n_child && rc == 0; ++s) {
    const size_t lo = s * per_snod, hi = lo + per_snod < w->n_ent ? lo + per_snod : w->n_ent;
    const size_t at = b.n;
    child[s].addr = heap_addr + at;
    child[s].max_key = w->ent[hi - 1].name_off;
    put(&b, "SNOD", 4);
    put_u8(&b, 1); put_u8(&b, 0); put_u16(&b, (unsigned)(hi - lo));
    for (size_t i = lo; i < hi; ++i)
      symbol_entry(&b, w->ent[i].name_off, w->ent[i].header, 1, w->ent[i].btree, w->ent[i].heap);
    put_zero(&b, at + SNOD_BYTES - b.n);
  }
  /* B-tree levels, bottom up, 2*INTERNAL_K children per node */
  unsigned long long root_btree = 0;
  unsigned level = 0;
  while (rc == 0) {
    const size_t n_nodes = n_child ? (n_child + per_node - 1) / per_node : 1;
    const unsigned long long first_addr = heap_addr + b.n;
    for (size_t nd = 0; nd < n_nodes; ++nd) {
      const size_t lo = nd * per_node, hi = lo + per_node < n_child ? lo + per_node : n_child;
      const size_t at = b.n;
      put(&b, "TREE", 4);
      put_u8(&b, 0); put_u8(&b, level); put_u16(&b, (unsigned)(hi - lo));
      put_u64(&b, nd ? first_addr + (nd - 1) * TREE_BYTES : H5_UNDEF);
      put_u64(&b, nd + 1 < n_nodes ? first_addr + (nd + 1) * TREE_BYTES : H5_UNDEF);
      put_u64(&b, lo ? child[lo - 1].max_key : 0);
      for (size_t c = lo; c < hi; ++c) {
        put_u64(&b, child[c].addr);
        put_u64(&b, child[c].max_key);
      }
      put_zero(&b, at + TREE_BYTES - b.n);
      parent[nd].addr = first_addr + nd * TREE_BYTES;
      parent[nd].max_key = hi > lo ? child[hi - 1].max_key : 0;
    }
    child_ref *t = child; child = parent; parent = t;
    n_child = n_nodes;
    ++level;
    if (n_nodes == 1) {
      root_btree = child[0].addr;
      break;
    }
  }
  free(child);
  free(parent);
  if (b.oom || names.oom) rc = -1;
  if (rc == 0 && fwrite(b.p, 1, b.n, w->f) != b.n) rc = -1;
  w->eof += b.n;
  /* superblock v0 + root group header at offset 0 */
  bytes s = {0, 0, 0, 0};
  static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', 0x0d, 0x0a, 0x1a, 0x0a};
  put(&s, sig, 8);
  put_u8(&s, 0); put_u8(&s, 0); put_u8(&s, 0); put_u8(&s, 0); /* superblock, free space, root group, reserved */
  put_u8(&s, 0); put_u8(&s, 8); put_u8(&s, 8); put_u8(&s, 0); /* shared header, offsets, lengths, reserved */
  put_u16(&s, LEAF_K); put_u16(&s, INTERNAL_K);
  put_u32(&s, 0);            /* consistency flags */
  put_u64(&s, 0);            /* base address */
  put_u64(&s, H5_UNDEF);     /* free-space info */
  put_u64(&s, w->eof);       /* end of file */
  put_u64(&s, H5_UNDEF);     /* driver info */
  symbol_entry(&s, 0, SUPERBLOCK_BYTES, 1, root_btree, heap_addr);
  size_t bt_at, hp_at;
  group_header(&s, &bt_at, &hp_at);
  if (!s.oom) {
    poke_u64(s.p + bt_at, root_btree);
    poke_u64(s.p + hp_at, heap_addr);
  }
  if (s.oom) rc = -1;
  if (rc == 0 && (fseek(w->f, 0, SEEK_SET) != 0 || fwrite(s.p, 1, s.n, w->f) != s.n)) rc = -1;
  if (fclose(w->f) != 0) rc = -1;
  if (rc) h5fail("could not finish the HDF5 file");
  free(s.p);
  free(b.p);
  free(names.p);
  free(w->tmpl.p);
  free(w->scratch);
  free(w->ent);
  free(w);
  return rc;
}

/* ------------------------------------------------------------------------------------------
 * reader
 * ------------------------------------------------------------------------------------------ */
struct heroam_h5_reader {
  FILE *f;
  unsigned long long base;
  long n_objects;
  unsigned long long *step_header; /* object header address of /Step_k, 0 = absent */
  size_t n_steps_cap;
  unsigned long long file_bytes;   /* sizes and counts read from the file are checked against this */
};

static int read_at(heroam_h5_reader *r, unsigned long long addr, void *dst, size_t n) {
  if (fseek(r->f, (long)(r->base + addr), SEEK_SET) != 0) return h5fail("seek to %llu failed", addr);
  if (fread(dst, 1, n, r->f) != n) return h5fail("short read of %zu bytes at %llu", n, addr);
  return 0;
}

/* calls fn(user, type, data, size) for every message of a version-1 object header, following
 * continuation blocks; fn returns non-zero to stop */
typedef int (*msg_fn)(void *user, unsigned type, const unsigned char *data, size_t size);
static int walk_header(heroam_h5_reader *r, unsigned long long addr, msg_fn fn, void *user) {
  unsigned char pre[16];
  if (read_at(r, addr, pre, 16)) return -1;
  if (pre[0] != 1) return h5fail("object header version %u at %llu is not supported", pre[0], addr);
  unsigned left = (unsigned)peek_le(pre + 2, 2);
  unsigned long long block_addr = addr + 16, block_size = peek_le(pre + 8, 4);
  unsigned long long cont_addr[16], cont_size[16];
  int n_cont = 0, at_cont = 0;
  for (;;) {
    if (block_size > r->file_bytes) return h5fail("header block of %llu bytes in a file of %llu", block_size, r->file_bytes);
    unsigned char *blk = (unsigned char *)malloc(block_size ? block_size : 1);
    if (!blk) return h5fail("out of memory");
    if (read_at(r, block_addr, blk, block_size)) { free(blk); return -1; }
    size_t pos = 0;
    while (left && pos + 8 <= block_size) {
      const unsigned type = (unsigned)peek_le(blk + pos, 2);
      const size_t size = (size_t)peek_le(blk + pos + 2, 2);
      if (pos + 8 + size > block_size) { free(blk); return h5fail("message overruns its header block"); }
      --left;
      if (type == 0x0010 && size >= 16 && n_cont < 16) {
        cont_addr[n_cont] = peek_le(blk + pos + 8, 8);
        cont_size[n_cont++] = peek_le(blk + pos + 16, 8);
      } else if (fn(user, type, blk + pos + 8, size)) {
        free(blk);
        return 0;
      }
      pos += 8 + size;
    }
    free(blk);
    if (!left || at_cont == n_cont) return 0;
    block_addr = cont_addr[at_cont];
    block_size = cont_size[at_cont++];
  }
}

typedef struct {
  unsigned long long btree, heap;
  int found;
} stab_ref;
static int take_stab(void *user, unsigned type, const unsigned char *data, size_t size) {
  stab_ref *s = (stab_ref *)user;
  if (type != 0x0011 || size < 16) return 0;
  s->btree = peek_le(data, 8);
  s->heap = peek_le(data + 8, 8);
  s->found = 1;
  return 1;
}

/* every (name, header address) of a symbol-table group */
typedef int (*entry_fn)(void *user, const char *name, unsigned long long header);
static int walk_tree(heroam_h5_reader *r, unsigned long long node, const char *heap, size_t heap_size, entry_fn fn,
                     void *user, int depth) {
  unsigned char head[24];
  if (depth > 32) return h5fail("B-tree too deep");
  if (read_at(r, node, head, 24)) return -1;
  if (memcmp(head, "TREE", 4) != 0 || head[4] != 0) return h5fail("no group B-tree node at %llu", node);
  const unsigned level = head[5], used = (unsigned)peek_le(head + 6, 2);
  unsigned char *body = (unsigned char *)malloc(16u * used + 8u);
  if (!body) return h5fail("out of memory");
  if (read_at(r, node + 24, body, 16u * used + 8u)) { free(body); return -1; }
  int rc = 0;
  for (unsigned i = 0; i < used && rc == 0; ++i) {
    const unsigned long long child = peek_le(body + 16u * i + 8u, 8);
    if (level > 0) {
      rc = walk_tree(r, child, heap, heap_size, fn, user, depth + 1);
      continue;
    }
    unsigned char sn[8];
    if (read_at(r, child, sn, 8)) { rc = -1; break; }
    if (memcmp(sn, "SNOD", 4) != 0) { rc = h5fail("no symbol node at %llu", child); break; }
    const unsigned count = (unsigned)peek_le(sn + 6, 2);
    unsigned char *ents = (unsigned char *)malloc(40u * count + 1u);
    if (!ents) { rc = h5fail("out of memory"); break; }
    if (read_at(r, child + 8, ents, 40u * count)) rc = -1;
    for (unsigned e = 0; e < count && rc == 0; ++e) {
      const unsigned long long name_off = peek_le(ents + 40u * e, 8);
      if (name_off >= heap_size || !memchr(heap + name_off, 0, heap_size - name_off)) {
        rc = h5fail("link name outside the heap");
        break;
      }
      rc = fn(user, heap + name_off, peek_le(ents + 40u * e + 8u, 8));
    }
    free(ents);
  }
  free(body);
  return rc;
}

static int walk_group(heroam_h5_reader *r, unsigned long long header, entry_fn fn, void *user) {
  stab_ref s = {0, 0, 0};
  if (walk_header(r, header, take_stab, &s)) return -1;
  if (!s.found) return h5fail("object at %llu is not a symbol-table group", header);
  unsigned char hh[32];
  if (read_at(r, s.heap, hh, 32)) return -1;
  if (memcmp(hh, "HEAP", 4) != 0) return h5fail("no local heap at %llu", s.heap);
  const size_t heap_size = (size_t)peek_le(hh + 8, 8);
  char *heap = (char *)malloc(heap_size + 1);
  if (!heap) return h5fail("out of memory");
  int rc = read_at(r, peek_le(hh + 24, 8), heap, heap_size);
  if (rc == 0) rc = walk_tree(r, s.btree, heap, heap_size, fn, user, 0);
  free(heap);
  return rc;
}

static int index_step(void *user, const char *name, unsigned long long header) {
  heroam_h5_reader *r = (heroam_h5_reader *)user;
  r->n_objects++;
  unsigned step;
  char tail;
  if (sscanf(name, "Step_%u%c", &step, &tail) != 1) return 0; /* not one of ours: counted, not indexed */
  /* every group costs well over 64 bytes of file: a step number beyond that is not from a writer of this format */
  if ((unsigned long long)step > r->file_bytes / 64u) return h5fail("group %s in a file of %llu bytes", name, r->file_bytes);
  if (step >= r->n_steps_cap) {
    size_t cap = r->n_steps_cap ? r->n_steps_cap : 1024;
    while (cap <= step) cap *= 2;
    unsigned long long *q = (unsigned long long *)realloc(r->step_header, cap * sizeof *q);
    if (!q) return h5fail("out of memory");
    memset(q + r->n_steps_cap, 0, (cap - r->n_steps_cap) * sizeof *q);
    r->step_header = q;
    r->n_steps_cap = cap;
  }
  r->step_header[step] = header;
  return 0;
}

heroam_h5_reader *heroam_h5_open(const char *filename) {
  FILE *f = fopen(filename, "rb");
  if (!f) {
    h5fail("cannot open %s", filename);
    return NULL;
  }
  heroam_h5_reader *r = (heroam_h5_reader *)calloc(1, sizeof *r);
  if (!r) {
    fclose(f);
    return NULL;
  }
  r->f = f;
  if (fseek(f, 0, SEEK_END) == 0 && ftell(f) > 0) r->file_bytes = (unsigned long long)ftell(f);
  static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', 0x0d, 0x0a, 0x1a, 0x0a};
  unsigned char sb[SUPERBLOCK_BYTES + 8];
  unsigned long long at = 0;
  int ok = 0;
  for (int probe = 0; probe < 24 && !ok; ++probe) { /* 0, 512, 1024, 2048, ... (a user block may precede) */
    if (fseek(f, (long)at, SEEK_SET) != 0 || fread(sb, 1, sizeof sb, f) < SUPERBLOCK_BYTES) break;
    if (memcmp(sb, sig, 8) == 0) ok = 1;
    else at = at ? at * 2 : 512;
  }
  if (!ok) { h5fail("%s: no HDF5 signature", filename); goto bad; }
  if (sb[8] > 1) { h5fail("superblock version %u is not supported (only 0 and 1)", sb[8]); goto bad; }
  if (sb[13] != 8 || sb[14] != 8) { h5fail("only 8-byte offsets and lengths are supported"); goto bad; }
  {
    const size_t fields = sb[8] == 1 ? 28 : 24; /* v1 adds indexed-storage K + 2 reserved bytes */
    r->base = peek_le(sb + fields, 8);
    const unsigned char *root = sb + fields + 32;
    const unsigned long long root_header = peek_le(root + 8, 8);
    if (r->base != at) r->base = at; /* addresses are relative to where the superblock really is */
    if (walk_group(r, root_header, index_step, r)) goto bad;
  }
  return r;
bad:
  heroam_h5_reader_close(r);
  return NULL;
}

long heroam_h5_num_steps(const heroam_h5_reader *r) { return r ? r->n_objects : -1; }

typedef struct {
  unsigned long long header;
} find_particles;
static int take_particles(void *user, const char *name, unsigned long long header) {
  if (strcmp(name, "particles") == 0) ((find_particles *)user)->header = header;
  return 0;
}

/* encoded length of a datatype message starting at p (classes used by Particle only), or 0 */
static size_t dt_length(const unsigned char *p, size_t avail, int depth) {
  if (avail < 8 || depth > 4) return 0;
  const unsigned cls = p[0] & 0x0f, ver = p[0] >> 4;
  if (cls == 0) return avail >= 12 ? 12 : 0;
  if (cls == 1) return avail >= 20 ? 20 : 0;
  if (cls == 10) {
    if (avail < 9) return 0;
    const unsigned nd = p[8];
    const size_t head = ver >= 3 ? 9 + 4u * nd : 12 + 8u * nd;
    if (avail < head) return 0;
    const size_t base = dt_length(p + head, avail - head, depth + 1);
    return base ? head + base : 0;
  }
  return 0;
}

typedef struct {
  unsigned long long count, data_addr, data_size;
  int have_space, have_type, have_layout, bad;
} dset_info;

static int check_particle_type(const unsigned char *p, size_t size) {
  static const struct { const char *name; unsigned off; } want[10] = {
      {"index", 0}, {"success", 4}, {"time", 8}, {"force_mag", 12}, {"velocity_sq", 16},
      {"orientation", 20}, {"position", 36}, {"velocity", 48}, {"ang_velocity", 60}, {"force", 72}};
  if (size < 8 || (p[0] & 0x0f) != 6) return h5fail("particles: not a compound datatype");
  const unsigned ver = p[0] >> 4, members = (unsigned)peek_le(p + 1, 2);
  if (peek_le(p + 4, 4) != sizeof(heroam_particle) || members != 10)
    return h5fail("particles: compound of %llu bytes / %u members, expected 84 / 10", peek_le(p + 4, 4), members);
  size_t pos = 8;
  for (unsigned m = 0; m < 10; ++m) {
    const unsigned char *nul = (const unsigned char *)memchr(p + pos, 0, size - pos);
    if (!nul) return h5fail("particles: truncated datatype");
    const size_t len = (size_t)(nul - (p + pos)) + 1;
    if (strcmp((const char *)p + pos, want[m].name) != 0) return h5fail("particles: member %u is '%s'", m, p + pos);
    pos += ver >= 3 ? len : (len + 7) & ~(size_t)7;
    unsigned off;
    if (ver >= 3) { off = p[pos]; pos += 1; }  /* 84-byte compound: one byte of offset */
    else { off = (unsigned)peek_le(p + pos, 4); pos += 4; }
    if (ver == 1) pos += 28;                   /* dimensionality, permutation, 4 dimension sizes */
    if (off != want[m].off) return h5fail("particles: member %s at offset %u", want[m].name, off);
    const size_t len_t = dt_length(p + pos, size - pos, 0);
    if (!len_t) return h5fail("particles: unsupported member datatype");
    pos += len_t;
  }
  return 0;
}

static int take_dataset(void *user, unsigned type, const unsigned char *d, size_t size) {
  dset_info *s = (dset_info *)user;
  if (type == 0x0001 && size >= 8) {
    const unsigned ver = d[0], rank = d[1];
    const size_t head = ver == 1 ? 8 : 4;
    if (rank != 1 || size < head + 8) { s->bad = h5fail("particles: dataspace of rank %u", rank); return 1; }
    s->count = peek_le(d + head, 8);
    s->have_space = 1;
  } else if (type == 0x0003) {
    if (check_particle_type(d, size)) { s->bad = -1; return 1; }
    s->have_type = 1;
  } else if (type == 0x0008 && size >= 2) {
    if (d[0] == 3 && d[1] == 1 && size >= 18) {
      s->data_addr = peek_le(d + 2, 8);
      s->data_size = peek_le(d + 10, 8);
      s->have_layout = 1;
    } else if ((d[0] == 1 || d[0] == 2) && d[2] == 1 && size >= 16) {
      s->data_addr = peek_le(d + 8, 8);
      s->data_size = 0; /* implied by the dataspace */
      s->have_layout = 1;
    } else {
      s->bad = h5fail("particles: only contiguous storage is supported (layout v%u class %u)", d[0], d[0] == 3 ? d[1] : d[2]);
      return 1;
    }
  }
  return 0;
}

int heroam_h5_read_timestep(heroam_h5_reader *r, unsigned int step, heroam_particles *out) {
  if (!r || !out) return h5fail("null argument");
  if (step >= r->n_steps_cap || !r->step_header[step]) return h5fail("no group /Step_%u", step);
  find_particles fp = {0};
  if (walk_group(r, r->step_header[step], take_particles, &fp)) return -1;
  if (!fp.header) return h5fail("/Step_%u has no dataset 'particles'", step);
  dset_info s;
  memset(&s, 0, sizeof s);
  if (walk_header(r, fp.header, take_dataset, &s)) return -1;
  if (s.bad) return -1;
  if (!s.have_space || !s.have_type || !s.have_layout) return h5fail("/Step_%u/particles: incomplete header", step);
  const size_t bytes_wanted = (size_t)s.count * sizeof(heroam_particle);
  if (s.data_size && s.data_size != bytes_wanted) return h5fail("/Step_%u/particles: storage size mismatch", step);
  out->no = (uint32_t)s.count;
  out->index = step; /* h5file.c:129 */
  out->particle = (heroam_particle *)malloc(bytes_wanted ? bytes_wanted : 1);
  if (!out->particle) return h5fail("Error allocating memory for particles");
  if (bytes_wanted && read_at(r, s.data_addr, out->particle, bytes_wanted)) {
    free(out->particle);
    out->particle = NULL;
    return -1;
  }
  return 0;
}

void heroam_h5_reader_close(heroam_h5_reader *r) {
  if (!r) return;
  if (r->f) fclose(r->f);
  free(r->step_header);
  free(r);
}
