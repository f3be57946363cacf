// heroam_gpu.cu -- C ABI (include/heroam_gpu.h) over the CUDA kernels.
// Host side of the engine: context, buffers, the pass loop, copies, timing.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <thread>
#include <vector>

#include "heroam_kernels.cuh"

using namespace heroam;

// ---------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

#define CU(call)                                                                        \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess)                                                              \
      return fail(HEROAM_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                  __FILE__, __LINE__);                                                  \
  } while (0)

// a device allocation that is released on every exit path of the function owning it
template <class T>
struct DevBuf {
  T *ptr = nullptr;
  ~DevBuf() { if (ptr) cudaFree(ptr); }
  cudaError_t alloc(size_t count) { return cudaMalloc(&ptr, count * sizeof(T)); }
};

// page-locked host memory with the same lifetime rule
template <class T>
struct PinnedBuf {
  T *ptr = nullptr;
  ~PinnedBuf() { if (ptr) cudaFreeHost(ptr); }
  cudaError_t alloc(size_t count) { return cudaMallocHost(&ptr, count * sizeof(T)); }
};

// ---------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------
struct heroam_ctx {
  int device = 0;
  heroam_config conf;
  heroam_options opt;
  cudaStream_t main = nullptr;
  cudaStream_t bin_stream[N_BINS] = {};
  cudaEvent_t bin_done[N_BINS] = {};
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_i0 = nullptr, ev_i1 = nullptr, ev_r = nullptr, ev_c0 = nullptr,
              ev_c1 = nullptr;
  float *d_table = nullptr;    // table_len + 1 floats: the table and a trailing 0 (the out-of-table bin)
  float *d_table_k = nullptr;  // the same multiplied by table_kappa (Turbo kernels)
  float table_kappa = 0.0f;
  long table_len = 0;
  // batch buffers (grown on demand)
  heroam_particle *d_particles = nullptr;
  size_t cap_particles = 0;
  TrajMeta *d_meta = nullptr;
  uint32_t *d_act = nullptr, *d_first = nullptr, *d_counts = nullptr, *d_bin_list = nullptr;
  uint32_t *d_live_list = nullptr;     // 2 x cap_traj: the live trajectories listed by the last two resolve passes
  uint8_t *d_nbr_pool = nullptr;       // partner lists of the warp kernels
  uint32_t *d_nbr_start = nullptr, *d_nbr_cnt = nullptr;
  unsigned long long *d_nbr_used = nullptr;
  unsigned long long nbr_cap = 0;
  float f_max = 0.0f;
  uint2 *d_free_list = nullptr;  // (trajectory, active slot) per particle in free flight, N_FREE_TIERS lists
  uint32_t free_stride = 0;      // entries per tier
  float4 *d_ckpt_q = nullptr;    // take-off orientation of loners
  heroam_traj_result *d_results = nullptr;
  size_t cap_traj = 0;
  uint32_t *d_bin_count = nullptr;
  Counters *d_counters = nullptr;
  uint32_t *h_bin_count = nullptr;  // pinned
  Counters *h_counters = nullptr;   // pinned
  // page-locked staging of heroam_gpu_simulate_batch (the reference's pointer-per-trajectory layout is
  // gathered into it, uploaded in one copy, and scattered back from it), grown on demand
  heroam_particle *h_stage = nullptr;
  size_t cap_stage = 0;
  // current batch
  unsigned long long he_number = 0;
  uint32_t n_traj = 0;
  size_t n_particles = 0;
  bool uploaded = false;
  heroam_run_info info;
  int sm_count = 0;
};

extern "C" int heroam_gpu_abi_version(void) { return HEROAM_ABI_VERSION; }

extern "C" int heroam_gpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" const char *heroam_gpu_last_error(void) { return g_err; }

// Steps until the reference's float clock reaches max_time, iterated exactly as the
// reference accumulates it (particle->time += dt, simulation.c:507).  Returns false when
// the clock stops advancing first: the reference would then loop forever (Q10).
static bool clock_steps(float dt, float max_time, uint32_t *steps_out, uint32_t *exact_out) {
  float t = 0.0f;
  uint64_t k = 0;
  bool exact = true;  // is the accumulated clock the product (float)k * dt all the way?
  while (t < max_time) {
    const float next = t + dt;
    if (!(next > t)) return false;
    t = next;
    if (++k >= 0xfffffff0ull) return false;
    exact = exact && k < (1ull << 24) && t == (float)k * dt;
  }
  *steps_out = (uint32_t)k;
  *exact_out = exact ? 1u : 0u;
  return true;
}

static int make_consts(const heroam_config &conf, unsigned long long he_number, Consts *out) {
  Consts c;
  c.dt = conf.dt;
  c.half_dt = (float)((double)conf.dt / 2.0);
  c.radius = (float)(2.22 * pow((double)(float)he_number, 1.0 / 3.0));
  c.acc_scale = (float)(1.0 / (double)conf.mass / (double)c.radius / (double)c.radius);
  c.inv_grid = (float)(1.0 / (double)conf.force_grid);
  c.grid_start = conf.force_grid_start;
  c.icd2 = conf.icd_dist * conf.icd_dist;
  c.max_time = conf.max_time;
  c.grid_len = conf.force_grid_length > 0x7fffffffL ? 0x7fffffff : (int)conf.force_grid_length;
  if (!(conf.dt > 0.0f)) return fail(HEROAM_E_ARG, "dt must be positive (got %g)", (double)conf.dt);
  if (!clock_steps(conf.dt, conf.max_time, &c.max_steps, &c.clock_exact))
    return fail(HEROAM_E_CLOCK,
                "dt=%g cannot advance a float clock up to max_time=%g: the reference's loop would never end",
                (double)conf.dt, (double)conf.max_time);
  {  // beyond this squared distance a pair is outside the table and outside icd_dist, with margin
    const double table_end = (double)conf.force_grid_start + ((double)c.grid_len + 2.0) * (double)conf.force_grid * 1.000001;
    const double reach2 = table_end > (double)c.icd2 ? table_end : (double)c.icd2;
    c.far2 = (float)(reach2 * 1.0001 + 1.0e-3);
  }
  c.kick = c.acc_scale * c.half_dt;
  c.kappa = c.kick * c.half_dt;
  c.inv_kappa = 1.0f / c.kappa;
  c.inv_half_dt = 1.0f / c.half_dt;
  c.pos_bias = -c.grid_start * c.inv_grid;
  c.inv_mr = (float)(1.0 / ((double)conf.mass * (double)c.radius));
  c.radius2 = 2.0f * c.radius;
  *out = c;
  return HEROAM_OK;
}

static int init_context(heroam_ctx *ctx, const float *force_list);

extern "C" int heroam_gpu_create(const heroam_config *conf, const float *force_list, int device,
                                 const heroam_options *opt, heroam_ctx **out) {
  if (!conf || !force_list || !out) return fail(HEROAM_E_ARG, "heroam_gpu_create: null argument");
  if (conf->force_grid_length <= 0) return fail(HEROAM_E_ARG, "force_grid_length must be positive");
  if (!(conf->force_grid > 0.0f)) return fail(HEROAM_E_ARG, "force_grid must be positive");
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (ndev <= 0) return fail(HEROAM_E_CUDA, "no CUDA device: this engine has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(HEROAM_E_ARG, "device %d out of range (0..%d)", device, ndev - 1);
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(HEROAM_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                prop.major, prop.minor);
  heroam_options options;
  memset(&options, 0, sizeof options);
  if (opt) options = *opt;
  if (options.mode != HEROAM_MODE_EXACT && options.mode != HEROAM_MODE_FAST && options.mode != HEROAM_MODE_FAST_PLAIN &&
      options.mode != HEROAM_MODE_FAST_RELAXED && options.mode != HEROAM_MODE_FAST_CLOSED)
    return fail(HEROAM_E_ARG, "unknown mode %d", options.mode);
  heroam_ctx *ctx = new heroam_ctx();
  ctx->device = device;
  ctx->conf = *conf;
  // seed = 0: the reference seeds from the clock (simulation.c:283-286).  Resolved once per context; a
  // caller driving several contexts as one run resolves it itself and passes the same seed to all
  // (host/heroam_main.c does).
  if (ctx->conf.seed == 0) ctx->conf.seed = (long)time(nullptr);
  ctx->opt = options;
  ctx->sm_count = prop.multiProcessorCount;
  const int rc = init_context(ctx, force_list);
  if (rc) {  // whatever was created so far is released; the error message stays
    heroam_gpu_destroy(ctx);
    return rc;
  }
  *out = ctx;
  return HEROAM_OK;
}

// streams, events, the force table and the small fixed buffers of a context
static int init_context(heroam_ctx *ctx, const float *force_list) {
  const heroam_config *conf = &ctx->conf;
  if (ctx->opt.segment_steps == 0) ctx->opt.segment_steps = 8192;
  memset(&ctx->info, 0, sizeof ctx->info);
  ctx->table_len = conf->force_grid_length;
  CU(cudaStreamCreateWithFlags(&ctx->main, cudaStreamNonBlocking));
  for (int b = 0; b < N_BINS; ++b) {
    CU(cudaStreamCreateWithFlags(&ctx->bin_stream[b], cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&ctx->bin_done[b], cudaEventDisableTiming));
  }
  CU(cudaEventCreate(&ctx->ev_start));
  CU(cudaEventCreate(&ctx->ev_stop));
  CU(cudaEventCreate(&ctx->ev_i0));
  CU(cudaEventCreate(&ctx->ev_i1));
  CU(cudaEventCreate(&ctx->ev_r));
  CU(cudaEventCreate(&ctx->ev_c0));
  CU(cudaEventCreate(&ctx->ev_c1));
  CU(cudaMalloc(&ctx->d_table, ((size_t)ctx->table_len + 1) * sizeof(float)));
  CU(cudaMalloc(&ctx->d_table_k, ((size_t)ctx->table_len + 1) * sizeof(float)));
  CU(cudaMemset(ctx->d_table, 0, ((size_t)ctx->table_len + 1) * sizeof(float)));
  CU(cudaMemcpy(ctx->d_table, force_list, (size_t)ctx->table_len * sizeof(float), cudaMemcpyHostToDevice));
  {  // largest force magnitude a pair that has not met can feel (squared distance >= icd^2)
    const double icd2 = (double)conf->icd_dist * (double)conf->icd_dist;
    long i0 = (long)floor((icd2 - (double)conf->force_grid_start) / (double)conf->force_grid) - 1;
    if (i0 < 0) i0 = 0;
    float fm = 0.0f;
    for (long i = i0; i < ctx->table_len; ++i) fm = fmaxf(fm, fabsf(force_list[i]));
    ctx->f_max = fm;
  }
  ctx->nbr_cap = 64ull << 20;
  CU(cudaMalloc(&ctx->d_nbr_pool, ctx->nbr_cap));
  CU(cudaMalloc(&ctx->d_nbr_used, sizeof(unsigned long long)));
  CU(cudaMalloc(&ctx->d_bin_count, (size_t)MAX_PARTS * N_BIN_COUNTERS * sizeof(uint32_t)));
  CU(cudaMalloc(&ctx->d_counters, sizeof(Counters)));
  CU(cudaMallocHost(&ctx->h_bin_count, (size_t)MAX_PARTS * N_BIN_COUNTERS * sizeof(uint32_t)));
  CU(cudaMallocHost(&ctx->h_counters, sizeof(Counters)));
  return HEROAM_OK;
}

extern "C" void heroam_gpu_destroy(heroam_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  cudaFree(ctx->d_table);
  cudaFree(ctx->d_table_k);
  cudaFree(ctx->d_nbr_pool);
  cudaFree(ctx->d_nbr_start);
  cudaFree(ctx->d_nbr_cnt);
  cudaFree(ctx->d_nbr_used);
  cudaFree(ctx->d_particles);
  cudaFree(ctx->d_meta);
  cudaFree(ctx->d_act);
  cudaFree(ctx->d_free_list);
  cudaFree(ctx->d_ckpt_q);
  cudaFree(ctx->d_first);
  cudaFree(ctx->d_counts);
  cudaFree(ctx->d_bin_list);
  cudaFree(ctx->d_live_list);
  cudaFree(ctx->d_results);
  cudaFree(ctx->d_bin_count);
  cudaFree(ctx->d_counters);
  cudaFreeHost(ctx->h_bin_count);
  cudaFreeHost(ctx->h_counters);
  cudaFreeHost(ctx->h_stage);
  for (int b = 0; b < N_BINS; ++b) {
    if (ctx->bin_stream[b]) cudaStreamDestroy(ctx->bin_stream[b]);
    if (ctx->bin_done[b]) cudaEventDestroy(ctx->bin_done[b]);
  }
  cudaEvent_t evs[] = {ctx->ev_start, ctx->ev_stop, ctx->ev_i0, ctx->ev_i1, ctx->ev_r, ctx->ev_c0, ctx->ev_c1};
  for (cudaEvent_t e : evs)
    if (e) cudaEventDestroy(e);
  if (ctx->main) cudaStreamDestroy(ctx->main);
  delete ctx;
}

extern "C" int heroam_gpu_set_mode(heroam_ctx *ctx, int mode) {
  if (!ctx) return fail(HEROAM_E_ARG, "null context");
  if (mode != HEROAM_MODE_EXACT && mode != HEROAM_MODE_FAST && mode != HEROAM_MODE_FAST_PLAIN &&
      mode != HEROAM_MODE_FAST_RELAXED && mode != HEROAM_MODE_FAST_CLOSED)
    return fail(HEROAM_E_ARG, "unknown mode %d", mode);
  ctx->opt.mode = mode;
  return HEROAM_OK;
}

// grow the per-batch device buffers
static int reserve(heroam_ctx *ctx, size_t n_traj, size_t n_particles) {
  if (n_particles > ctx->cap_particles) {
    cudaFree(ctx->d_particles);
    cudaFree(ctx->d_act);
    ctx->d_particles = nullptr;
    ctx->d_act = nullptr;
    ctx->cap_particles = 0;
    cudaFree(ctx->d_nbr_start);
    cudaFree(ctx->d_nbr_cnt);
    cudaFree(ctx->d_ckpt_q);
    cudaFree(ctx->d_free_list);
    ctx->d_nbr_start = ctx->d_nbr_cnt = nullptr;
    ctx->d_ckpt_q = nullptr;
    ctx->d_free_list = nullptr;
    CU(cudaMalloc(&ctx->d_particles, n_particles * sizeof(heroam_particle)));
    CU(cudaMalloc(&ctx->d_act, n_particles * sizeof(uint32_t)));
    CU(cudaMalloc(&ctx->d_nbr_start, n_particles * sizeof(uint32_t)));
    CU(cudaMalloc(&ctx->d_nbr_cnt, n_particles * sizeof(uint32_t)));
    CU(cudaMalloc(&ctx->d_ckpt_q, n_particles * sizeof(float4)));
    // every moving particle can be in flight at once: one entry each, per tier
    ctx->free_stride = (uint32_t)(n_particles > 0xffffffffull ? 0xffffffffull : n_particles);
    CU(cudaMalloc(&ctx->d_free_list, (size_t)N_FREE_TIERS * ctx->free_stride * sizeof(uint2)));
    ctx->cap_particles = n_particles;
  }
  if (n_traj > ctx->cap_traj) {
    cudaFree(ctx->d_meta);
    cudaFree(ctx->d_first);
    cudaFree(ctx->d_counts);
    cudaFree(ctx->d_bin_list);
    cudaFree(ctx->d_live_list);
    cudaFree(ctx->d_results);
    ctx->d_meta = nullptr;
    ctx->d_first = ctx->d_counts = ctx->d_bin_list = ctx->d_live_list = nullptr;
    ctx->d_results = nullptr;
    ctx->cap_traj = 0;
    CU(cudaMalloc(&ctx->d_meta, n_traj * sizeof(TrajMeta)));
    CU(cudaMalloc(&ctx->d_first, n_traj * sizeof(uint32_t)));
    CU(cudaMalloc(&ctx->d_counts, n_traj * sizeof(uint32_t)));
    CU(cudaMalloc(&ctx->d_bin_list, (size_t)N_BINS * n_traj * sizeof(uint32_t)));
    CU(cudaMalloc(&ctx->d_live_list, 2 * n_traj * sizeof(uint32_t)));
    CU(cudaMalloc(&ctx->d_results, n_traj * sizeof(heroam_traj_result)));
    ctx->cap_traj = n_traj;
  }
  return HEROAM_OK;
}

static DevView make_view(heroam_ctx *ctx, const Consts &c) {
  DevView v;
  v.c = c;
  v.c.f_max = ctx->f_max;
  v.table = ctx->d_table;
  v.table_k = ctx->d_table_k;
  v.particles = ctx->d_particles;
  v.meta = ctx->d_meta;
  v.act = ctx->d_act;
  v.ckpt_q = ctx->d_ckpt_q;
  v.nbr_pool = ctx->d_nbr_pool;
  v.nbr_start = ctx->d_nbr_start;
  v.nbr_cnt = ctx->d_nbr_cnt;
  v.nbr_used = ctx->d_nbr_used;
  v.nbr_cap = ctx->nbr_cap;
  v.counters = ctx->d_counters;
  v.n_traj = ctx->n_traj;
  return v;
}

extern "C" int heroam_gpu_upload(heroam_ctx *ctx, unsigned long long he_number, const uint32_t *counts,
                                 const heroam_particle *flat, int count) {
  if (!ctx || !counts || (!flat && count > 0) || count < 0) return fail(HEROAM_E_ARG, "heroam_gpu_upload: bad argument");
  CU(cudaSetDevice(ctx->device));
  std::vector<uint32_t> first((size_t)count);
  size_t total = 0;
  for (int i = 0; i < count; ++i) {
    if (counts[i] == 0) return fail(HEROAM_E_ARG, "trajectory %d has no particles", i);
    first[i] = (uint32_t)total;
    total += counts[i];
    if (total > 0xfffffff0ull) return fail(HEROAM_E_ARG, "batch too large: more than 2^32 particle records");
  }
  int rc = reserve(ctx, (size_t)count > 0 ? (size_t)count : 1, total > 0 ? total : 1);
  if (rc) return rc;
  ctx->he_number = he_number;
  ctx->n_traj = (uint32_t)count;
  ctx->n_particles = total;
  memset(&ctx->info, 0, sizeof ctx->info);
  CU(cudaEventRecord(ctx->ev_c0, ctx->main));
  if (count > 0) {
    CU(cudaMemcpyAsync(ctx->d_particles, flat, total * sizeof(heroam_particle), cudaMemcpyHostToDevice, ctx->main));
    CU(cudaMemcpyAsync(ctx->d_counts, counts, (size_t)count * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->main));
    CU(cudaMemcpyAsync(ctx->d_first, first.data(), (size_t)count * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->main));
  }
  CU(cudaEventRecord(ctx->ev_c1, ctx->main));
  CU(cudaStreamSynchronize(ctx->main));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, ctx->ev_c0, ctx->ev_c1));
  ctx->info.h2d_ms = ms;
  ctx->info.h2d_bytes = total * sizeof(heroam_particle) + 2ull * count * sizeof(uint32_t);
  ctx->uploaded = true;
  return HEROAM_OK;
}

// Longest window for which the acceleration part of the reach bound stays around 2 A per linked
// partner: R dt^2 f_max / (m R) W^2 / 2 <= 2  =>  W <= 2 sqrt(m / f_max) / dt.
static uint32_t split_window_max(const heroam_ctx *ctx) {
  if (!(ctx->f_max > 0.0f)) return 0xffffffffu;
  const double w = 2.0 * sqrt((double)ctx->conf.mass / (double)ctx->f_max) / (double)ctx->conf.dt;
  return w > 4.0e9 ? 0xffffffffu : (uint32_t)w;
}

// the kappa-scaled copy of the table the Turbo kernels gather from
static int refresh_scaled_table(heroam_ctx *ctx, const Consts &c) {
  if (ctx->table_kappa == c.kappa) return HEROAM_OK;
  const unsigned n = (unsigned)ctx->table_len + 1u;
  k_scale_table<<<(n + 255u) / 256u, 256, 0, ctx->main>>>(ctx->d_table_k, ctx->d_table, n, c.kappa);
  CU(cudaGetLastError());
  ctx->table_kappa = c.kappa;
  ctx->info.kernel_launches++;
  return HEROAM_OK;
}

// How a run is being scheduled at the moment.
struct Regime {
  bool draining = false;      // fewer trajectories in the full kernels than fill the GPU: latency-bound segment table
};

// What a pass is launched with under regime rg.
static Segments pass_segments(const heroam_ctx *ctx, const Regime &rg) {
  Segments sg = make_segments(ctx->opt.segment_steps, rg.draining, split_window_max(ctx));
  sg.free_stride = ctx->free_stride;
  sg.no_flight = ctx->opt.no_flight ? 1u : 0u;
  return sg;
}

// Decide the regime of the next pass from the populations the resolve kernel just reported: fewer
// trajectories in the full kernels than fill the GPU once means latency-bound from here on.
static void update_regime(const heroam_ctx *ctx, Regime &rg, const uint64_t *bin_total = nullptr) {
  uint64_t in_kernels = 0;
  for (int b = 1; b < N_BINS; ++b)
    if (b < BIN_FREE || b >= BIN_CORE) in_kernels += bin_total ? bin_total[b] : ctx->h_bin_count[b];
  rg.draining = in_kernels < (uint64_t)ctx->sm_count * 768;
}

// P: arithmetic of the integrate kernels, F: of the free flights
template <class P, class F>
static int launch_bin(heroam_ctx *ctx, const DevView &v, int bin, uint32_t cnt, bool draining = false,
                      uint32_t list_offset = 0, uint32_t free_offset = 0) {
  const uint32_t *list = ctx->d_bin_list + (size_t)bin * ctx->n_traj + list_offset;
  cudaStream_t s = ctx->bin_stream[bin];
  const unsigned grid = (cnt + INTEGRATE_TPB - 1) / INTEGRATE_TPB;
  if (bin >= BIN_FREE && bin < BIN_CORE) {
    const uint2 *flyers = ctx->d_free_list + (size_t)(bin - BIN_FREE) * ctx->free_stride + free_offset;
    if (v.c.clock_exact) k_free_flight<F, true><<<(cnt + 127) / 128, 128, 0, s>>>(v, flyers, cnt);
    else k_free_flight<F, false><<<(cnt + 127) / 128, 128, 0, s>>>(v, flyers, cnt);
    CU(cudaGetLastError());
    return HEROAM_OK;
  }
  const int width = bin >= BIN_CORE ? bin - BIN_CORE : bin;  // the core of a split trajectory runs in the kernel of its size
  static const bool lookahead_on = !(getenv("HEROAM_LOOKAHEAD") && atoi(getenv("HEROAM_LOOKAHEAD")) == 0);
  // latency-bound: prefetch the table two steps ahead (while the GPU is full the prefetch costs more than it
  // hides: 4e6 trajectories 24.1-24.4 s with it on for the narrow bins, 23.1 s without)
  if (draining && lookahead_on && width >= 2 && width <= 4) {
    if (width == 2) k_integrate_thread<2, P, true><<<grid, INTEGRATE_TPB, 0, s>>>(v, list, cnt);
    else if (width == 3) k_integrate_thread<3, P, true><<<grid, INTEGRATE_TPB, 0, s>>>(v, list, cnt);
    else k_integrate_thread<4, P, true><<<grid, INTEGRATE_TPB, 0, s>>>(v, list, cnt);
    CU(cudaGetLastError());
    return HEROAM_OK;
  }
  switch (width) {
    case 1: k_integrate_thread<1, P><<<grid, INTEGRATE_TPB, 0, s>>>(v, list, cnt); break;
    case 2: k_integrate_thread<2, P><<<grid, INTEGRATE_TPB, 0, s>>>(v, list, cnt); break;
    case 3: k_integrate_thread<3, P><<<grid, INTEGRATE_TPB, 0, s>>>(v, list, cnt); break;
    case 4: k_integrate_thread<4, P><<<grid, INTEGRATE_TPB, 0, s>>>(v, list, cnt); break;
    case 5: k_integrate_thread<5, P><<<grid, INTEGRATE_TPB, 0, s>>>(v, list, cnt); break;
    case 6: k_integrate_thread<6, P><<<grid, INTEGRATE_TPB, 0, s>>>(v, list, cnt); break;
    case 7: k_integrate_thread<7, P><<<grid, INTEGRATE_TPB, 0, s>>>(v, list, cnt); break;
    case 8: k_integrate_thread<8, P><<<grid, INTEGRATE_TPB, 0, s>>>(v, list, cnt); break;
    default: {
      const unsigned wgrid = (cnt + WARP_KERNEL_WARPS - 1) / WARP_KERNEL_WARPS;
      const unsigned sgrid = (cnt + SMEM_TPB - 1) / SMEM_TPB;
      if (ctx->opt.wide_kernel == 1) {
        k_integrate_records<P><<<(cnt + 63) / 64, 64, 0, s>>>(v, list, cnt);
#define HEROAM_HYBRID_CASE(A)                                                                                     \
  } else if (bin == A && ctx->opt.wide_kernel == 0) {                                                             \
    CU(cudaFuncSetAttribute(k_integrate_hybrid<A, P>, cudaFuncAttributeMaxDynamicSharedMemorySize,                \
                            (int)hybrid_smem_bytes<A>()));                                                        \
    k_integrate_hybrid<A, P><<<sgrid, SMEM_TPB, hybrid_smem_bytes<A>(), s>>>(v, list, cnt);
      HEROAM_HYBRID_CASE(9)
      HEROAM_HYBRID_CASE(10)
      HEROAM_HYBRID_CASE(11)
      HEROAM_HYBRID_CASE(12)
#undef HEROAM_HYBRID_CASE
      } else if (bin <= BIN_SMEM16 && ctx->opt.wide_kernel != 2) {
        const size_t bytes = 5ull * 16 * SMEM_TPB * sizeof(float4);
        CU(cudaFuncSetAttribute(k_integrate_smem<16, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        k_integrate_smem<16, P><<<sgrid, SMEM_TPB, bytes, s>>>(v, list, cnt);
      } else if (bin <= BIN_WARP1) {
        // partner lists exist for the warp bins proper (schedule_segment), not for narrower bins
        k_integrate_warp<1, P><<<wgrid, 32 * WARP_KERNEL_WARPS, 0, s>>>(v, list, cnt, bin == BIN_WARP1);
      } else if (bin == BIN_WARP2) {
        k_integrate_warp<2, P><<<wgrid, 32 * WARP_KERNEL_WARPS, 0, s>>>(v, list, cnt, true);
      } else {
        k_integrate_warp<4, P><<<wgrid, 32 * WARP_KERNEL_WARPS, 0, s>>>(v, list, cnt, true);
      }
      break;
    }
  }
  CU(cudaGetLastError());
  return HEROAM_OK;
}

// R: arithmetic of the once-per-pass resolve kernel; H: arithmetic of the integrate kernels; F: of the free flights
template <class R, class H, class F = H>
static int run_passes(heroam_ctx *ctx, const Consts &c) {
  DevView v = make_view(ctx, c);
  if (H::turbo) {
    int rc = refresh_scaled_table(ctx, c);
    if (rc) return rc;
  }
  const uint32_t n = ctx->n_traj;
  const unsigned tgrid = (n + 127) / 128;
  const uint32_t cap = ctx->opt.max_steps == 0 || ctx->opt.max_steps > 0xffffffffull ? 0xffffffffu
                                                                                    : (uint32_t)ctx->opt.max_steps;
  double integrate_ms = 0.0;
  bool pending_timing = false;
  Regime rg;
  LiveLists lists = {nullptr, n, ctx->d_live_list};
  CU(cudaEventRecord(ctx->ev_start, ctx->main));
  CU(cudaMemsetAsync(ctx->d_counters, 0, sizeof(Counters), ctx->main));
  k_prepare<<<tgrid, 128, 0, ctx->main>>>(v, ctx->d_first, ctx->d_counts);
  CU(cudaGetLastError());
  ctx->info.kernel_launches++;
  for (;;) {
    CU(cudaMemsetAsync(ctx->d_bin_count, 0, N_BIN_COUNTERS * sizeof(uint32_t), ctx->main));
    CU(cudaMemsetAsync(ctx->d_nbr_used, 0, sizeof(unsigned long long), ctx->main));
    const Segments sg = pass_segments(ctx, rg);
    if (lists.count)
      k_resolve_and_bin<R, F><<<(lists.count + 127) / 128, 128, 0, ctx->main>>>(v, sg, cap, ctx->d_bin_count, ctx->d_bin_list,
                                                                               ctx->d_free_list, lists);
    CU(cudaGetLastError());
    ctx->info.kernel_launches++;
    CU(cudaMemcpyAsync(ctx->h_bin_count, ctx->d_bin_count, N_BIN_COUNTERS * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->main));
    CU(cudaEventRecord(ctx->ev_r, ctx->main));
    CU(cudaStreamSynchronize(ctx->main));
    // from the second pass on only the trajectories still live are visited
    lists.in = lists.out;
    lists.count = ctx->h_bin_count[LIVE_SLOT];
    lists.out = lists.in == ctx->d_live_list ? ctx->d_live_list + ctx->cap_traj : ctx->d_live_list;
    if (pending_timing) {
      float ms = 0;
      CU(cudaEventElapsedTime(&ms, ctx->ev_i0, ctx->ev_i1));
      integrate_ms += ms;
      CU(cudaEventElapsedTime(&ms, ctx->ev_i1, ctx->ev_r));
      ctx->info.resolve_ms += ms;
      pending_timing = false;
    }
    uint64_t live = 0;
    for (int b = 0; b < N_BINS; ++b) live += ctx->h_bin_count[b];
    if (live == 0) break;
    update_regime(ctx, rg);
    ctx->info.passes++;
    CU(cudaEventRecord(ctx->ev_i0, ctx->main));
    for (int b = 1; b < N_BINS; ++b) {
      const uint32_t cnt = ctx->h_bin_count[b];
      if (!cnt) continue;
      // the bin streams start after ev_i0, i.e. after this pass's resolve kernel
      CU(cudaStreamWaitEvent(ctx->bin_stream[b], ctx->ev_i0, 0));
      int rc = launch_bin<H, F>(ctx, v, b, cnt, rg.draining);
      if (rc) return rc;
      ctx->info.kernel_launches++;
      CU(cudaEventRecord(ctx->bin_done[b], ctx->bin_stream[b]));
      CU(cudaStreamWaitEvent(ctx->main, ctx->bin_done[b], 0));
    }
    CU(cudaEventRecord(ctx->ev_i1, ctx->main));
    pending_timing = true;
  }
  k_results<<<tgrid, 128, 0, ctx->main>>>(v, ctx->d_results);
  CU(cudaGetLastError());
  ctx->info.kernel_launches++;
  CU(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->main));
  CU(cudaEventRecord(ctx->ev_stop, ctx->main));
  CU(cudaStreamSynchronize(ctx->main));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_stop));
  ctx->info.device_ms = ms;
  ctx->info.integrate_ms = integrate_ms;
  ctx->info.particle_steps = ctx->h_counters->particle_steps;
  ctx->info.pair_steps = ctx->h_counters->pair_steps;
  ctx->info.inrange_steps = ctx->h_counters->inrange_steps;
  ctx->info.pair_skipped = ctx->h_counters->pair_skipped;
  ctx->info.free_steps = ctx->h_counters->free_steps;
  ctx->info.trajectories = n;
  ctx->info.particles = ctx->n_particles;
  if (ctx->h_counters->unsupported)
    return fail(HEROAM_E_UNSUPPORTED, "%llu trajectories have more than %d moving particles", ctx->h_counters->unsupported,
                MAX_ACTIVE);
  return HEROAM_OK;
}

extern "C" int heroam_gpu_run(heroam_ctx *ctx) {
  if (!ctx) return fail(HEROAM_E_ARG, "null context");
  if (!ctx->uploaded) return fail(HEROAM_E_ARG, "heroam_gpu_run: nothing uploaded");
  CU(cudaSetDevice(ctx->device));
  Consts c;
  int rc = make_consts(ctx->conf, ctx->he_number, &c);
  if (rc) return rc;
  ctx->uploaded = false;  // the resident batch is consumed
  if (ctx->n_traj == 0) return HEROAM_OK;
  rc = ctx->opt.mode == HEROAM_MODE_EXACT          ? run_passes<Exact, Exact>(ctx, c)
       : ctx->opt.mode == HEROAM_MODE_FAST_PLAIN   ? run_passes<Fast, Fast>(ctx, c)
       : ctx->opt.mode == HEROAM_MODE_FAST_RELAXED ? run_passes<Fast, Turbo8>(ctx, c)
       : ctx->opt.mode == HEROAM_MODE_FAST_CLOSED  ? run_passes<Fast, Turbo, TurboClosed>(ctx, c)
       : ctx->opt.count_work                     ? run_passes<Fast, TurboCount>(ctx, c)
                                                 : run_passes<Fast, Turbo>(ctx, c);
  return rc;
}

extern "C" int heroam_gpu_download(heroam_ctx *ctx, heroam_particle *flat, heroam_traj_result *results) {
  if (!ctx) return fail(HEROAM_E_ARG, "null context");
  CU(cudaSetDevice(ctx->device));
  CU(cudaEventRecord(ctx->ev_c0, ctx->main));
  size_t bytes = 0;
  if (flat && ctx->n_particles) {
    CU(cudaMemcpyAsync(flat, ctx->d_particles, ctx->n_particles * sizeof(heroam_particle), cudaMemcpyDeviceToHost, ctx->main));
    bytes += ctx->n_particles * sizeof(heroam_particle);
  }
  if (results && ctx->n_traj) {
    CU(cudaMemcpyAsync(results, ctx->d_results, (size_t)ctx->n_traj * sizeof(heroam_traj_result), cudaMemcpyDeviceToHost, ctx->main));
    bytes += (size_t)ctx->n_traj * sizeof(heroam_traj_result);
  }
  CU(cudaEventRecord(ctx->ev_c1, ctx->main));
  CU(cudaStreamSynchronize(ctx->main));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, ctx->ev_c0, ctx->ev_c1));
  ctx->info.d2h_ms = ms;
  ctx->info.d2h_bytes = bytes;
  return HEROAM_OK;
}

extern "C" int heroam_gpu_simulate_flat(heroam_ctx *ctx, unsigned long long he_number, const uint32_t *counts,
                                        heroam_particle *flat, int count, heroam_traj_result *results) {
  int rc = heroam_gpu_upload(ctx, he_number, counts, flat, count);
  if (rc) return rc;
  rc = heroam_gpu_run(ctx);
  if (rc) return rc;
  return heroam_gpu_download(ctx, flat, results);
}

// records of trajectories [lo, hi) between the caller's per-trajectory arrays and the flat staging buffer
static void stage_copy(heroam_particles *trajectories, const std::vector<size_t> &first, heroam_particle *stage, int lo,
                       int hi, bool gather) {
  for (int i = lo; i < hi; ++i) {
    const size_t bytes = (size_t)trajectories[i].no * sizeof(heroam_particle);
    if (gather) memcpy(stage + first[i], trajectories[i].particle, bytes);
    else memcpy(trajectories[i].particle, stage + first[i], bytes);
  }
}

static void stage_parallel(heroam_particles *trajectories, const std::vector<size_t> &first, heroam_particle *stage,
                           int count, bool gather) {
  // a copy of tens of bytes per trajectory is bound by memory latency: a few host threads for large batches
  unsigned hw = std::thread::hardware_concurrency();
  int n_threads = count < 65536 ? 1 : (int)(hw ? (hw > 16 ? 16 : hw) : 4);
  if (n_threads <= 1) {
    stage_copy(trajectories, first, stage, 0, count, gather);
    return;
  }
  std::vector<std::thread> pool;
  for (int t = 0; t < n_threads; ++t) {
    const int lo = (int)((long long)count * t / n_threads), hi = (int)((long long)count * (t + 1) / n_threads);
    pool.emplace_back(stage_copy, trajectories, std::cref(first), stage, lo, hi, gather);
  }
  for (auto &th : pool) th.join();
}

static int reserve_stage(heroam_ctx *ctx, size_t records) {
  if (records <= ctx->cap_stage) return HEROAM_OK;
  cudaFreeHost(ctx->h_stage);
  ctx->h_stage = nullptr;
  ctx->cap_stage = 0;
  CU(cudaMallocHost(&ctx->h_stage, records * sizeof(heroam_particle)));
  ctx->cap_stage = records;
  return HEROAM_OK;
}

extern "C" int heroam_gpu_reserve(heroam_ctx *ctx, unsigned long long n_trajectories, unsigned long long n_records) {
  if (!ctx) return fail(HEROAM_E_ARG, "null context");
  if (n_records > 0xfffffff0ull || n_trajectories > 0xfffffff0ull) return fail(HEROAM_E_ARG, "batch too large: more than 2^32 records");
  CU(cudaSetDevice(ctx->device));
  int rc = reserve(ctx, n_trajectories ? (size_t)n_trajectories : 1, n_records ? (size_t)n_records : 1);
  if (rc) return rc;
  return reserve_stage(ctx, (size_t)n_records);
}

extern "C" int heroam_gpu_simulate_batch(heroam_ctx *ctx, unsigned long long he_number,
                                         heroam_particles *trajectories, int count, heroam_traj_result *results) {
  if (!ctx || (!trajectories && count > 0) || count < 0) return fail(HEROAM_E_ARG, "heroam_gpu_simulate_batch: bad argument");
  CU(cudaSetDevice(ctx->device));
  std::vector<uint32_t> counts((size_t)count);
  std::vector<size_t> first((size_t)count);
  size_t total = 0;
  for (int i = 0; i < count; ++i) {
    if (!trajectories[i].particle || trajectories[i].no == 0)
      return fail(HEROAM_E_ARG, "trajectory %d is empty", i);
    counts[i] = trajectories[i].no;
    first[i] = total;
    total += counts[i];
  }
  int rc = reserve_stage(ctx, total);
  if (rc) return rc;
  stage_parallel(trajectories, first, ctx->h_stage, count, true);
  rc = heroam_gpu_simulate_flat(ctx, he_number, counts.data(), ctx->h_stage, count, results);
  if (rc) return rc;
  stage_parallel(trajectories, first, ctx->h_stage, count, false);
  return HEROAM_OK;
}

// ---------------------------------------------------------------------------------
// Per-step capture (saveflag = 1, simulation.c:460-476, :522-523).  The batch advances in
// chunks of `chunk` loop iterations; k_integrate_capture leaves the snapshots of a chunk in a
// device buffer, which is copied to page-locked host memory on a second stream and handed to
// the sink while the next chunk is being integrated (two buffers).
// ---------------------------------------------------------------------------------
struct CaptureBuffers {
  DevBuf<heroam_particle> d_snap[2];
  DevBuf<uint32_t> d_taken[2], d_from[2];
  PinnedBuf<heroam_particle> h_snap[2];
  PinnedBuf<uint32_t> h_taken[2], h_from[2];
  cudaStream_t copy = nullptr;
  cudaEvent_t filled[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
  ~CaptureBuffers() {
    for (int i = 0; i < 2; ++i) {
      if (filled[i]) cudaEventDestroy(filled[i]);
      if (copied[i]) cudaEventDestroy(copied[i]);
    }
    if (copy) cudaStreamDestroy(copy);
  }
};

static int deliver_chunk(const heroam_ctx *ctx, const uint32_t *counts, const std::vector<uint32_t> &first, uint32_t chunk,
                         const heroam_particle *snap, const uint32_t *taken, const uint32_t *from,
                         std::vector<uint32_t> &next_step, heroam_step_sink sink, void *user) {
  for (uint32_t t = 0; t < ctx->n_traj; ++t) {
    if (!taken[t]) continue;
    if (from[t] != next_step[t])
      return fail(HEROAM_E_CUDA, "capture: trajectory %u resumed at step %u, expected %u", t, from[t], next_step[t]);
    const heroam_particle *rec = snap + (size_t)first[t] * chunk;
    for (uint32_t i = 0; i < taken[t]; ++i)
      if (sink(user, t, from[t] + i, rec + (size_t)i * counts[t], counts[t]))
        return fail(HEROAM_E_ARG, "capture: the sink refused step %u of trajectory %u", from[t] + i, t);
    next_step[t] += taken[t];
  }
  return HEROAM_OK;
}

template <class R, class H>
static int run_traced(heroam_ctx *ctx, const Consts &c, const uint32_t *counts, const std::vector<uint32_t> &first,
                      std::vector<uint32_t> &next_step, heroam_step_sink sink, void *user) {
  DevView v = make_view(ctx, c);
  const uint32_t n = ctx->n_traj;
  const unsigned tgrid = (n + 127) / 128;
  const uint32_t cap = ctx->opt.max_steps == 0 || ctx->opt.max_steps > 0xffffffffull ? 0xffffffffu
                                                                                    : (uint32_t)ctx->opt.max_steps;
  // snapshots of one chunk: n_particles * chunk records, at most 256 MB per buffer
  uint32_t chunk = 4096;
  while (chunk > 1 && (size_t)ctx->n_particles * chunk * sizeof(heroam_particle) > (256ull << 20)) chunk /= 2;
  if (ctx->opt.segment_steps && ctx->opt.segment_steps < chunk) chunk = ctx->opt.segment_steps;
  const size_t snap_records = (size_t)ctx->n_particles * chunk;
  CaptureBuffers cb;
  CU(cudaStreamCreateWithFlags(&cb.copy, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    CU(cb.d_snap[i].alloc(snap_records));
    CU(cb.d_taken[i].alloc(n));
    CU(cb.d_from[i].alloc(n));
    CU(cb.h_snap[i].alloc(snap_records));
    CU(cb.h_taken[i].alloc(n));
    CU(cb.h_from[i].alloc(n));
    CU(cudaEventCreateWithFlags(&cb.filled[i], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&cb.copied[i], cudaEventDisableTiming));
  }
  CU(cudaEventRecord(ctx->ev_start, ctx->main));
  CU(cudaMemsetAsync(ctx->d_counters, 0, sizeof(Counters), ctx->main));
  k_prepare<<<tgrid, 128, 0, ctx->main>>>(v, ctx->d_first, ctx->d_counts);
  CU(cudaGetLastError());
  ctx->info.kernel_launches++;
  const Segments sg = make_capture_segments(chunk);
  int cur = 0;
  bool pending = false;  // buffer cur ^ 1 is on its way to the host
  int rc = HEROAM_OK;
  for (;;) {
    CU(cudaMemsetAsync(ctx->d_bin_count, 0, N_BIN_COUNTERS * sizeof(uint32_t), ctx->main));
    const LiveLists everyone = {nullptr, n, ctx->d_live_list};
    k_resolve_and_bin<R, H><<<tgrid, 128, 0, ctx->main>>>(v, sg, cap, ctx->d_bin_count, ctx->d_bin_list, ctx->d_free_list, everyone);
    CU(cudaGetLastError());
    ctx->info.kernel_launches++;
    CU(cudaMemcpyAsync(ctx->h_bin_count, ctx->d_bin_count, N_BIN_COUNTERS * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->main));
    CU(cudaStreamSynchronize(ctx->main));
    uint64_t live = 0;
    for (int b = 0; b < N_BINS; ++b) live += ctx->h_bin_count[b];
    if (live) {
      ctx->info.passes++;
      CaptureView cv;
      cv.snap = cb.d_snap[cur].ptr;
      cv.taken = cb.d_taken[cur].ptr;
      cv.from_step = cb.d_from[cur].ptr;
      cv.chunk = chunk;
      CU(cudaMemsetAsync(cv.taken, 0, (size_t)n * sizeof(uint32_t), ctx->main));
      for (int b = 1; b < BIN_FREE; ++b) {
        const uint32_t cnt = ctx->h_bin_count[b];
        if (!cnt) continue;
        k_integrate_capture<H><<<(cnt + 63) / 64, 64, 0, ctx->main>>>(v, cv, ctx->d_bin_list + (size_t)b * n, cnt);
        CU(cudaGetLastError());
        ctx->info.kernel_launches++;
      }
      CU(cudaEventRecord(cb.filled[cur], ctx->main));
      CU(cudaStreamWaitEvent(cb.copy, cb.filled[cur], 0));
      CU(cudaMemcpyAsync(cb.h_snap[cur].ptr, cv.snap, snap_records * sizeof(heroam_particle), cudaMemcpyDeviceToHost, cb.copy));
      CU(cudaMemcpyAsync(cb.h_taken[cur].ptr, cv.taken, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, cb.copy));
      CU(cudaMemcpyAsync(cb.h_from[cur].ptr, cv.from_step, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, cb.copy));
      CU(cudaEventRecord(cb.copied[cur], cb.copy));
      ctx->info.d2h_bytes += snap_records * sizeof(heroam_particle) + 2ull * n * sizeof(uint32_t);
    }
    if (pending) {  // hand the previous chunk to the sink while this one is integrated
      const int prev = cur ^ 1;
      CU(cudaEventSynchronize(cb.copied[prev]));
      rc = deliver_chunk(ctx, counts, first, chunk, cb.h_snap[prev].ptr, cb.h_taken[prev].ptr, cb.h_from[prev].ptr,
                         next_step, sink, user);
      pending = false;
      if (rc) break;
    }
    if (!live) break;
    pending = true;
    cur ^= 1;
  }
  CU(cudaStreamSynchronize(cb.copy));
  CU(cudaStreamSynchronize(ctx->main));
  if (rc) return rc;
  k_results<<<tgrid, 128, 0, ctx->main>>>(v, ctx->d_results);
  CU(cudaGetLastError());
  ctx->info.kernel_launches++;
  CU(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->main));
  CU(cudaEventRecord(ctx->ev_stop, ctx->main));
  CU(cudaStreamSynchronize(ctx->main));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_stop));
  ctx->info.device_ms = ms;
  ctx->info.particle_steps = ctx->h_counters->particle_steps;
  ctx->info.pair_steps = ctx->h_counters->pair_steps;
  ctx->info.inrange_steps = ctx->h_counters->inrange_steps;
  ctx->info.trajectories = n;
  ctx->info.particles = ctx->n_particles;
  if (ctx->h_counters->unsupported)
    return fail(HEROAM_E_UNSUPPORTED, "%llu trajectories have more than %d moving particles", ctx->h_counters->unsupported,
                MAX_ACTIVE);
  return HEROAM_OK;
}

extern "C" int heroam_gpu_simulate_traced(heroam_ctx *ctx, unsigned long long he_number, const uint32_t *counts,
                                          heroam_particle *flat, int count, heroam_traj_result *results,
                                          heroam_step_sink sink, void *user) {
  if (!sink) return fail(HEROAM_E_ARG, "heroam_gpu_simulate_traced: null sink");
  int rc = heroam_gpu_upload(ctx, he_number, counts, flat, count);
  if (rc) return rc;
  ctx->uploaded = false;
  if (count == 0) return HEROAM_OK;
  Consts c;
  rc = make_consts(ctx->conf, he_number, &c);
  if (rc) return rc;
  std::vector<uint32_t> first((size_t)count), next_step((size_t)count, 0u);
  size_t total = 0;
  for (int i = 0; i < count; ++i) {
    first[i] = (uint32_t)total;
    total += counts[i];
  }
  rc = ctx->opt.mode == HEROAM_MODE_EXACT ? run_traced<Exact, Exact>(ctx, c, counts, first, next_step, sink, user)
                                          : run_traced<Fast, Fast>(ctx, c, counts, first, next_step, sink, user);
  if (rc) return rc;
  std::vector<heroam_traj_result> res((size_t)count);
  rc = heroam_gpu_download(ctx, flat, res.data());
  if (rc) return rc;
  // A trajectory nobody can move in (every particle frozen at import) still runs one iteration of the
  // reference's loop (Q2): its only snapshot is the state it was imported in, which that
  // iteration leaves unchanged.
  size_t at = 0;
  for (int i = 0; i < count; ++i) {
    if (next_step[i] == 0 && res[i].steps == 1) {
      if (sink(user, (uint32_t)i, 0u, flat + at, counts[i]))
        return fail(HEROAM_E_ARG, "capture: the sink refused step 0 of trajectory %d", i);
      next_step[i] = 1;
    }
    if (next_step[i] != res[i].steps)
      return fail(HEROAM_E_CUDA, "capture: trajectory %d ran %u iterations but %u snapshots were taken", i, res[i].steps,
                  next_step[i]);
    at += counts[i];
  }
  if (results) memcpy(results, res.data(), (size_t)count * sizeof(heroam_traj_result));
  return HEROAM_OK;
}

extern "C" int heroam_gpu_last_run_info(const heroam_ctx *ctx, heroam_run_info *out) {
  if (!ctx || !out) return fail(HEROAM_E_ARG, "null argument");
  *out = ctx->info;
  return HEROAM_OK;
}

extern "C" void heroam_gpu_free_particles(heroam_particles *p) {
  if (p) {
    free(p->particle);
    p->particle = nullptr;
  }
}

extern "C" int heroam_gpu_fp32_peak(heroam_ctx *ctx, double *tflops, double *sm_clock_mhz) {
  if (!ctx || !tflops) return fail(HEROAM_E_ARG, "null argument");
  CU(cudaSetDevice(ctx->device));
  DevBuf<float> out_buf;
  CU(out_buf.alloc(1));
  float *d_out = out_buf.ptr;
  const int iters = 4096, blocks = ctx->sm_count * 8, threads = 256;
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CU(cudaEventRecord(ctx->ev_c0, ctx->main));
    k_fma_peak<<<blocks, threads, 0, ctx->main>>>(d_out, iters, 0.999f, 0.001f);
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->ev_c1, ctx->main));
    CU(cudaStreamSynchronize(ctx->main));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ctx->ev_c0, ctx->ev_c1));
    const double flops = 2.0 * 8.0 * 16.0 * (double)iters * (double)blocks * (double)threads;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  *tflops = best;
  if (sm_clock_mhz) {
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
    *sm_clock_mhz = khz / 1000.0;
  }
  return HEROAM_OK;
}

extern "C" int heroam_gpu_fp32_peak_x2(heroam_ctx *ctx, double *tflops) {
  if (!ctx || !tflops) return fail(HEROAM_E_ARG, "null argument");
  CU(cudaSetDevice(ctx->device));
  DevBuf<float> out_buf, coef_buf;
  CU(out_buf.alloc(1));
  CU(coef_buf.alloc(8));
  const float coef[8] = {0.999f, 0.998f, 0.997f, 0.996f, 0.001f, 0.002f, 0.003f, 0.004f};
  CU(cudaMemcpy(coef_buf.ptr, coef, sizeof coef, cudaMemcpyHostToDevice));
  const int iters = 4096, blocks = ctx->sm_count * 8, threads = 256;
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CU(cudaEventRecord(ctx->ev_c0, ctx->main));
    k_fma_peak_x2<<<blocks, threads, 0, ctx->main>>>(out_buf.ptr, coef_buf.ptr, iters);
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->ev_c1, ctx->main));
    CU(cudaStreamSynchronize(ctx->main));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ctx->ev_c0, ctx->ev_c1));
    const double flops = 2.0 * 16.0 * 16.0 * (double)iters * (double)blocks * (double)threads;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  *tflops = best;
  return HEROAM_OK;
}

extern "C" int heroam_gpu_fp32_peak_rrr(heroam_ctx *ctx, double *tflops) {
  if (!ctx || !tflops) return fail(HEROAM_E_ARG, "null argument");
  CU(cudaSetDevice(ctx->device));
  DevBuf<float> out_buf, coef_buf;
  CU(out_buf.alloc(1));
  CU(coef_buf.alloc(8));
  const float coef[8] = {0.999f, 0.998f, 0.997f, 0.996f, 0.001f, 0.002f, 0.003f, 0.004f};
  CU(cudaMemcpy(coef_buf.ptr, coef, sizeof coef, cudaMemcpyHostToDevice));
  const int iters = 4096, blocks = ctx->sm_count * 8, threads = 256;
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CU(cudaEventRecord(ctx->ev_c0, ctx->main));
    k_fma_peak_rrr<<<blocks, threads, 0, ctx->main>>>(out_buf.ptr, coef_buf.ptr, iters);
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->ev_c1, ctx->main));
    CU(cudaStreamSynchronize(ctx->main));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ctx->ev_c0, ctx->ev_c1));
    const double flops = 2.0 * 8.0 * 16.0 * (double)iters * (double)blocks * (double)threads;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  *tflops = best;
  return HEROAM_OK;
}

// ---------------------------------------------------------------------------------
// sampler and fused statistics: see heroam_sampler.cuh
// ---------------------------------------------------------------------------------
#include "heroam_sampler.cuh"
