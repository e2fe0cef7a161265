// api.cu — the extern "C" surface declared in include/hydrocore.h, the context and its arena.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"
#include "tma.cuh"

static thread_local char g_err[512] = "";

void hc_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char *hc_last_error(void) { return g_err; }
extern "C" int hc_version(void) { return 100; }

extern "C" int hc_debug_read(hc_ctx *ctx, unsigned long long *out16) {
  if (!ctx || !out16) return HC_ERR_ARG;
  hc_guard hc_guard__(ctx);
  HC_CUDA(cudaSetDevice(ctx->device));
  HC_CUDA(cudaDeviceSynchronize());
  HC_CUDA(cudaMemcpy(out16, ctx->d_dbg, sizeof(unsigned long long) * 16, cudaMemcpyDeviceToHost));
  HC_CUDA(cudaMemset(ctx->d_dbg, 0, sizeof(unsigned long long) * 16));
  return HC_OK;
}


// ---- TMA tensor maps ----------------------------------------------------------------------------
// cuTensorMapEncodeTiled is a driver-API call; it is fetched through the runtime (no link against libcuda.so,
// which does not exist on a build box without a GPU driver).
typedef CUresult (*hc_pfn_encode_tiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                        const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int hc_tmap_2d(hc_tmap *out, const void *base, int elem_bytes, int is_float, uint64_t width, uint64_t height,
               uint64_t pitch_bytes, uint32_t box_w, uint32_t box_h) {
  static hc_pfn_encode_tiled fn = nullptr;
  static int tried = 0;
  out->ok = 0;
  if (getenv("HC_NO_TMA")) return 1;
  if (!tried) {
    tried = 1;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (hc_pfn_encode_tiled)p;
    else
      cudaGetLastError();
  }
  if (!fn || !hc_tmap_ok(base, pitch_bytes) || box_w > 256 || box_h > 256 || ((uint64_t)box_w * elem_bytes) % 16) return 1;
  const cuuint64_t dims[2] = {width, height};
  const cuuint64_t strides[1] = {pitch_bytes};
  const cuuint32_t box[2] = {box_w, box_h};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                 : (is_float ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT32);
  CUresult r = fn(&out->map, dt, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  is_float ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA : CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return 1;
  out->ok = 1;
  return 0;
}

// ---- arena ------------------------------------------------------------------------------------
int arena_reset(hc_ctx *ctx) {
  if (ctx->n_blocks > 1) {
    // coalesce: next time the same workload fits one block and allocates nothing
    size_t total = 0;
    HC_CUDA(cudaDeviceSynchronize());
    for (int i = 0; i < ctx->n_blocks; i++) {
      total += ctx->blocks[i].cap;
      cudaFree(ctx->blocks[i].p);
    }
    ctx->n_blocks = 0;
    char *p = nullptr;
    if (cudaMalloc(&p, total) == cudaSuccess) {
      ctx->blocks[0].p = p;
      ctx->blocks[0].cap = total;
      ctx->blocks[0].off = 0;
      ctx->n_blocks = 1;
    } else {
      cudaGetLastError();  // fall back to on-demand blocks
    }
  }
  for (int i = 0; i < ctx->n_blocks; i++) ctx->blocks[i].off = 0;
  return HC_OK;
}

void *arena_take(hc_ctx *ctx, size_t bytes) {
  size_t b = hc_align(bytes ? bytes : 1);
  for (int i = 0; i < ctx->n_blocks; i++) {
    hc_block *k = &ctx->blocks[i];
    if (k->cap - k->off >= b) {
      void *p = k->p + k->off;
      k->off += b;
      return p;
    }
  }
  if (ctx->n_blocks == HC_MAX_BLOCKS) { hc_set_error("arena: too many blocks"); return nullptr; }
  char *p = nullptr;
  cudaError_t e = cudaMalloc(&p, b);
  if (e != cudaSuccess) {
    cudaGetLastError();
    hc_set_error("arena: cudaMalloc(%zu) -> %s", b, cudaGetErrorString(e));
    return nullptr;
  }
  hc_block *k = &ctx->blocks[ctx->n_blocks++];
  k->p = p; k->cap = b; k->off = b;
  return p;
}

// ---- per-kernel event timing --------------------------------------------------------------------
void prof_begin(hc_ctx *ctx, cudaStream_t st, const char *name) {
  if (ctx->prof_n >= ctx->prof_cap) return;  // pool exhausted: stop recording, keep running
  int i = ctx->prof_n;
  ctx->prof_name[i] = name;
  ctx->prof_stream = st;
  cudaEventRecord(ctx->prof_ev[2 * i], st);
  ctx->prof_open = 1;
}
void prof_end(hc_ctx *ctx) {
  cudaEventRecord(ctx->prof_ev[2 * ctx->prof_n + 1], ctx->prof_stream);
  ctx->prof_n++;
  ctx->prof_open = 0;
}

extern "C" int hc_profile_enable(hc_ctx *ctx, int on, int max_records) {
  if (!ctx) return HC_ERR_ARG;
  hc_guard hc_guard__(ctx);
  HC_CUDA(cudaSetDevice(ctx->device));
  if (on && !ctx->prof_ev) {
    int cap = max_records > 0 ? max_records : 8192;
    ctx->prof_ev = (cudaEvent_t *)calloc((size_t)cap * 2, sizeof(cudaEvent_t));
    ctx->prof_name = (const char **)calloc((size_t)cap, sizeof(char *));
    if (!ctx->prof_ev || !ctx->prof_name) return HC_ERR_NOMEM;
    for (int i = 0; i < cap * 2; i++) HC_CUDA(cudaEventCreate(&ctx->prof_ev[i]));
    ctx->prof_cap = cap;
  }
  ctx->prof_on = on ? 1 : 0;
  ctx->prof_n = 0;
  ctx->prof_open = 0;
  return HC_OK;
}

// Writes "name count total_ms max_ms\n" per kernel name into buf (NUL terminated); resets the log.
extern "C" int hc_profile_read(hc_ctx *ctx, char *buf, int64_t buflen) {
  if (!ctx || !buf || buflen < 2) return HC_ERR_ARG;
  hc_guard hc_guard__(ctx);
  HC_CUDA(cudaSetDevice(ctx->device));
  HC_CUDA(cudaDeviceSynchronize());
  struct agg { const char *name; int64_t cnt; double tot, mx; };
  agg a[128];
  int na = 0;
  for (int i = 0; i < ctx->prof_n; i++) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->prof_ev[2 * i], ctx->prof_ev[2 * i + 1]) != cudaSuccess) { cudaGetLastError(); continue; }
    int j = 0;
    for (; j < na; j++) if (a[j].name == ctx->prof_name[i] || !strcmp(a[j].name, ctx->prof_name[i])) break;
    if (j == na) { if (na == 128) continue; a[na].name = ctx->prof_name[i]; a[na].cnt = 0; a[na].tot = 0; a[na].mx = 0; na++; }
    a[j].cnt++; a[j].tot += ms; if (ms > a[j].mx) a[j].mx = ms;
  }
  int64_t off = 0;
  buf[0] = 0;
  for (int j = 0; j < na; j++) {
    int w = snprintf(buf + off, (size_t)(buflen - off), "%s %lld %.6f %.6f\n", a[j].name, (long long)a[j].cnt, a[j].tot, a[j].mx);
    if (w < 0 || off + w >= buflen) break;
    off += w;
  }
  ctx->prof_n = 0;
  return HC_OK;
}

// ---- context ----------------------------------------------------------------------------------
extern "C" int hc_create(hc_ctx **out, int device) {
  if (!out) { hc_set_error("hc_create: null out"); return HC_ERR_ARG; }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    hc_set_error("hc_create: no CUDA device (%s) - libhydrocore has no CPU path", cudaGetErrorString(e));
    cudaGetLastError();
    return HC_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) { hc_set_error("hc_create: device %d out of range", device); return HC_ERR_ARG; }
  HC_CUDA(cudaSetDevice(device));
  hc_ctx *ctx = (hc_ctx *)calloc(1, sizeof(hc_ctx));
  if (!ctx) return HC_ERR_NOMEM;
  ctx->device = device;
  {
    pthread_mutexattr_t at;
    pthread_mutexattr_init(&at);
    pthread_mutexattr_settype(&at, PTHREAD_MUTEX_RECURSIVE);
    pthread_mutex_init(&ctx->mu, &at);
    pthread_mutexattr_destroy(&at);
  }
  cudaDeviceProp prop;
  HC_CUDA(cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  HC_CUDA(cudaMallocHost(&ctx->h_mail, 256));
  memset(ctx->h_mail, 0, 256);
  HC_CUDA(cudaMalloc(&ctx->d_dbg, 16 * sizeof(unsigned long long)));
  HC_CUDA(cudaMemset(ctx->d_dbg, 0, 16 * sizeof(unsigned long long)));
  HC_CUDA(cudaStreamCreateWithFlags(&ctx->s_compute, cudaStreamNonBlocking));
  HC_CUDA(cudaStreamCreateWithFlags(&ctx->s_copy, cudaStreamNonBlocking));
  HC_CUDA(cudaEventCreateWithFlags(&ctx->ev_fill, cudaEventDisableTiming));
  HC_CUDA(cudaEventCreateWithFlags(&ctx->ev_dir, cudaEventDisableTiming));
  HC_CUDA(cudaEventCreateWithFlags(&ctx->ev_acc, cudaEventDisableTiming));
  HC_CUDA(cudaEventCreateWithFlags(&ctx->ev_done[0], cudaEventDisableTiming));
  HC_CUDA(cudaEventCreateWithFlags(&ctx->ev_done[1], cudaEventDisableTiming));
  // the flat pass's slow tier runs latency-bound rounds beside full-grid work of the chain: its streams get the highest
  // priority so that their (few) CTAs are scheduled ahead of the queued CTAs of the big kernel
  int prio_least = 0, prio_greatest = 0;
  HC_CUDA(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
  HC_CUDA(cudaStreamCreateWithPriority(&ctx->s_aux, cudaStreamNonBlocking, prio_greatest));
  HC_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  HC_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
  HC_CUDA(cudaStreamCreateWithPriority(&ctx->s_aux2, cudaStreamNonBlocking, prio_greatest));
  HC_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork2, cudaEventDisableTiming));
  HC_CUDA(cudaEventCreateWithFlags(&ctx->ev_join2, cudaEventDisableTiming));
  *out = ctx;
  return HC_OK;
}

extern "C" int hc_destroy(hc_ctx *ctx) {
  if (!ctx) return HC_OK;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (int i = 0; i < ctx->n_blocks; i++) cudaFree(ctx->blocks[i].p);
  for (int i = 0; i < 10; i++) if (ctx->io[i]) cudaFree(ctx->io[i]);
  if (ctx->h_mail) cudaFreeHost(ctx->h_mail);
  if (ctx->s_compute) cudaStreamDestroy(ctx->s_compute);
  if (ctx->s_copy) cudaStreamDestroy(ctx->s_copy);
  if (ctx->ev_fill) cudaEventDestroy(ctx->ev_fill);
  if (ctx->ev_dir) cudaEventDestroy(ctx->ev_dir);
  if (ctx->ev_acc) cudaEventDestroy(ctx->ev_acc);
  for (int i = 0; i < 2; i++) if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
  if (ctx->s_aux) cudaStreamDestroy(ctx->s_aux);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->s_aux2) cudaStreamDestroy(ctx->s_aux2);
  if (ctx->ev_fork2) cudaEventDestroy(ctx->ev_fork2);
  if (ctx->ev_join2) cudaEventDestroy(ctx->ev_join2);
  pthread_mutex_destroy(&ctx->mu);
  free(ctx);
  return HC_OK;
}

extern "C" int64_t hc_workspace_bytes(const hc_ctx *ctx) {
  int64_t t = 0;
  for (int i = 0; i < ctx->n_blocks; i++) t += (int64_t)ctx->blocks[i].cap;
  return t;
}
extern "C" int64_t hc_launch_count(const hc_ctx *ctx) { return ctx->launches; }
extern "C" int hc_fill_stats(const hc_ctx *ctx, int64_t *n_basins, int32_t *n_rounds) {
  if (n_basins) *n_basins = ctx->fill_basins;
  if (n_rounds) *n_rounds = ctx->fill_rounds == -1 ? (int32_t)ctx->h_mail[32] : ctx->fill_rounds;   // -1: fetched asynchronously, valid once the caller has synchronised
  return HC_OK;
}

#define CHECK_CTX(ctx) HC_ENTER(ctx)

static int check_shape(int64_t ny, int64_t nx) {
  if (ny < 0 || nx < 0) { hc_set_error("negative shape"); return HC_ERR_ARG; }
  if (ny > 0 && nx > 0 && (double)ny * (double)nx >= 2147483632.0) {
    hc_set_error("raster of %lld x %lld cells exceeds the 31-bit cell index", (long long)ny, (long long)nx);
    return HC_ERR_TOO_LARGE;
  }
  return HC_OK;
}

// ---- device entry points ------------------------------------------------------------------------
extern "C" int hc_fill_dev(hc_ctx *ctx, const float *z, float *out, int64_t ny, int64_t nx, float nodata,
                           int seed_mode, void *stream) {
  CHECK_CTX(ctx);
  HC_TRY(check_shape(ny, nx));
  if (ny && nx && (!z || !out)) { hc_set_error("hc_fill: null pointer"); return HC_ERR_ARG; }
  if (seed_mode != HC_SEED_WBT && seed_mode != HC_SEED_TAUDEM) { hc_set_error("hc_fill: bad seed_mode"); return HC_ERR_ARG; }
  return fill_run(ctx, z, out, ny, nx, nodata, seed_mode, (cudaStream_t)stream);
}

extern "C" int hc_d8_dev(hc_ctx *ctx, const float *z, uint8_t *dir, float *slope, int64_t ny, int64_t nx,
                         float nodata, double dx, double dy, int table, void *stream) {
  CHECK_CTX(ctx);
  HC_TRY(check_shape(ny, nx));
  if (ny && nx && (!z || !dir)) { hc_set_error("hc_d8: null pointer"); return HC_ERR_ARG; }
  return d8_run(ctx, z, dir, slope, ny, nx, nodata, dx, dy, table, (cudaStream_t)stream);
}

extern "C" int hc_resolve_flats_dev(hc_ctx *ctx, const float *z, uint8_t *dir, int64_t ny, int64_t nx,
                                    float nodata, int table, void *stream) {
  CHECK_CTX(ctx);
  HC_TRY(check_shape(ny, nx));
  if (ny && nx && (!z || !dir)) { hc_set_error("hc_resolve_flats: null pointer"); return HC_ERR_ARG; }
  return flats_run(ctx, z, dir, ny, nx, nodata, table, (cudaStream_t)stream);
}

extern "C" int hc_accum_dev(hc_ctx *ctx, const uint8_t *dir, uint32_t *acc, int64_t ny, int64_t nx, int table,
                            void *stream) {
  CHECK_CTX(ctx);
  HC_TRY(check_shape(ny, nx));
  if (ny && nx && (!dir || !acc)) { hc_set_error("hc_accum: null pointer"); return HC_ERR_ARG; }
  return accum_run(ctx, dir, acc, ny, nx, table, (cudaStream_t)stream);
}

extern "C" int hc_fill_d8_accum_dev(hc_ctx *ctx, const float *z, float *filled, uint8_t *dir, float *slope,
                                    uint32_t *acc, int64_t ny, int64_t nx, float nodata, double dx, double dy,
                                    int seed_mode, int table, int resolve_flats, void *stream) {
  CHECK_CTX(ctx);
  HC_TRY(check_shape(ny, nx));
  if (ny && nx && (!z || !filled || !dir || !acc)) { hc_set_error("hc_fill_d8_accum: null pointer"); return HC_ERR_ARG; }
  if (seed_mode != HC_SEED_WBT && seed_mode != HC_SEED_TAUDEM) { hc_set_error("bad seed_mode"); return HC_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  HC_TRY(fill_d8_flats_accum_run(ctx, z, filled, dir, slope, acc, ny, nx, nodata, dx, dy, seed_mode, table, resolve_flats, st));
  return HC_OK;
}


// padding columns [nx, ld) of a pitched float raster := nodata (see hc_fill_d8_accum_ld_dev)
__global__ void k_pad_nodata(float *z, int ny, int nx, int ld, float nodata) {
  const int np = ld - nx;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)ny * np) return;
  const int r = (int)(i / np), c = nx + (int)(i - (long long)r * np);
  z[(size_t)r * ld + c] = nodata;
}

// Pitched ("leading dimension") form of the chain: every raster has ld >= nx elements per row.  With ld a multiple of
// 16 (and 16-byte aligned bases) every row is vector- and TMA-addressable whatever nx is (a 10801-wide dense raster
// is neither).  The padding columns are turned into nodata and the chain runs on the ny x ld raster: under both tools'
// conventions a nodata column right of the raster acts exactly like the raster edge (column nx - 1 is a seed / is
// contaminated either way; the padding is exterior nodata; it holds no direction and no count), so the first nx
// columns of every output equal the dense call's result.  Output padding columns hold unspecified values.
extern "C" int hc_fill_d8_accum_ld_dev(hc_ctx *ctx, float *z, float *filled, uint8_t *dir, float *slope,
                                       uint32_t *acc, int64_t ny, int64_t nx, int64_t ld, float nodata, double dx,
                                       double dy, int seed_mode, int table, int resolve_flats, void *stream) {
  CHECK_CTX(ctx);
  if (ld < nx) { hc_set_error("hc_fill_d8_accum_ld: ld < nx"); return HC_ERR_ARG; }
  HC_TRY(check_shape(ny, ld));
  if (ny && nx && (!z || !filled || !dir || !acc)) { hc_set_error("hc_fill_d8_accum_ld: null pointer"); return HC_ERR_ARG; }
  if (seed_mode != HC_SEED_WBT && seed_mode != HC_SEED_TAUDEM) { hc_set_error("bad seed_mode"); return HC_ERR_ARG; }
  if (ny <= 0 || nx <= 0) return HC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (ld > nx) {
    const long long np = (long long)ny * (ld - nx);
    k_pad_nodata<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(z, (int)ny, (int)nx, (int)ld, nodata);
    HC_LAUNCH_CHECK(ctx);
  }
  ctx->pad_from = ld > nx ? (int)nx : 0;   // (the WBT seed rule's exterior-nodata search need not label the padding)
  int rc = fill_d8_flats_accum_run(ctx, z, filled, dir, slope, acc, ny, ld, nodata, dx, dy, seed_mode, table, resolve_flats, st);
  ctx->pad_from = 0;
  HC_TRY(rc);
  return HC_OK;
}

extern "C" int hc_label_dev(hc_ctx *ctx, const uint8_t *mask, int32_t *lab, int64_t ny, int64_t nx, int conn,
                            int32_t *nlab, void *stream) {
  CHECK_CTX(ctx);
  HC_TRY(check_shape(ny, nx));
  if (ny && nx && (!mask || !lab)) { hc_set_error("hc_label: null pointer"); return HC_ERR_ARG; }
  return label_run(ctx, mask, lab, ny, nx, conn, nlab, (cudaStream_t)stream);
}

// ---- host entry points --------------------------------------------------------------------------
static int io_get(hc_ctx *ctx, int slot, size_t bytes, void **out) {
  if (ctx->io_cap[slot] < bytes) {
    if (ctx->io[slot]) { HC_CUDA(cudaDeviceSynchronize()); cudaFree(ctx->io[slot]); ctx->io[slot] = nullptr; ctx->io_cap[slot] = 0; }
    cudaError_t e = cudaMalloc(&ctx->io[slot], bytes);
    if (e != cudaSuccess) { cudaGetLastError(); hc_set_error("cudaMalloc(%zu) -> %s", bytes, cudaGetErrorString(e)); return HC_ERR_NOMEM; }
    ctx->io_cap[slot] = bytes;
  }
  *out = ctx->io[slot];
  return HC_OK;
}
struct dev_buf {
  void *p = nullptr;
  hc_ctx *ctx; int slot;
  dev_buf(hc_ctx *c, int s) : ctx(c), slot(s) {}
  int alloc(size_t b) { return io_get(ctx, slot, b ? b : 1, &p); }
};

extern "C" int hc_fill_host(hc_ctx *ctx, const float *z, float *out, int64_t ny, int64_t nx, float nodata,
                            int seed_mode) {
  CHECK_CTX(ctx);
  HC_TRY(check_shape(ny, nx));
  size_t n = (size_t)ny * nx;
  if (!n) return HC_OK;
  if (!z || !out) { hc_set_error("hc_fill_host: null pointer"); return HC_ERR_ARG; }
  // shares the staging buffers of hc_fill_d8_accum_host_begin(slot 0): wait until that call's copies have landed
  HC_CUDA(cudaEventSynchronize(ctx->ev_done[0]));
  HC_CUDA(cudaEventSynchronize(ctx->ev_done[1]));
  HC_CUDA(cudaStreamSynchronize(ctx->s_compute));
  dev_buf dz(ctx, 0), dout(ctx, 1);
  HC_TRY(dz.alloc(n * 4));
  HC_TRY(dout.alloc(n * 4));
  HC_CUDA(cudaMemcpyAsync(dz.p, z, n * 4, cudaMemcpyHostToDevice, 0));
  HC_TRY(hc_fill_dev(ctx, (const float *)dz.p, (float *)dout.p, ny, nx, nodata, seed_mode, nullptr));
  HC_CUDA(cudaMemcpyAsync(out, dout.p, n * 4, cudaMemcpyDeviceToHost, 0));
  HC_CUDA(cudaStreamSynchronize(0));
  return HC_OK;
}

// Begin / end form of the host chain: `begin` copies the DEM in, runs the chain and QUEUES the device->host copies of
// the results; `end` waits until the slot's last copy has landed.  Two slots (0 / 1) own separate device staging
// buffers, so begin(slot ^ 1) may be called before end(slot): the PCIe link carries one unit's results (9 B/cell)
// while the next unit is uploaded (other direction) and computed.  Host buffers must be pinned for the overlap to
// happen and must stay untouched until `end`.
extern "C" int hc_fill_d8_accum_host_begin(hc_ctx *ctx, int slot, const float *z, float *filled, uint8_t *dir,
                                           float *slope, uint32_t *acc, int64_t ny, int64_t nx, float nodata,
                                           double dx, double dy, int seed_mode, int table, int resolve_flats) {
  CHECK_CTX(ctx);
  HC_TRY(check_shape(ny, nx));
  if (slot != 0 && slot != 1) { hc_set_error("hc_fill_d8_accum_host_begin: slot must be 0 or 1"); return HC_ERR_ARG; }
  size_t n = (size_t)ny * nx;
  if (!n) return HC_OK;
  if (!z || !filled || !dir || !acc) { hc_set_error("hc_fill_d8_accum_host: null pointer"); return HC_ERR_ARG; }
  if (seed_mode != HC_SEED_WBT && seed_mode != HC_SEED_TAUDEM) { hc_set_error("bad seed_mode"); return HC_ERR_ARG; }
  if (table != HC_TABLE_WBT && table != HC_TABLE_TAUDEM) { hc_set_error("bad table"); return HC_ERR_ARG; }
  // the slot's staging buffers may still be the source of the copies of the call that used it last
  HC_CUDA(cudaEventSynchronize(ctx->ev_done[slot]));
  // device staging rasters are PITCHED: ld = nx rounded up to 64 elements, so that every row is 16-byte aligned for
  // all element sizes (TMA and vector loads need that; a dense 10801-wide raster offers neither).  The padding columns
  // of z hold nodata (hc_fill_d8_accum_ld_dev writes them); the 2-D copies move only the nx real columns.
  const int64_t ld = (nx + 63) & ~(int64_t)63;
  HC_TRY(check_shape(ny, ld));
  const size_t np = (size_t)ny * ld;
  const int b = 5 * slot;
  dev_buf dz(ctx, b + 0), df(ctx, b + 1), dd(ctx, b + 2), ds(ctx, b + 3), da(ctx, b + 4);
  HC_TRY(dz.alloc(np * 4));
  HC_TRY(df.alloc(np * 4));
  HC_TRY(dd.alloc(np));
  if (slope) HC_TRY(ds.alloc(np * 4));
  HC_TRY(da.alloc(np * 4));
  // compute on one stream, device->host copies on another: the filled raster travels while D8 / flats run, the
  // direction raster while the accumulation runs (the PCIe link, not the kernels, bounds this entry point)
  cudaStream_t sc = ctx->s_compute, sx = ctx->s_copy;
  HC_CUDA(cudaMemcpy2DAsync(dz.p, (size_t)ld * 4, z, (size_t)nx * 4, (size_t)nx * 4, (size_t)ny, cudaMemcpyHostToDevice, sc));
  if (ld > nx) {
    const long long npad = (long long)ny * (ld - nx);
    k_pad_nodata<<<(unsigned)((npad + 255) / 256), 256, 0, sc>>>((float *)dz.p, (int)ny, (int)nx, (int)ld, nodata);
    HC_LAUNCH_CHECK(ctx);
  }
  ctx->pad_from = ld > nx ? (int)nx : 0;
  int rc = fill_d8_flats_accum_run(ctx, (const float *)dz.p, (float *)df.p, (uint8_t *)dd.p, (float *)ds.p, (uint32_t *)da.p, ny, ld,
                                   nodata, dx, dy, seed_mode, table, resolve_flats, sc);
  ctx->pad_from = 0;
  HC_TRY(rc);
  HC_CUDA(cudaEventRecord(ctx->ev_dir, sc));
  HC_CUDA(cudaStreamWaitEvent(sx, ctx->ev_dir, 0));
  HC_CUDA(cudaMemcpy2DAsync(filled, (size_t)nx * 4, df.p, (size_t)ld * 4, (size_t)nx * 4, (size_t)ny, cudaMemcpyDeviceToHost, sx));
  HC_CUDA(cudaMemcpy2DAsync(dir, (size_t)nx, dd.p, (size_t)ld, (size_t)nx, (size_t)ny, cudaMemcpyDeviceToHost, sx));
  if (slope) HC_CUDA(cudaMemcpy2DAsync(slope, (size_t)nx * 4, ds.p, (size_t)ld * 4, (size_t)nx * 4, (size_t)ny, cudaMemcpyDeviceToHost, sx));
  HC_CUDA(cudaEventRecord(ctx->ev_acc, sc));
  HC_CUDA(cudaStreamWaitEvent(sx, ctx->ev_acc, 0));
  HC_CUDA(cudaMemcpy2DAsync(acc, (size_t)nx * 4, da.p, (size_t)ld * 4, (size_t)nx * 4, (size_t)ny, cudaMemcpyDeviceToHost, sx));
  HC_CUDA(cudaEventRecord(ctx->ev_done[slot], sx));
  return HC_OK;
}

extern "C" int hc_fill_d8_accum_host_end(hc_ctx *ctx, int slot) {
  CHECK_CTX(ctx);
  if (slot != 0 && slot != 1) { hc_set_error("hc_fill_d8_accum_host_end: slot must be 0 or 1"); return HC_ERR_ARG; }
  HC_CUDA(cudaEventSynchronize(ctx->ev_done[slot]));
  return HC_OK;
}

extern "C" int hc_fill_d8_accum_host(hc_ctx *ctx, const float *z, float *filled, uint8_t *dir, float *slope,
                                     uint32_t *acc, int64_t ny, int64_t nx, float nodata, double dx, double dy,
                                     int seed_mode, int table, int resolve_flats) {
  HC_TRY(hc_fill_d8_accum_host_begin(ctx, 0, z, filled, dir, slope, acc, ny, nx, nodata, dx, dy, seed_mode, table,
                                     resolve_flats));
  return hc_fill_d8_accum_host_end(ctx, 0);
}

extern "C" int hc_label_host(hc_ctx *ctx, const uint8_t *mask, int32_t *lab, int64_t ny, int64_t nx, int conn,
                             int32_t *nlab) {
  CHECK_CTX(ctx);
  HC_TRY(check_shape(ny, nx));
  size_t n = (size_t)ny * nx;
  if (nlab) *nlab = 0;
  if (!n) return HC_OK;
  if (!mask || !lab) { hc_set_error("hc_label_host: null pointer"); return HC_ERR_ARG; }
  HC_CUDA(cudaEventSynchronize(ctx->ev_done[0]));
  HC_CUDA(cudaEventSynchronize(ctx->ev_done[1]));
  HC_CUDA(cudaStreamSynchronize(ctx->s_compute));
  dev_buf dm(ctx, 2), dl(ctx, 4);
  HC_TRY(dm.alloc(n));
  HC_TRY(dl.alloc(n * 4));
  HC_CUDA(cudaMemcpyAsync(dm.p, mask, n, cudaMemcpyHostToDevice, 0));
  HC_TRY(hc_label_dev(ctx, (const uint8_t *)dm.p, (int32_t *)dl.p, ny, nx, conn, nlab, nullptr));
  HC_CUDA(cudaMemcpyAsync(lab, dl.p, n * 4, cudaMemcpyDeviceToHost, 0));
  HC_CUDA(cudaStreamSynchronize(0));
  return HC_OK;
}
