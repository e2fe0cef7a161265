// common.cuh — context, workspace arena, error plumbing shared by every pass of libhydrocore.
#pragma once
#include <cuda_runtime.h>
#include <pthread.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/hydrocore.h"

#define HC_TILE 64             // square tile edge used by the tiled passes (fill, flats)
#define HC_TAG 0x80000000u     // top bit marks a resolved basin id in the pointer/label array
#define HC_LAB_NODATA 0x7FFFFFFFu
#define HC_NONE 0xFFFFFFFFu
#define HC_NFCAP 1024          // NOFLOW cells listed per 64 x 64 tile (flat pass); a tile holding more is "dense"
#define HC_NF_HI 0x8000u       // list entry = local index (12 bits) | HC_NF_HI when the cell touches higher ground

typedef unsigned int u32;
typedef unsigned long long u64;

void hc_set_error(const char *fmt, ...);

#define HC_CUDA(call)                                                                     \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      hc_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return HC_ERR_CUDA;                                                                 \
    }                                                                                     \
  } while (0)

#define HC_TRY(call)          \
  do {                        \
    int r__ = (call);         \
    if (r__ != HC_OK) return r__; \
  } while (0)

// optional per-kernel timing: HC_PROF(ctx, st, "name") before a launch, the matching
// HC_LAUNCH_CHECK records the closing event (see prof_begin / prof_end in api.cu)
#define HC_PROF(ctx, st, name)                              \
  do {                                                      \
    if ((ctx)->prof_on) prof_begin((ctx), (st), (name));    \
  } while (0)

#define HC_LAUNCH_CHECK(ctx)                                                            \
  do {                                                                                  \
    (ctx)->launches++;                                                                  \
    if ((ctx)->prof_open) prof_end(ctx);                                                \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) {                                                           \
      hc_set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__));  \
      return HC_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

// Row-band sharding: a band's rasters carry one halo row above (row -1) and below (row nb) when the
// band has a neighbour there.  Kernels may READ rows [rlo, rhi); they own rows [0, nb).
struct hc_rows { int rlo, rhi, top, bot; };
static inline hc_rows hc_make_rows(int64_t nb, int flags) {
  hc_rows rw = {(flags & 1) ? -1 : 0, (int)nb + ((flags & 2) ? 1 : 0), flags & 1, (flags >> 1) & 1};
  return rw;
}

// state a band keeps between the phases of one pass (pointers into the arena)
struct hc_band_state {
  int flags; int64_t nb, nx, band_index;
  // fill
  int fill_ready; const float *z; u32 *P; float *level0; u32 *troot; u32 K1, nopen; u32 *cnt;
  // flats
  int flats_ready, table; float nodata; const float *fz; const uint8_t *dirA; uint8_t *dirB;
  uint8_t *pending, *active, *visited, *tmask; u32 *dlo, *dhi, *flag, *tcount; unsigned short *tlist; int n_pending;
  u32 *qA, *qB, *queued;   // dirty-tile queues of the slow tier (current / next) and the per-tile "already queued" marks
  // accumulation
  int acc_ready; uint32_t *acc_out; const uint8_t *dir; unsigned long long *pstate; unsigned short *pexit; u32 *pdown, *pinflow, *bout, *acc_err;
};

// banded labelling: state between the phases (pointers into the arena)
struct band_label_state { u32 *L, *G, *R, *flag; int64_t nb, nx; int conn, ready; u32 base, owned; };

#define HC_MAX_BLOCKS 64
struct hc_block { char *p; size_t cap, off; };

struct hc_ctx {
  int device;
  // one call at a time per context: every entry point holds this (recursive) lock while it touches the arena, the
  // mailbox, the staging buffers or the band state; concurrent callers of ONE context serialise instead of
  // overwriting each other's workspace.  Use one context per thread / stream for concurrency.
  pthread_mutex_t mu;
  int sm_count;
  // chunked bump arena: blocks are only ever ADDED inside a pass (pointers handed out stay
  // valid); arena_reset() at the start of a public call coalesces them into one block, so a
  // steady-state call does no cudaMalloc at all.
  hc_block blocks[HC_MAX_BLOCKS];
  int n_blocks;
  // pinned mailbox for the few scalar read-backs
  u32 *h_mail;
  int64_t launches;
  // stats of the last fill
  int64_t fill_basins;
  int32_t fill_rounds;
  float deg_thr[90]; int deg_thr_n;   // slope -> truncated-degree thresholds (hc_slope_thresholds_set), 0 = atanf form
  uint32_t *chain_acc;   // accumulation raster of the running hc_fill_d8_accum_* chain (fill_d8_flats_accum_run)
  int pad_from;   // > 0: columns >= pad_from are the nodata padding of a pitched raster (hc_fill_d8_accum_ld_dev)
  // band diagnostics: terminal-graph edges emitted, MSF rounds, global seam-graph rounds
  int64_t band_term_edges;
  int32_t band_msf_rounds, band_global_rounds;
  hc_band_state band;
  band_label_state band_label;
  // grow-only device staging buffers of the *_host entry points (z, filled, dir, slope, acc)
  void *io[10];        // [5 * slot + k]: two sets, so that one call's results can travel while the next call computes
  size_t io_cap[10];
  // streams / events of the *_host chain: compute and device->host copies overlap
  cudaStream_t s_compute, s_copy;
  cudaEvent_t ev_fill, ev_dir;
  cudaEvent_t ev_acc, ev_done[2];   // begin/end form of the host chain: accumulation ready; slot's last copy landed
  // auxiliary stream of the flat pass (light and window relaxation kernels run side by side)
  cudaStream_t s_aux;
  cudaEvent_t ev_fork, ev_join;
  // second auxiliary stream: accumulation of the tiles without pending flats beside the flats' slow tier
  cudaStream_t s_aux2;
  cudaEvent_t ev_fork2, ev_join2;
  unsigned long long *d_dbg;  // 16 device work counters
  // per-kernel event timing (off by default)
  int prof_on, prof_open;
  int prof_n, prof_cap;
  cudaEvent_t *prof_ev;      // 2 per record
  const char **prof_name;
  cudaStream_t prof_stream;
};

struct hc_guard {
  hc_ctx *c;
  explicit hc_guard(hc_ctx *ctx) : c(ctx) { pthread_mutex_lock(&c->mu); }
  ~hc_guard() { pthread_mutex_unlock(&c->mu); }
  hc_guard(const hc_guard &) = delete;
  hc_guard &operator=(const hc_guard &) = delete;
};
// first statement of every extern "C" entry point that takes a context
#define HC_ENTER(ctx)                                                \
  if (!(ctx)) { hc_set_error("null context"); return HC_ERR_ARG; }  \
  hc_guard hc_guard__(ctx);                                          \
  HC_CUDA(cudaSetDevice((ctx)->device))

void prof_begin(hc_ctx *ctx, cudaStream_t st, const char *name);
void prof_end(hc_ctx *ctx);

int arena_reset(hc_ctx *ctx);
void *arena_take(hc_ctx *ctx, size_t bytes);  // 256-B aligned; nullptr (+error set) on failure
static inline size_t hc_align(size_t b) { return (b + 255) & ~(size_t)255; }

__device__ __forceinline__ bool d_is_nodata(float v, float nodata) { return v == nodata || v != v; }

// monotone float -> uint map: a < b  <=>  ford(a) < ford(b)  (NaN never passed in)
__device__ __forceinline__ u32 d_ford(float f) {
  u32 b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float d_funord(u32 u) {
  u32 b = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
  return __uint_as_float(b);
}

// debug counters (cheap; read and reset with hc_debug_read)
#define HC_DBG(i, v) atomicAdd(&dbg[(i)], (unsigned long long)(v))

// neighbour tables, in each tool's own search order (see oracle/hydro_oracle.c)
__constant__ const int c_dr[2][8] = {{-1, 0, 1, 1, 1, 0, -1, -1}, {0, -1, -1, -1, 0, 1, 1, 1}};
__constant__ const int c_dc[2][8] = {{1, 1, 1, 0, -1, -1, -1, 0}, {1, 1, 0, -1, -1, -1, 0, 1}};
__constant__ const int c_code[2][8] = {{1, 2, 4, 8, 16, 32, 64, 128}, {1, 2, 3, 4, 5, 6, 7, 8}};

// code -> slot k (0..7), or -1
__device__ __forceinline__ int d_code_to_k(int table, int code) {
  if (table == HC_TABLE_WBT) {
    if (code == 0 || (code & (code - 1)) || code > 128) return -1;
    return __ffs(code) - 1;
  }
  return (code >= 1 && code <= 8) ? code - 1 : -1;
}

// internal pass entry points (device pointers, arena already begun by the caller when noted)
int fill_run(hc_ctx *ctx, const float *z, float *out, int64_t ny, int64_t nx, float nodata, int seed_mode,
             cudaStream_t st);
size_t fill_workspace(int64_t ny, int64_t nx);
int fill_d8_flats_accum_run(hc_ctx *ctx, const float *z, float *filled, uint8_t *dir, float *slope, uint32_t *acc, int64_t ny,
                            int64_t nx, float nodata, double dx, double dy, int seed_mode, int table, int resolve_flats,
                            cudaStream_t st);
int fill_d8_flats_run(hc_ctx *ctx, const float *z, float *filled, uint8_t *dir, float *slope, int64_t ny, int64_t nx,
                      float nodata, double dx, double dy, int seed_mode, int table, int resolve_flats, cudaStream_t st);
// exterior-nodata raster of the WBT seed rule (label.cu); ext comes from the arena, no arena_reset inside
int ext_nodata_mask(hc_ctx *ctx, const float *z, uint8_t *ext, int64_t ny, int64_t nx, float nodata,
                    int *has_interior, cudaStream_t st);
int d8_run(hc_ctx *ctx, const float *z, uint8_t *dir, float *slope, int64_t ny, int64_t nx, float nodata,
           double dx, double dy, int table, cudaStream_t st);
int flats_run(hc_ctx *ctx, const float *z, uint8_t *dir, int64_t ny, int64_t nx, float nodata, int table,
              cudaStream_t st);
size_t flats_workspace(int64_t ny, int64_t nx);
int d8_flats_run(hc_ctx *ctx, const float *z, uint8_t *dir, float *slope, int64_t ny, int64_t nx, float nodata,
                 double dx, double dy, int table, cudaStream_t st);
int accum_run(hc_ctx *ctx, const uint8_t *dir, uint32_t *acc, int64_t ny, int64_t nx, int table,
              cudaStream_t st);
size_t accum_workspace(int64_t ny, int64_t nx);
int accum_split_prepare(hc_ctx *ctx, const uint8_t *dir, uint32_t *acc, int64_t ny, int64_t nx);
int accum_split_tiles(hc_ctx *ctx, int table, const uint8_t *tile_class, int only_class, cudaStream_t st);
int accum_split_finish(hc_ctx *ctx, int table, cudaStream_t st);
// flat pass in two halves (whole raster): fast tier, then the slow tier for what it left pending
// per-tile lists of the NOFLOW cells of a D8 raster, as the flat pass consumes them: entries[tile * HC_NFCAP + k] = local
// index | HC_NF_HI, eq[...] = which of the 8 neighbours (3 x 3 row-major, centre skipped) have the same elevation,
// counts[tile] = NOFLOW cells in the tile (entries beyond HC_NFCAP are not stored)
struct hc_nf_lists { unsigned short *entries; uint8_t *eq; u32 *counts; };
int flats_whole_begin(hc_ctx *ctx, const float *z, const uint8_t *dirA, uint8_t *dirB, int64_t ny, int64_t nx, float nodata,
                      int table, cudaStream_t st, const uint8_t **tile_pending, int *n_pending, const hc_nf_lists *pre = nullptr);
int flats_whole_finish(hc_ctx *ctx, cudaStream_t st);
int label_run(hc_ctx *ctx, const uint8_t *mask, int32_t *lab, int64_t ny, int64_t nx, int conn, int32_t *nlab,
              cudaStream_t st);
size_t label_workspace(int64_t ny, int64_t nx);
// row-band phases (band.cu holds the extern "C" wrappers)
int64_t band_fill_edge_cap(int64_t nx);
int band_fill_local(hc_ctx *ctx, const float *z, int64_t nb, int64_t nx, float nodata, int seed_mode, int flags,
                    int64_t band_index, uint32_t *edges_out, cudaStream_t st, const uint8_t *ext = nullptr);
int band_fill_finish(hc_ctx *ctx, const uint32_t *all_edges, int64_t nbands, float *out, cudaStream_t st);
int d8_run_rows(hc_ctx *ctx, const float *z, uint8_t *dir, float *slope, int64_t ny, int64_t nx, float nodata,
                double dx, double dy, int table, hc_rows rw, cudaStream_t st, uint8_t *dir2 = nullptr);
int band_flats_begin(hc_ctx *ctx, const float *z, const uint8_t *dirA, uint8_t *dirB, int64_t nb, int64_t nx,
                     float nodata, int table, int flags, cudaStream_t st);
int band_flats_get_seams(hc_ctx *ctx, uint32_t *out, cudaStream_t st);
int band_flats_put_halos(hc_ctx *ctx, const uint32_t *top, const uint32_t *bot, int *changed, cudaStream_t st);
int band_flats_put_halos_flag(hc_ctx *ctx, const uint32_t *top, const uint32_t *bot, int *flag_dev, cudaStream_t st);
int band_flats_relax(hc_ctx *ctx, cudaStream_t st);
int band_flats_finish(hc_ctx *ctx, cudaStream_t st);
int band_accum_local(hc_ctx *ctx, const uint8_t *dir, uint32_t *acc, int64_t nb, int64_t nx, int table, int flags,
                     uint32_t *seam_out, cudaStream_t st);
int band_accum_finish(hc_ctx *ctx, const uint32_t *all_seams, int64_t nbands, int64_t band_index, uint32_t *acc,
                      cudaStream_t st);
int scp_run(hc_ctx *ctx, const float *z, float *out, int64_t ny, int64_t nx, float nodata, int breach, cudaStream_t st);
