// common.cuh — context, workspace arena, error plumbing shared by every pass of libhydrocore.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/hydrocore.h"

#define HC_TILE 64             // square tile edge used by the tiled passes (fill, flats)
#define HC_TAG 0x80000000u     // top bit marks a resolved basin id in the pointer/label array
#define HC_LAB_NODATA 0x7FFFFFFFu
#define HC_NONE 0xFFFFFFFFu

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

#define HC_MAX_BLOCKS 64
struct hc_block { char *p; size_t cap, off; };

struct hc_ctx {
  int device;
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
  // grow-only device staging buffers of the *_host entry points (z, filled, dir, slope, acc)
  void *io[5];
  size_t io_cap[5];
  unsigned long long *d_dbg;  // 16 device work counters
  // per-kernel event timing (off by default)
  int prof_on, prof_open;
  int prof_n, prof_cap;
  cudaEvent_t *prof_ev;      // 2 per record
  const char **prof_name;
  cudaStream_t prof_stream;
};

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
int label_run(hc_ctx *ctx, const uint8_t *mask, int32_t *lab, int64_t ny, int64_t nx, int conn, int32_t *nlab,
              cudaStream_t st);
size_t label_workspace(int64_t ny, int64_t nx);
int scp_run(hc_ctx *ctx, const float *z, float *out, int64_t ny, int64_t nx, float nodata, int breach, cudaStream_t st);
