// accum.cu — D8 flow accumulation (cell counts, self-inclusive).
//
// Replaces wbt.d8_flow_accumulation(pntr=True) (run_aws.py:544) and TauDEM AreaD8 -nc
// (run_aws.py:608-617): in-degree counting, then dependency-driven propagation.
//
// The dependency chain of a D8 forest is as long as the longest river, so a flat "one atomic per
// hop in global memory" walk is latency bound.  It is therefore done in two levels (the structure
// of Barnes 2017, "Parallel non-divergent flow accumulation for trillion cell DEMs", with a GPU
// tile as the unit):
//
//   A  k_acc_tile_local   per 64x64 tile, in shared memory: in-tile in-degrees, dependency walk
//                         with one shared-memory atomicAdd per hop, giving every cell's LOCAL
//                         count (written to the raster); each perimeter cell walks its in-tile
//                         path to the perimeter cell through which it leaves the tile.  Perimeter
//                         records go to 256 slots per tile.
//   B  k_acc_link*        the same dependency walk on the graph of tile-exit cells (one global
//                         atomic per TILE crossing instead of per cell), which also deposits the
//                         inflow every tile-entry cell receives from outside its tile.
//   C  k_acc_tile_final   counts are linear in the entry inflows: every entry cell with inflow
//                         adds it along its in-tile path (fire-and-forget shared reductions, no
//                         dependency waits) on top of the local counts.
//
// A cell's state word packs [count : high bits | pending-upstream counter : low bits]; a delivery
// is ONE atomicAdd of ((count << S) - 1): it credits the count and retires a dependency at once,
// and the returned old word tells the thread whether it was the last contributor, in which case
// it owns the target and walks on.  Integer adds only: the result is exact and order independent.
#include "common.cuh"

#define T HC_TILE
#define TH (T + 2)
#define NTHREADS 256
#define PSLOTS 256
#define NOSLOT 0xFFFFu
#define DN_NONE 0xFFFFu  // no downstream inside or outside the tile (terminal)
#define DN_EXIT 0xFFFEu  // downstream is a valid cell of another tile
// pdown of a slot whose downstream cell lies in a neighbouring BAND: top bit | side << 30 | column
#define BAND_EXIT 0x80000000u

__device__ __forceinline__ int perim_slot(int lr, int lc, int hr, int hc) {
  if (lr == 0) return lc;
  if (lr == hr - 1) return T + lc;
  if (lc == 0) return 2 * T + lr;
  if (lc == hc - 1) return 3 * T + lr;
  return -1;
}

// neighbour offsets from packed immediates (a dynamic index into __constant__ tables serialises
// when the lanes of a warp disagree): 2-bit fields holding offset + 1, slot k = 0..7
__host__ __device__ constexpr u32 pack_offsets(int a0, int a1, int a2, int a3, int a4, int a5, int a6, int a7) {
  return (u32)(a0 + 1) | (u32)(a1 + 1) << 2 | (u32)(a2 + 1) << 4 | (u32)(a3 + 1) << 6 | (u32)(a4 + 1) << 8 |
         (u32)(a5 + 1) << 10 | (u32)(a6 + 1) << 12 | (u32)(a7 + 1) << 14;
}
template <int TABLE>
__device__ __forceinline__ void k_to_drdc(int k, int &dr, int &dc) {
  constexpr u32 DRW = TABLE == HC_TABLE_WBT ? pack_offsets(-1, 0, 1, 1, 1, 0, -1, -1) : pack_offsets(0, -1, -1, -1, 0, 1, 1, 1);
  constexpr u32 DCW = TABLE == HC_TABLE_WBT ? pack_offsets(1, 1, 1, 0, -1, -1, -1, 0) : pack_offsets(1, 1, 0, -1, -1, -1, 0, 1);
  dr = (int)((DRW >> (2 * k)) & 3u) - 1;
  dc = (int)((DCW >> (2 * k)) & 3u) - 1;
}

// packed per-cell word of the local pass: [count:14 @17 | pending-upstream:4 @13 | downstream:13 @0]
#define W_DN 0x1FFFu
#define W_NONE 0x1FFFu   // no downstream (terminal)
#define W_EXIT 0x1FFEu   // downstream is a valid cell outside the tile
#define W_PEND_ONE (1u << 13)
#define W_PEND(w) (((w) >> 13) & 15u)
#define W_CNT(w) ((w) >> 17)
#define ACC_HOPS 4

// Phase A: per tile, local counts by a dependency walk in shared memory.  One 32-bit shared atomic
// per hop: it credits the count, retires a dependency, and its return value carries the target's
// own downstream pointer, so the walk needs no other load.  Writes the local counts to `acc` and the
// perimeter records (exit slot, downstream slot, exit count) phase B works on.
template <int TABLE>
__global__ void __launch_bounds__(NTHREADS)
k_acc_tile_local(const uint8_t *__restrict__ dir, u64 *__restrict__ pstate, unsigned short *__restrict__ pexit,
                 u32 *__restrict__ pdown, u32 *__restrict__ pinflow, uint32_t *__restrict__ acc, int ny, int nx,
                 hc_rows rw, const uint8_t *__restrict__ tile_class, int only_class) {
  __shared__ uint8_t sd[TH * TH];
  __shared__ u32 st[T * T];

  const int tile = blockIdx.y * gridDim.x + blockIdx.x;
  // (the whole-raster chain runs the tiles without pending flat cells while the flats' slow tier is still relaxing,
  //  and the other tiles afterwards: tile_class = the flat pass's per-tile "pending" flag)
  if (tile_class && (int)tile_class[tile] != only_class) return;
  const int r0 = blockIdx.y * T, c0 = blockIdx.x * T;
  const int hr = min(T, ny - r0), hc = min(T, nx - c0);
  const int tid = threadIdx.x;

#pragma unroll 6
  for (int i = tid; i < TH * TH; i += NTHREADS) {
    int lr = i / TH, lc = i - lr * TH;
    int r = r0 + lr - 1, c = c0 + lc - 1;
    bool inb = r >= rw.rlo && r < rw.rhi && c >= 0 && c < nx;
    uint8_t d = __ldg(dir + (inb ? (long long)r * nx + c : 0));  // r = -1 / ny: a band's halo rows
    sd[i] = inb ? d : (uint8_t)HC_DIR_NODATA;
  }
  __syncthreads();

  for (int i = tid; i < T * T; i += NTHREADS) {
    int lr = i >> 6, lc = i & (T - 1);
    u32 down = W_NONE, weight = 0;
    int d = sd[(lr + 1) * TH + lc + 1];
    if (d != HC_DIR_NODATA && lr < hr && lc < hc) {
      int k = d_code_to_k(TABLE, d);
      if (k >= 0) {
        int dr, dc;
        k_to_drdc<TABLE>(k, dr, dc);
        int nr = lr + dr, nc = lc + dc;
        if (sd[(nr + 1) * TH + nc + 1] != HC_DIR_NODATA)
          down = (nr >= 0 && nr < hr && nc >= 0 && nc < hc) ? (u32)(nr * T + nc) : W_EXIT;
      }
      weight = 1;
    }
    st[i] = down | (weight << 17);
  }
  __syncthreads();
  // in-tile in-degree by scatter: one shared atomic per cell
  for (int i = tid; i < T * T; i += NTHREADS) {
    u32 d = st[i] & W_DN;   // the low 13 bits never change
    if (d < W_EXIT) atomicAdd(&st[d], W_PEND_ONE);
  }
  __syncthreads();
  // sources = valid cells nobody in the tile drains into; remembered before any counter can reach 0 again
  u32 srcmask = 0;
#pragma unroll
  for (int j = 0; j < T * T / NTHREADS; j++) {
    u32 w = st[tid + j * NTHREADS];
    if (W_PEND(w) == 0u && W_CNT(w) != 0u) srcmask |= 1u << j;
  }
  __syncthreads();
  // one merged loop: a lane whose chain ended fetches its next source in the same iteration the other
  // lanes hop, so a warp iterates max over lanes of (sum of its chains), not sum over sources of the
  // longest chain among the lanes.  A chain that runs longer than ACC_HOPS hops (a stream: one lane would
  // keep its whole warp issuing for it) is parked as a continuation (count, next cell) and finished in a
  // second phase where the parked chains of the whole tile share one or two warps.
  __shared__ u32 s_nq;
  __shared__ u32 sq[T * T / ACC_HOPS];
  if (tid == 0) s_nq = 0;
  __syncthreads();
  {
    u32 a = 0, d = W_NONE;
    int hops = 0;
    for (;;) {
      if (d >= W_EXIT) {
        if (!srcmask) break;
        int j = __ffs(srcmask) - 1;
        srcmask &= srcmask - 1;
        u32 w = st[tid + j * NTHREADS];
        a = W_CNT(w);
        d = w & W_DN;
        hops = 0;
      }
      if (d < W_EXIT) {
        if (hops == ACC_HOPS) {
          sq[atomicAdd(&s_nq, 1u)] = (a << 13) | d;   // count < 2^14, cell < 2^13
          d = W_NONE;
        } else {
          u32 old = atomicAdd(&st[d], (a << 17) - W_PEND_ONE);
          if (W_PEND(old) != 1u) d = W_NONE;
          else { a += W_CNT(old); d = old & W_DN; hops++; }
        }
      }
    }
  }
  __syncthreads();
  for (u32 k = tid; k < s_nq; k += NTHREADS) {
    u32 e = sq[k];
    u32 a = e >> 13, d = e & W_DN;
    while (d < W_EXIT) {
      u32 old = atomicAdd(&st[d], (a << 17) - W_PEND_ONE);
      if (W_PEND(old) != 1u) break;
      a += W_CNT(old);
      d = old & W_DN;
    }
  }
  __syncthreads();

  // local counts -> raster (phase C adds the inflow from other tiles along the in-tile paths)
  for (int i = tid; i < T * T; i += NTHREADS) {
    int lr = i >> 6, lc = i & (T - 1);
    if (lr < hr && lc < hc) acc[(size_t)(r0 + lr) * nx + (c0 + lc)] = W_CNT(st[i]);
  }

  // perimeter records: thread t owns perimeter slot t
  u64 my_state = 0;
  u32 my_down = HC_NONE;
  unsigned short ex = NOSLOT;
  {
    int s = tid, lr = -1, lc = -1;
    if (s < T) { lr = 0; lc = s; }
    else if (s < 2 * T) { lr = hr - 1; lc = s - T; if (hr == 1) lr = -1; }
    else if (s < 3 * T) { lr = s - 2 * T; lc = 0; if (lr == 0 || lr >= hr - 1) lr = -1; }
    else { lr = s - 3 * T; lc = hc - 1; if (lr == 0 || lr >= hr - 1 || hc == 1) lr = -1; }
    if (lr >= 0 && lr < hr && lc >= 0 && lc < hc) {
      const int my_cell = lr * T + lc;
      u32 w = st[my_cell];
      if ((w & W_DN) == W_EXIT) {
        int k = d_code_to_k(TABLE, sd[(lr + 1) * TH + lc + 1]);
        int dr, dc;
        k_to_drdc<TABLE>(k, dr, dc);
        int gr = r0 + lr + dr, gc = c0 + lc + dc;
        if (gr < 0) my_down = BAND_EXIT | (u32)gc;
        else if (gr >= ny) my_down = BAND_EXIT | 0x40000000u | (u32)gc;
        else {
          int tyy = gr / T, txx = gc / T;
          int hr2 = min(T, ny - tyy * T), hc2 = min(T, nx - txx * T);
          int s2 = perim_slot(gr - tyy * T, gc - txx * T, hr2, hc2);
          my_down = (u32)((tyy * gridDim.x + txx) * PSLOTS + s2);
        }
        my_state = (u64)W_CNT(w) << 16;  // [local count : 48 | feeders : 16]
      }
      // the perimeter cell through which this cell's in-tile path leaves the tile (a D8 forest has no
      // cycles; the hop cap only guards against a malformed direction raster)
      int cur = my_cell;
      u32 d = w & W_DN;
      for (int hop = 0; d < W_EXIT && hop < T * T; hop++) { cur = (int)d; d = st[cur] & W_DN; }
      if (d == W_EXIT) ex = (unsigned short)perim_slot(cur >> 6, cur & (T - 1), hr, hc);
    }
  }
  size_t g = (size_t)tile * PSLOTS + tid;
  pexit[g] = ex;
  pdown[g] = my_down;
  pstate[g] = my_state;
  pinflow[g] = 0;
}

// Phase C: add the inflow every tile-entry cell receives from outside its tile along that cell's
// in-tile path (plain shared-memory reductions, no dependency waits), on top of the local counts.
template <int TABLE>
__global__ void __launch_bounds__(NTHREADS)
k_acc_tile_final(const uint8_t *__restrict__ dir, const u32 *__restrict__ pinflow, uint32_t *__restrict__ acc,
                 int ny, int nx) {
  __shared__ uint8_t sd[T * T];
  __shared__ unsigned short dn[T * T];
  __shared__ u32 sacc[T * T];
  const int tile = blockIdx.y * gridDim.x + blockIdx.x;
  const int r0 = blockIdx.y * T, c0 = blockIdx.x * T;
  const int hr = min(T, ny - r0), hc = min(T, nx - c0);
  const int tid = threadIdx.x;

  const u32 X = __ldg(&pinflow[(size_t)tile * PSLOTS + tid]);
  if (!__syncthreads_or(X != 0u)) return;  // nothing flows into this tile: the local counts are final

#pragma unroll 4
  for (int i = tid; i < T * T; i += NTHREADS) {
    int lr = i >> 6, lc = i & (T - 1);
    bool inb = lr < hr && lc < hc;
    size_t g = inb ? (size_t)(r0 + lr) * nx + (c0 + lc) : 0;
    uint8_t d = __ldg(&dir[g]);
    u32 a = __ldcg(&acc[g]);
    sd[i] = inb ? d : (uint8_t)HC_DIR_NODATA;
    sacc[i] = inb ? a : 0u;
  }
  __syncthreads();
  for (int i = tid; i < T * T; i += NTHREADS) {
    int lr = i >> 6, lc = i & (T - 1);
    unsigned short down = DN_NONE;
    int d = sd[i];
    if (d != HC_DIR_NODATA) {
      int k = d_code_to_k(TABLE, d);
      if (k >= 0) {
        int dr, dc;
        k_to_drdc<TABLE>(k, dr, dc);
        int nr = lr + dr, nc = lc + dc;
        if (nr >= 0 && nr < hr && nc >= 0 && nc < hc && sd[nr * T + nc] != HC_DIR_NODATA) down = (unsigned short)(nr * T + nc);
      }
    }
    dn[i] = down;
  }
  __syncthreads();
  if (X) {
    int s = tid, lr = -1, lc = -1;
    if (s < T) { lr = 0; lc = s; }
    else if (s < 2 * T) { lr = hr - 1; lc = s - T; if (hr == 1) lr = -1; }
    else if (s < 3 * T) { lr = s - 2 * T; lc = 0; if (lr == 0 || lr >= hr - 1) lr = -1; }
    else { lr = s - 3 * T; lc = hc - 1; if (lr == 0 || lr >= hr - 1 || hc == 1) lr = -1; }
    if (lr >= 0 && lr < hr && lc >= 0 && lc < hc && sd[lr * T + lc] != HC_DIR_NODATA) {
      int cur = lr * T + lc;
      for (int hop = 0; hop < T * T; hop++) {
        atomicAdd(&sacc[cur], X);
        unsigned short d = dn[cur];
        if (d >= DN_EXIT) break;
        cur = d;
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < T * T; i += NTHREADS) {
    int lr = i >> 6, lc = i & (T - 1);
    if (lr < hr && lc < hc) acc[(size_t)(r0 + lr) * nx + (c0 + lc)] = sacc[i];
  }
}

// B1: count, for every exit slot, how many exit slots feed it through the neighbouring tile
__global__ void __launch_bounds__(256)
k_acc_link_count(u64 *pstate, const unsigned short *__restrict__ pexit, const u32 *__restrict__ pdown, size_t nslots) {
  size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslots) return;
  u32 e = pdown[s];
  if (e == HC_NONE || (e & BAND_EXIT)) return;
  unsigned short x2 = pexit[e];
  if (x2 == NOSLOT) return;
  atomicAdd(&pstate[(size_t)(e / PSLOTS) * PSLOTS + x2], 1ull);
}

// B1b: exits nobody feeds are the sources of the link walk
__global__ void __launch_bounds__(256)
k_acc_link_mark(u64 *pstate, const u32 *__restrict__ pdown, size_t nslots) {
  size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslots) return;
  if (pdown[s] == HC_NONE) return;
  u64 w = pstate[s];
  if ((w & 0xFFFFull) == 0) pstate[s] = w | 0xFFFFull;
}

// B2: dependency walk over tile crossings
__global__ void __launch_bounds__(256)
k_acc_link_walk(u64 *pstate, const unsigned short *__restrict__ pexit, const u32 *__restrict__ pdown,
                u32 *pinflow, size_t nslots, u32 *bout, int nx) {
  size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslots) return;
  if (pdown[s] == HC_NONE) return;
  u64 w = pstate[s];
  if ((w & 0xFFFFull) != 0xFFFFull) return;
  u64 total = w >> 16;
  size_t x = s;
  for (;;) {
    u32 e = pdown[x];
    if (e & BAND_EXIT) {  // leaves the band: credit the neighbouring band's seam cell
      atomicAdd(&bout[(size_t)((e >> 30) & 1u) * nx + (e & 0x3FFFFFFFu)], (u32)total);
      return;
    }
    atomicAdd(&pinflow[e], (u32)total);
    unsigned short x2 = pexit[e];
    if (x2 == NOSLOT) return;
    size_t gx = (size_t)(e / PSLOTS) * PSLOTS + x2;
    u64 old = atomicAdd(&pstate[gx], (total << 16) - 1ull);
    if ((old & 0xFFFFull) != 1ull) return;
    total += old >> 16;
    x = gx;
  }
}

size_t accum_workspace(int64_t ny, int64_t nx) {
  size_t tiles = (size_t)((ny + T - 1) / T) * ((nx + T - 1) / T);
  return hc_align(tiles * PSLOTS * 8) + hc_align(tiles * PSLOTS * 2) + 2 * hc_align(tiles * PSLOTS * 4);
}

// phases A + B (tile-local counts, then the walk over tile crossings); state left in ctx->band
template <int TABLE>
static int accum_local_impl(hc_ctx *ctx, const uint8_t *dir, uint32_t *acc, int64_t ny, int64_t nx, hc_rows rw,
                            cudaStream_t st) {
  hc_band_state *bs = &ctx->band;
  const int tx = (int)((nx + T - 1) / T), ty = (int)((ny + T - 1) / T);
  const size_t nslots = (size_t)tx * ty * PSLOTS;
  u64 *pstate = (u64 *)arena_take(ctx, nslots * 8);
  unsigned short *pexit = (unsigned short *)arena_take(ctx, nslots * 2);
  u32 *pdown = (u32 *)arena_take(ctx, nslots * 4);
  u32 *pinflow = (u32 *)arena_take(ctx, nslots * 4);
  u32 *bout = (u32 *)arena_take(ctx, (size_t)2 * nx * 4);
  if (!pstate || !pexit || !pdown || !pinflow || !bout) return HC_ERR_NOMEM;
  bs->pstate = pstate; bs->pexit = pexit; bs->pdown = pdown; bs->pinflow = pinflow; bs->bout = bout;
  bs->dir = dir; bs->nb = ny; bs->nx = nx;
  if (rw.top || rw.bot) HC_CUDA(cudaMemsetAsync(bout, 0, (size_t)2 * nx * 4, st));
  dim3 tg(tx, ty);
  unsigned sb = (unsigned)((nslots + 255) / 256);
  HC_PROF(ctx, st, "k_acc_tile_local");
  k_acc_tile_local<TABLE><<<tg, NTHREADS, 0, st>>>(dir, pstate, pexit, pdown, pinflow, acc, (int)ny, (int)nx, rw, nullptr, 0);
  HC_LAUNCH_CHECK(ctx);
  HC_PROF(ctx, st, "k_acc_link_count");
  k_acc_link_count<<<sb, 256, 0, st>>>(pstate, pexit, pdown, nslots);
  HC_LAUNCH_CHECK(ctx);
  HC_PROF(ctx, st, "k_acc_link_mark");
  k_acc_link_mark<<<sb, 256, 0, st>>>(pstate, pdown, nslots);
  HC_LAUNCH_CHECK(ctx);
  HC_PROF(ctx, st, "k_acc_link_walk");
  k_acc_link_walk<<<sb, 256, 0, st>>>(pstate, pexit, pdown, pinflow, nslots, bout, (int)nx);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

template <int TABLE>
static int accum_final_impl(hc_ctx *ctx, uint32_t *acc, hc_rows rw, cudaStream_t st) {
  hc_band_state *bs = &ctx->band;
  const int64_t ny = bs->nb, nx = bs->nx;
  dim3 tg((unsigned)((nx + T - 1) / T), (unsigned)((ny + T - 1) / T));
  HC_PROF(ctx, st, "k_acc_tile_final");
  (void)rw;
  k_acc_tile_final<TABLE><<<tg, NTHREADS, 0, st>>>(bs->dir, bs->pinflow, acc, (int)ny, (int)nx);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

// ---- split form for the whole-raster chain (fill.cu fill_d8_flats_accum_run): phase A per tile class on any stream,
// ---- then links + phase C.  No arena_reset: the workspace is taken from the current arena generation.
template <int TABLE>
static int accum_tiles_impl(hc_ctx *ctx, const uint8_t *tile_class, int only_class, cudaStream_t st) {
  hc_band_state *bs = &ctx->band;
  const int64_t ny = bs->nb, nx = bs->nx;
  dim3 tg((unsigned)((nx + T - 1) / T), (unsigned)((ny + T - 1) / T));
  HC_PROF(ctx, st, "k_acc_tile_local");
  k_acc_tile_local<TABLE><<<tg, NTHREADS, 0, st>>>(bs->dir, bs->pstate, bs->pexit, bs->pdown, bs->pinflow, bs->acc_out, (int)ny,
                                                   (int)nx, hc_make_rows(ny, 0), tile_class, only_class);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

int accum_split_prepare(hc_ctx *ctx, const uint8_t *dir, uint32_t *acc, int64_t ny, int64_t nx) {
  hc_band_state *bs = &ctx->band;
  const size_t nslots = (size_t)((nx + T - 1) / T) * ((ny + T - 1) / T) * PSLOTS;
  bs->pstate = (u64 *)arena_take(ctx, nslots * 8);
  bs->pexit = (unsigned short *)arena_take(ctx, nslots * 2);
  bs->pdown = (u32 *)arena_take(ctx, nslots * 4);
  bs->pinflow = (u32 *)arena_take(ctx, nslots * 4);
  bs->bout = (u32 *)arena_take(ctx, (size_t)2 * nx * 4);
  if (!bs->pstate || !bs->pexit || !bs->pdown || !bs->pinflow || !bs->bout) return HC_ERR_NOMEM;
  bs->dir = dir; bs->nb = ny; bs->nx = nx; bs->acc_out = acc;
  return HC_OK;
}

int accum_split_tiles(hc_ctx *ctx, int table, const uint8_t *tile_class, int only_class, cudaStream_t st) {
  return table == HC_TABLE_WBT ? accum_tiles_impl<HC_TABLE_WBT>(ctx, tile_class, only_class, st)
                               : accum_tiles_impl<HC_TABLE_TAUDEM>(ctx, tile_class, only_class, st);
}

int accum_split_finish(hc_ctx *ctx, int table, cudaStream_t st) {
  hc_band_state *bs = &ctx->band;
  const size_t nslots = (size_t)((bs->nx + T - 1) / T) * ((bs->nb + T - 1) / T) * PSLOTS;
  unsigned sb = (unsigned)((nslots + 255) / 256);
  HC_PROF(ctx, st, "k_acc_link_count");
  k_acc_link_count<<<sb, 256, 0, st>>>(bs->pstate, bs->pexit, bs->pdown, nslots);
  HC_LAUNCH_CHECK(ctx);
  HC_PROF(ctx, st, "k_acc_link_mark");
  k_acc_link_mark<<<sb, 256, 0, st>>>(bs->pstate, bs->pdown, nslots);
  HC_LAUNCH_CHECK(ctx);
  HC_PROF(ctx, st, "k_acc_link_walk");
  k_acc_link_walk<<<sb, 256, 0, st>>>(bs->pstate, bs->pexit, bs->pdown, bs->pinflow, nslots, bs->bout, (int)bs->nx);
  HC_LAUNCH_CHECK(ctx);
  hc_rows rw = hc_make_rows(bs->nb, 0);
  return table == HC_TABLE_WBT ? accum_final_impl<HC_TABLE_WBT>(ctx, bs->acc_out, rw, st)
                               : accum_final_impl<HC_TABLE_TAUDEM>(ctx, bs->acc_out, rw, st);
}

int accum_run(hc_ctx *ctx, const uint8_t *dir, uint32_t *acc, int64_t ny, int64_t nx, int table,
              cudaStream_t st) {
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (table != HC_TABLE_WBT && table != HC_TABLE_TAUDEM) { hc_set_error("accum: bad table %d", table); return HC_ERR_ARG; }
  HC_TRY(arena_reset(ctx));
  hc_rows rw = hc_make_rows(ny, 0);
  if (table == HC_TABLE_WBT) {
    HC_TRY(accum_local_impl<HC_TABLE_WBT>(ctx, dir, acc, ny, nx, rw, st));
    return accum_final_impl<HC_TABLE_WBT>(ctx, acc, rw, st);
  }
  HC_TRY(accum_local_impl<HC_TABLE_TAUDEM>(ctx, dir, acc, ny, nx, rw, st));
  return accum_final_impl<HC_TABLE_TAUDEM>(ctx, acc, rw, st);
}

// ================================================================================================
// row bands (the one-exchange scheme of Barnes 2017 with a band as the unit).
//   local   phases A + B with flow that leaves the band parked in bout[side][col] (credited to the
//           neighbouring band's seam cell); then, for every own seam cell, the walk over tile
//           crossings tells where a unit of inflow arriving there leaves the band again (exit_of).
//           The band publishes seam_out[side][{exit_of, bout}][nx].
//   finish  every rank solves the forest over all seam cells (X = inflow a seam cell receives from
//           the other side = neighbour's bout + sum of X of the seam cells that exit into it), adds
//           X along each own seam cell's path through the tile-entry slots, and runs phase C.
// Integer adds only => identical to the whole-raster result.
// ================================================================================================
#define SEAM_NONE 0xFFFFFFFFu

__global__ void __launch_bounds__(256)
k_acc_seam_exit(const unsigned short *__restrict__ pexit, const u32 *__restrict__ pdown, const uint8_t *__restrict__ dir,
                int ny, int nx, int tx, hc_rows rw, u32 *__restrict__ seam_out, const u32 *__restrict__ bout, u32 *err) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * nx) return;
  const int side = i / nx, c = i - side * nx;
  u32 *o = seam_out + (size_t)side * 2 * nx;
  o[nx + c] = bout[i];
  u32 ex = SEAM_NONE;
  if ((side == 0 && rw.top) || (side == 1 && rw.bot)) {
    const int r = side ? ny - 1 : 0;
    if (dir[(size_t)r * nx + c] != HC_DIR_NODATA) {
      int tyy = r / T, txx = c / T;
      int hr = min(T, ny - tyy * T), hc = min(T, nx - txx * T);
      u32 t = (u32)(tyy * tx + txx);
      int s = perim_slot(r - tyy * T, c - txx * T, hr, hc);
      for (int hop = 0;; hop++) {
        if (hop > (1 << 26)) { *err = 1u; break; }  // a cycle: not a D8 forest
        unsigned short x = pexit[(size_t)t * PSLOTS + s];
        if (x == NOSLOT) break;
        u32 d = pdown[(size_t)t * PSLOTS + x];
        if (d & BAND_EXIT) { ex = d & 0x7FFFFFFFu; break; }  // side << 30 | column of the receiving cell
        t = d / PSLOTS;
        s = (int)(d % PSLOTS);
      }
    }
  }
  o[c] = ex;
}

// global seam forest.  Node g = (band*2 + side)*nx + c.  all[band][side][{exit_of,bout}][nx].
__device__ __forceinline__ long long seam_recv(u32 ex, long long band, int nx) {
  // exit through the top (side bit 0) lands on the bottom seam of band-1, through the bottom on the top of band+1
  int via_bottom = (ex >> 30) & 1;
  long long col = ex & 0x3FFFFFFFu;
  return via_bottom ? ((band + 1) * 2 + 0) * (long long)nx + col : ((band - 1) * 2 + 1) * (long long)nx + col;
}

__global__ void k_seam_init(const u32 *__restrict__ all, int nbands, int nx, u64 *__restrict__ state) {
  long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long M = (long long)nbands * 2 * nx;
  if (g >= M) return;
  long long band = g / (2ll * nx);
  int side = (int)((g / nx) & 1), c = (int)(g % nx);
  // inflow from the neighbouring band's local sources: its bout record aimed at me
  long long nb_band = side ? band + 1 : band - 1;
  u64 x0 = 0;
  if (nb_band >= 0 && nb_band < nbands) x0 = all[((size_t)nb_band * 2 + (side ? 0 : 1)) * 2 * nx + nx + c];
  state[g] = x0 << 24;
}
__global__ void k_seam_count(const u32 *__restrict__ all, int nbands, int nx, u64 *__restrict__ state) {
  long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long M = (long long)nbands * 2 * nx;
  if (g >= M) return;
  long long band = g / (2ll * nx);
  int side = (int)((g / nx) & 1), c = (int)(g % nx);
  u32 ex = all[((size_t)band * 2 + side) * 2 * nx + c];
  if (ex == SEAM_NONE) return;
  long long t = seam_recv(ex, band, nx);
  if (t < 0 || t >= M) return;
  atomicAdd(&state[t], 1ull);
}
__global__ void k_seam_mark(const u32 *__restrict__ all, int nbands, int nx, u64 *__restrict__ state) {
  long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long M = (long long)nbands * 2 * nx;
  if (g >= M) return;
  u64 w = state[g];
  if ((w & 0xFFFFFFull) == 0) state[g] = w | 0xFFFFFFull;  // a source of the seam walk
}
__global__ void k_seam_walk(const u32 *__restrict__ all, int nbands, int nx, u64 *__restrict__ state) {
  long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long M = (long long)nbands * 2 * nx;
  if (g >= M) return;
  u64 w = state[g];
  if ((w & 0xFFFFFFull) != 0xFFFFFFull) return;
  u64 total = w >> 24;
  long long x = g;
  for (;;) {
    long long band = x / (2ll * nx);
    int side = (int)((x / nx) & 1), c = (int)(x % nx);
    u32 ex = all[((size_t)band * 2 + side) * 2 * nx + c];
    if (ex == SEAM_NONE) return;
    long long t = seam_recv(ex, band, nx);
    if (t < 0 || t >= M) return;
    u64 old = atomicAdd(&state[t], (total << 24) - 1ull);
    if ((old & 0xFFFFFFull) != 1ull) return;
    total += old >> 24;
    x = t;
  }
}

// inject X (inflow from the other band) at an own seam cell: the cell's own entry slot takes it directly, and the
// exit slot its in-tile path leads to receives it as the "local count" of a second walk over the tile crossings
// (k_acc_link_count / mark / walk on pstate2), which then spreads it to every entry slot further down — each slot
// visited once.  (Walking every seam cell's path separately costs seam cells x path length atomics, thousands of them
// on the same river slots: that was 14 ms on the downstream bands of an 8-band run.)
__global__ void __launch_bounds__(256)
k_acc_seam_inject(const unsigned short *__restrict__ pexit, u32 *pinflow, u64 *pstate2, const uint8_t *__restrict__ dir, int ny,
                  int nx, int tx, hc_rows rw, const u64 *__restrict__ state, long long gbase) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * nx) return;
  const int side = i / nx, c = i - side * nx;
  if (!((side == 0 && rw.top) || (side == 1 && rw.bot))) return;
  u32 X = (u32)(state[gbase + i] >> 24);
  if (!X) return;
  const int r = side ? ny - 1 : 0;
  if (dir[(size_t)r * nx + c] == HC_DIR_NODATA) return;
  int tyy = r / T, txx = c / T;
  int hr = min(T, ny - tyy * T), hc = min(T, nx - txx * T);
  size_t t = (size_t)(tyy * tx + txx);
  int s = perim_slot(r - tyy * T, c - txx * T, hr, hc);
  atomicAdd(&pinflow[t * PSLOTS + s], X);
  unsigned short x = pexit[t * PSLOTS + s];
  if (x != NOSLOT) atomicAdd(&pstate2[t * PSLOTS + x], (u64)X << 16);
}

int band_accum_local(hc_ctx *ctx, const uint8_t *dir, uint32_t *acc, int64_t nb, int64_t nx, int table, int flags,
                     uint32_t *seam_out, cudaStream_t st) {
  if (table != HC_TABLE_WBT && table != HC_TABLE_TAUDEM) { hc_set_error("accum: bad table %d", table); return HC_ERR_ARG; }
  if (nx >= (1 << 30)) { hc_set_error("band too wide"); return HC_ERR_TOO_LARGE; }
  if (ctx->h_mail[8]) {
    ctx->h_mail[8] = 0;
    hc_set_error("band accumulation: the previous direction raster contained a cycle (not a D8 forest)");
    return HC_ERR_ARG;
  }
  HC_TRY(arena_reset(ctx));
  hc_band_state *bs = &ctx->band;
  hc_rows rw = hc_make_rows(nb, flags);
  bs->flags = flags; bs->table = table;
  if (table == HC_TABLE_WBT) HC_TRY(accum_local_impl<HC_TABLE_WBT>(ctx, dir, acc, nb, nx, rw, st));
  else HC_TRY(accum_local_impl<HC_TABLE_TAUDEM>(ctx, dir, acc, nb, nx, rw, st));
  if (!(flags & 3)) HC_CUDA(cudaMemsetAsync(bs->bout, 0, (size_t)2 * nx * 4, st));
  u32 *err = (u32 *)arena_take(ctx, 256);
  if (!err) return HC_ERR_NOMEM;
  HC_CUDA(cudaMemsetAsync(err, 0, 4, st));
  const int tx = (int)((nx + T - 1) / T);
  HC_PROF(ctx, st, "k_acc_seam_exit");
  k_acc_seam_exit<<<(unsigned)((2 * nx + 255) / 256), 256, 0, st>>>(bs->pexit, bs->pdown, dir, (int)nb, (int)nx, tx, rw, seam_out, bs->bout, err);
  HC_LAUNCH_CHECK(ctx);
  bs->acc_err = err;   // checked in the finish phase (one read-back there instead of a sync here)
  bs->acc_ready = 1;
  return HC_OK;
}

int band_accum_finish(hc_ctx *ctx, const uint32_t *all_seams, int64_t nbands, int64_t band_index, uint32_t *acc,
                      cudaStream_t st) {
  hc_band_state *bs = &ctx->band;
  if (!bs->acc_ready) { hc_set_error("hc_band_accum_finish without hc_band_accum_local"); return HC_ERR_ARG; }
  bs->acc_ready = 0;
  const int64_t nb = bs->nb, nx = bs->nx;
  hc_rows rw = hc_make_rows(nb, bs->flags);
  if (bs->flags & 3) {
    const long long M = (long long)nbands * 2 * nx;
    u64 *state = (u64 *)arena_take(ctx, (size_t)M * 8);
    if (!state) return HC_ERR_NOMEM;
    unsigned gb = (unsigned)((M + 255) / 256);
    HC_PROF(ctx, st, "k_seam_init");
    k_seam_init<<<gb, 256, 0, st>>>(all_seams, (int)nbands, (int)nx, state);
    HC_LAUNCH_CHECK(ctx);
    HC_PROF(ctx, st, "k_seam_count");
    k_seam_count<<<gb, 256, 0, st>>>(all_seams, (int)nbands, (int)nx, state);
    HC_LAUNCH_CHECK(ctx);
    HC_PROF(ctx, st, "k_seam_mark");
    k_seam_mark<<<gb, 256, 0, st>>>(all_seams, (int)nbands, (int)nx, state);
    HC_LAUNCH_CHECK(ctx);
    HC_PROF(ctx, st, "k_seam_walk");
    k_seam_walk<<<gb, 256, 0, st>>>(all_seams, (int)nbands, (int)nx, state);
    HC_LAUNCH_CHECK(ctx);
    const int tx = (int)((nx + T - 1) / T), ty = (int)((nb + T - 1) / T);
    const size_t nslots = (size_t)tx * ty * PSLOTS;
    u64 *pstate2 = (u64 *)arena_take(ctx, nslots * 8);
    u32 *bout2 = (u32 *)arena_take(ctx, (size_t)2 * nx * 4);   // the second walk's band exits are already in X
    if (!pstate2 || !bout2) return HC_ERR_NOMEM;
    HC_CUDA(cudaMemsetAsync(pstate2, 0, nslots * 8, st));
    HC_PROF(ctx, st, "k_acc_seam_inject");
    k_acc_seam_inject<<<(unsigned)((2 * nx + 255) / 256), 256, 0, st>>>(bs->pexit, bs->pinflow, pstate2, bs->dir, (int)nb, (int)nx, tx, rw, state, (long long)band_index * 2 * nx);
    HC_LAUNCH_CHECK(ctx);
    unsigned sb = (unsigned)((nslots + 255) / 256);
    HC_PROF(ctx, st, "k_acc_link_count");
    k_acc_link_count<<<sb, 256, 0, st>>>(pstate2, bs->pexit, bs->pdown, nslots);
    HC_LAUNCH_CHECK(ctx);
    HC_PROF(ctx, st, "k_acc_link_mark");
    k_acc_link_mark<<<sb, 256, 0, st>>>(pstate2, bs->pdown, nslots);
    HC_LAUNCH_CHECK(ctx);
    HC_PROF(ctx, st, "k_acc_link_walk");
    k_acc_link_walk<<<sb, 256, 0, st>>>(pstate2, bs->pexit, bs->pdown, bs->pinflow, nslots, bout2, (int)nx);
    HC_LAUNCH_CHECK(ctx);
  }
  if (bs->table == HC_TABLE_WBT) HC_TRY(accum_final_impl<HC_TABLE_WBT>(ctx, acc, rw, st));
  else HC_TRY(accum_final_impl<HC_TABLE_TAUDEM>(ctx, acc, rw, st));
  // deferred check of the seam walk's cycle guard: the copy is queued behind the kernels, the flag is looked at
  // at the start of the NEXT banded accumulation on this context (no host synchronisation on the hot path)
  HC_CUDA(cudaMemcpyAsync(ctx->h_mail + 8, bs->acc_err, 4, cudaMemcpyDeviceToHost, st));
  return HC_OK;
}
