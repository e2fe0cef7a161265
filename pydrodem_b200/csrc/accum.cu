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
//   A  k_acc_tile<false>  per 64x64 tile, in shared memory: in-tile in-degrees, dependency walk
//                         with one shared-memory atomicAdd per hop, giving every cell's LOCAL
//                         count; pointer jumping gives, for each perimeter cell, the perimeter
//                         cell through which its in-tile path leaves the tile.  Only perimeter
//                         records leave the SM (256 slots per tile).
//   B  k_acc_link*        the same dependency walk on the graph of tile-exit cells (one global
//                         atomic per TILE crossing instead of per cell), which also deposits the
//                         inflow every tile-entry cell receives from outside its tile.
//   C  k_acc_tile<true>   phase A again with the entry cells' weights raised by their inflow;
//                         writes the final uint32 raster.
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

__device__ __forceinline__ int perim_slot(int lr, int lc, int hr, int hc) {
  if (lr == 0) return lc;
  if (lr == hr - 1) return T + lc;
  if (lc == 0) return 2 * T + lr;
  if (lc == hc - 1) return 3 * T + lr;
  return -1;
}

template <int TABLE, bool FINAL>
__global__ void __launch_bounds__(NTHREADS)
k_acc_tile(const uint8_t *__restrict__ dir, u64 *__restrict__ pstate, unsigned short *__restrict__ pexit,
           u32 *__restrict__ pdown, u32 *__restrict__ pinflow, uint32_t *__restrict__ acc, int ny, int nx,
           unsigned long long *dbg) {
  // 32-bit shared-memory atomics only (64-bit ones are CAS spin loops):
  //   local pass : st = [count:24 | pending:8]  (an in-tile count is at most 4096)
  //   final pass : counts reach 2^31, so count (sacc) and pending counter (st) are separate words
  // (measured alternatives that were NOT faster: in-degree by 8-neighbour gather, plain updates for
  //  single-contributor cells, one-hop-per-iteration walk with immediate source refill)
  __shared__ uint8_t sd[TH * TH];
  __shared__ unsigned short dn[T * T];
  __shared__ u32 st[T * T];
  __shared__ u32 sacc[FINAL ? T * T : 1];
  __shared__ unsigned short jmp[FINAL ? 1 : T * T];

  const int tile = blockIdx.y * gridDim.x + blockIdx.x;
  const int r0 = blockIdx.y * T, c0 = blockIdx.x * T;
  const int hr = min(T, ny - r0), hc = min(T, nx - c0);
  const int tid = threadIdx.x;

#pragma unroll 6
  for (int i = tid; i < TH * TH; i += NTHREADS) {
    int lr = i / TH, lc = i - lr * TH;
    int r = r0 + lr - 1, c = c0 + lc - 1;
    bool inb = r >= 0 && r < ny && c >= 0 && c < nx;
    uint8_t d = __ldg(&dir[inb ? (size_t)r * nx + c : 0]);
    sd[i] = inb ? d : (uint8_t)HC_DIR_NODATA;
  }
  __syncthreads();

  // downstream slot and initial state of every cell (pending counters start at 0)
  for (int i = tid; i < T * T; i += NTHREADS) {
    int lr = i >> 6, lc = i & (T - 1);
    unsigned short down = DN_NONE;
    u32 weight = 0;
    int ci = (lr + 1) * TH + lc + 1;
    int d = sd[ci];
    if (d != HC_DIR_NODATA && lr < hr && lc < hc) {
      int k = d_code_to_k(TABLE, d);
      if (k >= 0) {
        int nr = lr + c_dr[TABLE][k], nc = lc + c_dc[TABLE][k];
        if (sd[(nr + 1) * TH + nc + 1] != HC_DIR_NODATA)
          down = (nr >= 0 && nr < hr && nc >= 0 && nc < hc) ? (unsigned short)(nr * T + nc) : (unsigned short)DN_EXIT;
      }
      weight = 1;
      if (FINAL) {
        int s = perim_slot(lr, lc, hr, hc);
        if (s >= 0) weight += __ldcg(&pinflow[(size_t)tile * PSLOTS + s]);
      }
    }
    dn[i] = down;
    if (FINAL) { st[i] = 0; sacc[i] = weight; }
    else st[i] = weight << 8;
  }
  __syncthreads();
  // in-tile in-degree by scatter: one shared atomic per cell
  for (int i = tid; i < T * T; i += NTHREADS) {
    unsigned short d = dn[i];
    if (d < DN_EXIT) atomicAdd(&st[d], 1u);
  }
  __syncthreads();
  // sources = cells nobody in the tile drains into; remembered before any counter can reach 0 again
  u32 srcmask = 0;
#pragma unroll
  for (int j = 0; j < T * T / NTHREADS; j++) {
    int i = tid + j * NTHREADS;
    if ((st[i] & 0xFFu) == 0u && sd[((i >> 6) + 1) * TH + (i & (T - 1)) + 1] != HC_DIR_NODATA) srcmask |= 1u << j;
  }
  __syncthreads();

  // dependency walk from the in-tile sources
  for (int j = 0; j < T * T / NTHREADS; j++) {
    if (!(srcmask & (1u << j))) continue;
    int i = tid + j * NTHREADS;
    int cur = i;
    if (FINAL) {
      u32 a = sacc[i];
      for (;;) {
        unsigned short d = dn[cur];
        if (d >= DN_EXIT) break;
        atomicAdd(&sacc[d], a);
        __threadfence_block();
        u32 old = atomicSub(&st[d], 1u);
        if (old != 1u) break;
        a = *((volatile u32 *)&sacc[d]);
        cur = d;
      }
    } else {
      u32 a = st[i] >> 8;
      for (;;) {
        unsigned short d = dn[cur];
        if (d >= DN_EXIT) break;
        u32 old = atomicAdd(&st[d], (a << 8) - 1u);
        if ((old & 0xFFu) != 1u) break;
        a += old >> 8;
        cur = d;
      }
    }
  }
  __syncthreads();

  if (FINAL) {
    for (int i = tid; i < T * T; i += NTHREADS) {
      int lr = i >> 6, lc = i & (T - 1);
      if (lr < hr && lc < hc) acc[(size_t)(r0 + lr) * nx + (c0 + lc)] = sacc[i];
    }
    return;
  }

  // phase A epilogue: perimeter records
  u64 my_state = 0;
  u32 my_down = HC_NONE;
  int my_cell = -1;
  {
    // thread t owns perimeter slot t
    int s = tid, lr = -1, lc = -1;
    if (s < T) { lr = 0; lc = s; }
    else if (s < 2 * T) { lr = hr - 1; lc = s - T; if (hr == 1) lr = -1; }
    else if (s < 3 * T) { lr = s - 2 * T; lc = 0; if (lr == 0 || lr >= hr - 1) lr = -1; }
    else { lr = s - 3 * T; lc = hc - 1; if (lr == 0 || lr >= hr - 1 || hc == 1) lr = -1; }
    if (lr >= 0 && lr < hr && lc >= 0 && lc < hc) {
      my_cell = lr * T + lc;
      if (dn[my_cell] == DN_EXIT) {
        int k = d_code_to_k(TABLE, sd[(lr + 1) * TH + lc + 1]);
        int gr = r0 + lr + c_dr[TABLE][k], gc = c0 + lc + c_dc[TABLE][k];
        int tyy = gr / T, txx = gc / T;
        int hr2 = min(T, ny - tyy * T), hc2 = min(T, nx - txx * T);
        int s2 = perim_slot(gr - tyy * T, gc - txx * T, hr2, hc2);
        my_down = (u32)((tyy * gridDim.x + txx) * PSLOTS + s2);
        my_state = (u64)(st[my_cell] >> 8) << 16;  // [local count : 48 | feeders : 16]
      }
    }
  }
  for (int i = tid; i < T * T; i += NTHREADS) {
    unsigned short d = dn[i];
    jmp[i] = d >= DN_EXIT ? (unsigned short)i : d;
  }
  __syncthreads();
  for (;;) {
    int changed = 0;
    for (int i = tid; i < T * T; i += NTHREADS) {
      unsigned short p = jmp[i];
      unsigned short pp = jmp[p];
      if (pp != p) { jmp[i] = pp; changed = 1; }
    }
    if (!__syncthreads_or(changed)) break;
  }
  {
    size_t g = (size_t)tile * PSLOTS + tid;
    unsigned short ex = NOSLOT;
    if (my_cell >= 0) {
      int root = jmp[my_cell];
      if (dn[root] == DN_EXIT) ex = (unsigned short)perim_slot(root / T, root % T, hr, hc);
    }
    pexit[g] = ex;
    pdown[g] = my_down;
    pstate[g] = my_state;
    pinflow[g] = 0;
  }
}

// B1: count, for every exit slot, how many exit slots feed it through the neighbouring tile
__global__ void __launch_bounds__(256)
k_acc_link_count(u64 *pstate, const unsigned short *__restrict__ pexit, const u32 *__restrict__ pdown, size_t nslots) {
  size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslots) return;
  u32 e = pdown[s];
  if (e == HC_NONE) return;
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
                u32 *pinflow, size_t nslots) {
  size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslots) return;
  if (pdown[s] == HC_NONE) return;
  u64 w = pstate[s];
  if ((w & 0xFFFFull) != 0xFFFFull) return;
  u64 total = w >> 16;
  size_t x = s;
  for (;;) {
    u32 e = pdown[x];
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

template <int TABLE>
static int accum_impl(hc_ctx *ctx, const uint8_t *dir, uint32_t *acc, int64_t ny, int64_t nx, cudaStream_t st) {
  const int tx = (int)((nx + T - 1) / T), ty = (int)((ny + T - 1) / T);
  const size_t nslots = (size_t)tx * ty * PSLOTS;
  u64 *pstate = (u64 *)arena_take(ctx, nslots * 8);
  unsigned short *pexit = (unsigned short *)arena_take(ctx, nslots * 2);
  u32 *pdown = (u32 *)arena_take(ctx, nslots * 4);
  u32 *pinflow = (u32 *)arena_take(ctx, nslots * 4);
  if (!pstate || !pexit || !pdown || !pinflow) return HC_ERR_NOMEM;
  dim3 tg(tx, ty);
  unsigned sb = (unsigned)((nslots + 255) / 256);
  HC_PROF(ctx, st, "k_acc_tile_local");
  k_acc_tile<TABLE, false><<<tg, NTHREADS, 0, st>>>(dir, pstate, pexit, pdown, pinflow, acc, (int)ny, (int)nx, ctx->d_dbg);
  HC_LAUNCH_CHECK(ctx);
  HC_PROF(ctx, st, "k_acc_link_count");
  k_acc_link_count<<<sb, 256, 0, st>>>(pstate, pexit, pdown, nslots);
  HC_LAUNCH_CHECK(ctx);
  HC_PROF(ctx, st, "k_acc_link_mark");
  k_acc_link_mark<<<sb, 256, 0, st>>>(pstate, pdown, nslots);
  HC_LAUNCH_CHECK(ctx);
  HC_PROF(ctx, st, "k_acc_link_walk");
  k_acc_link_walk<<<sb, 256, 0, st>>>(pstate, pexit, pdown, pinflow, nslots);
  HC_LAUNCH_CHECK(ctx);
  HC_PROF(ctx, st, "k_acc_tile_final");
  k_acc_tile<TABLE, true><<<tg, NTHREADS, 0, st>>>(dir, pstate, pexit, pdown, pinflow, acc, (int)ny, (int)nx, ctx->d_dbg);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

int accum_run(hc_ctx *ctx, const uint8_t *dir, uint32_t *acc, int64_t ny, int64_t nx, int table,
              cudaStream_t st) {
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (table != HC_TABLE_WBT && table != HC_TABLE_TAUDEM) { hc_set_error("accum: bad table %d", table); return HC_ERR_ARG; }
  HC_TRY(arena_reset(ctx));
  if (table == HC_TABLE_WBT) return accum_impl<HC_TABLE_WBT>(ctx, dir, acc, ny, nx, st);
  return accum_impl<HC_TABLE_TAUDEM>(ctx, dir, acc, ny, nx, st);
}
