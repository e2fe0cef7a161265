// fill.cu — depression fill (epsilon = 0) as a watershed-graph minimax problem.
//
// Replaces the native work behind dep.fill_pits -> wbt.fill_depressions_wang_and_liu
// (tools/depressions.py:1236-1316, :1284) and TauDEM PitRemove (run_aws.py:582-588).
//
// The filled surface is out(c) = max(z(c), W(c)) with W(c) the minimax ("lowest spill") level
// over all 8-connected paths from c to a seed.  A serial priority-flood computes W with a heap;
// here it is computed in a constant number of bandwidth-bound grid passes plus O(log K) tiny
// table passes, K = number of local minima:
//
//   1. k_ptr      every valid cell points at its lowest 8-neighbour under the total order
//                 (z, index); local minima ("pits") get a compact basin id; seeds (and through
//                 them every cell that drains to a seed) belong to basin 0 = "outside".
//                 The 64x64 tile's pointer forest is compressed in shared memory before it is
//                 written, so a pointer leaves the kernel either resolved (tagged basin id) or
//                 aimed at a cell on another tile's perimeter.
//   2. k_perim    perimeter cells chase their chain across tiles to a tagged id.
//   3. (flatten)  every other cell takes its (now resolved) perimeter target's id: folded into the load of
//                 the first edge pass below (k_flatten only runs when there is no pit at all).
//   4. Boruvka rounds on the basin graph.  Edge weight between adjacent basins = the lowest pass
//                 min over adjacent cell pairs of max(z_a, z_b).  Each still-closed basin picks
//                 its cheapest outgoing edge straight from the grid (k_minedge, one 64-bit
//                 atomicMin per basin-boundary cell at most), the pointer forest of those picks
//                 is rooted and compressed on the K-entry tables, trees that reach basin 0 are
//                 resolved, the others contract into their root.  The second round also emits the compact
//                 list of (closed cell, foreign neighbour) pairs; later rounds run on that list.  Table
//                 work of a round is one cooperative launch (k_bor_persistent, grid.sync between phases).  A basin's final level is the
//                 max of the pick weights up its contraction chain (see DESIGN.md for the proof
//                 that this equals the minimax level).
//   5. k_apply    out = max(z, level[basin]).
//
// Only float comparisons/min/max touch elevations, so the result is bit-identical to the serial
// priority-flood (the eps=0 surface is unique).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"
#include "tma.cuh"

namespace cg = cooperative_groups;

#define T HC_TILE
#define TH (T + 2)
#define NTHREADS 256

// ---------------------------------------------------------------------------------------------
// 1. pointers + in-tile compression
// ---------------------------------------------------------------------------------------------
// TMA = true: the (tile + halo) window arrives by ONE bulk-tensor copy (tma.cuh; outside-raster elements NaN), the
// map's origin is the first readable row (rw.rlo: a band's halo row above); otherwise a row loader.
#define KP_WP 72   // window pitch = TMA box width.  A box must START on a 16-byte boundary of its row, so the box begins
#define KP_XO 3    // at column c0 - 4 and the logical window (column c0 - 1 first) sits KP_XO elements into it
template <bool TMA>
__global__ void __launch_bounds__(NTHREADS)
k_ptr(const float *__restrict__ z, const __grid_constant__ CUtensorMap mz, const uint8_t *__restrict__ ext,
      u32 *__restrict__ P, int ny, int nx, float nodata, u32 *__restrict__ pit_counter, hc_rows rw, u32 idbase,
      uint8_t *__restrict__ seam_seed) {
  __shared__ __align__(128) float sz_box[TH * KP_WP];
  float *sz = sz_box + KP_XO;   // logical window origin (row r0 - 1, column c0 - 1)
  __shared__ unsigned short snxt[T * T];
  __shared__ u32 stv[T * T];
  __shared__ u32 s_npit, s_base;
  __shared__ __align__(8) uint64_t mbar;

  const int r0 = blockIdx.y * T, c0 = blockIdx.x * T;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const float qnan = __int_as_float(0x7fc00000);
  if (tid == 0) s_npit = 0;

  if (TMA) {
    if (tid == 0) {
      hc_mbar_init(&mbar, 1);
      hc_mbar_expect_tx(&mbar, (u32)(TH * KP_WP * 4));
      hc_tma_load_2d(sz_box, &mz, &mbar, c0 - 1 - KP_XO, r0 - 1 - rw.rlo);
    }
    __syncthreads();
    hc_mbar_wait(&mbar, 0);
    // nodata -> NaN (outside-raster elements are NaN already)
    for (int i = tid; i < TH * KP_WP; i += NTHREADS) {
      const float v = sz_box[i];
      if (v == nodata) sz_box[i] = qnan;
    }
  } else {
    // tile + 1-cell halo -> shared; warp w takes rows w, w+8, ..., a lane columns lane, lane+32, lane+64
    for (int lr = warp; lr < TH; lr += NTHREADS / 32) {
      const int r = r0 + lr - 1;
      const bool rok = r >= rw.rlo && r < rw.rhi;
      const float *row = z + ((long long)r * nx + (c0 - 1));  // r may be -1 / ny: the band's halo rows
#pragma unroll
      for (int m = 0; m < 3; m++) {
        const int lc = lane + 32 * m;
        if (lc < TH) {
          const int c = c0 + lc - 1;
          float v = qnan;
          if (rok && c >= 0 && c < nx) {
            v = __ldg(row + lc);
            if (v == nodata) v = qnan;
          }
          sz[lr * KP_WP + lc] = v;
        }
      }
    }
  }
  __syncthreads();

  // pass A: classify, pick lowest neighbour, count pits.  Total order (z, row-major index): scanning
  // the 3x3 window in row-major order with a strict "<" keeps the first minimum, i.e. the smallest
  // index among equal elevations (the centre sits at position 4 of that scan).
  const int rb0 = rw.top ? -1 : 0, rb1 = rw.bot ? -1 : ny - 1;  // raster-border rows of this band (-1: none)
  for (int i = tid; i < T * T; i += NTHREADS) {
    int lr = i >> 6, lc = i & (T - 1);
    int r = r0 + lr, c = c0 + lc;
    unsigned short nxt = (unsigned short)i;
    u32 tv = HC_NONE;  // nodata / outside raster
    if (r < ny && c < nx) {
      const float *w = &sz[lr * KP_WP + lc];  // top-left corner of the 3x3 window
      const float v = w[KP_WP + 1];
      if (v == v) {
        // positions 3,2,1,0 (before the centre in index order) win ties against the centre and against
        // later ones: scan them backwards with "<="; positions 5..8 need a strictly lower elevation
        float bz = v;
        int bk = 4;
        bool anynan = false;
#define HC_SCAN(K, OFF, CMP)                   \
        {                                      \
          float zn = w[OFF];                   \
          anynan |= zn != zn;                  \
          if (zn CMP bz) { bz = zn; bk = K; }  \
        }
        HC_SCAN(3, KP_WP, <=) HC_SCAN(2, 2, <=) HC_SCAN(1, 1, <=) HC_SCAN(0, 0, <=)
        HC_SCAN(5, KP_WP + 2, <) HC_SCAN(6, 2 * KP_WP, <) HC_SCAN(7, 2 * KP_WP + 1, <) HC_SCAN(8, 2 * KP_WP + 2, <)
#undef HC_SCAN
        // seeds: raster border, or next to nodata (to EXTERIOR nodata only under the WBT rule).  A NaN
        // in the window that lies outside the raster can only belong to a border cell.
        bool seed = r == rb0 || r == rb1 || c == 0 || c == nx - 1;
        if (anynan && !seed) {
          if (ext == nullptr) seed = true;
          else {
            for (int dr = -1; dr <= 1; dr++)
              for (int dc = -1; dc <= 1; dc++) {
                float zn = w[(dr + 1) * KP_WP + dc + 1];
                if (zn != zn && ext[(long long)(r + dr) * nx + (c + dc)]) seed = true;  // (halo rows of a band: r + dr = -1 / ny)
              }
          }
        }
        // a band's seam rows (first / last own row next to another band) are not raster border
        const bool seam_t = rw.top && r == 0, seam_b = rw.bot && r == ny - 1;
        if (seam_t || seam_b) {
          // every seam cell is the root of its own "terminal" basin with a fixed id (Barnes 2016:
          // the tile perimeter); whether it is also a seed is remembered for the seam graph
          u32 sidx = (seam_t ? 0u : (u32)nx) + (u32)c;
          tv = HC_TAG | (1u + sidx);
          seam_seed[sidx] = seed ? 1 : 0;
        } else if (seed) {
          tv = HC_TAG;  // basin 0
        } else if (bk == 4) {
          tv = HC_TAG | 0x7FFFFFFEu;  // pit, id patched below
          atomicAdd(&s_npit, 1u);
        } else {
          const int bdr = bk / 3 - 1, bdc = bk - (bk / 3) * 3 - 1;
          int nlr = lr + bdr, nlc = lc + bdc;
          if (nlr >= 0 && nlr < T && nlc >= 0 && nlc < T) {
            nxt = (unsigned short)(nlr * T + nlc);
            tv = 0;  // not terminal
          } else {
            tv = (u32)((size_t)(r + bdr) * nx + (c + bdc));  // exits the tile: global index
          }
        }
      }
    }
    snxt[i] = nxt;
    stv[i] = tv;
  }
  __syncthreads();
  if (tid == 0) {
    s_base = s_npit ? atomicAdd(pit_counter, s_npit) : 0u;
    s_npit = 0;
  }
  __syncthreads();
  for (int i = tid; i < T * T; i += NTHREADS) {
    if (stv[i] == (HC_TAG | 0x7FFFFFFEu)) {
      u32 k = atomicAdd(&s_npit, 1u);
      stv[i] = HC_TAG | (idbase + s_base + k + 1u);
    }
  }
  __syncthreads();

  // pass B: pointer jumping inside the tile until every cell points at a terminal.  Terminals point at
  // themselves and nothing else does.  Four hops per round (paths shrink 4x between barriers); a cell whose new
  // target is its own target is resolved, drops out of the live mask and is never touched again.  Concurrent
  // writers only ever replace a pointer by one of its ancestors.
  u32 live = 0xFFFFu;
  for (;;) {
    int more = 0;
#pragma unroll
    for (int j = 0; j < T * T / NTHREADS; j++) {
      if (!(live & (1u << j))) continue;
      const int i = tid + j * NTHREADS;
      unsigned short p = snxt[i];
      p = snxt[p]; p = snxt[p]; p = snxt[p];
      const unsigned short t_ = snxt[p];
      snxt[i] = p;
      if (t_ == p) live &= ~(1u << j);
      else more = 1;
    }
    if (!__syncthreads_or(more)) break;
  }

  for (int i = tid; i < T * T; i += NTHREADS) {
    int lr = i >> 6, lc = i & (T - 1);
    int r = r0 + lr, c = c0 + lc;
    if (r < ny && c < nx) {
      u32 tv = stv[snxt[i]];
      if (tv == HC_NONE) tv = HC_TAG | HC_LAB_NODATA;
      P[(size_t)r * nx + c] = tv;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// 2./3. cross-tile resolution
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_perim(u32 *P, int ny, int nx) {
  // one block per tile, 4 sides x T threads; perimeter of the (possibly ragged) hr x hc tile
  const int r0 = blockIdx.y * T, c0 = blockIdx.x * T;
  const int hr = min(T, ny - r0), hc = min(T, nx - c0);
  const int side = threadIdx.x / T, j = threadIdx.x % T;
  int lr, lc;
  if (side == 0) { if (j >= hc) return; lr = 0; lc = j; }
  else if (side == 1) { if (j >= hc || hr < 2) return; lr = hr - 1; lc = j; }
  else if (side == 2) { if (j < 1 || j >= hr - 1) return; lr = j; lc = 0; }
  else { if (j < 1 || j >= hr - 1 || hc < 2) return; lr = j; lc = hc - 1; }
  size_t i = (size_t)(r0 + lr) * nx + (c0 + lc);
  volatile u32 *VP = P;
  u32 p = VP[i];
  if (p & HC_TAG) return;
  for (;;) {
    u32 q = VP[p];
    p = q;
    if (q & HC_TAG) break;
  }
  VP[i] = p;
}

__global__ void __launch_bounds__(256)
k_flatten(u32 *P, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    u32 p = P[i];
    if (!(p & HC_TAG)) {
      u32 q = __ldcg(&P[p]);
      // q is tagged: p is a perimeter cell already resolved by k_perim
      P[i] = q;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// 4. Boruvka rounds
// ---------------------------------------------------------------------------------------------
// EMIT: besides picking the cheapest edge, append every (closed cell, foreign neighbour) pair to a
// compact edge list; later rounds then run on that list instead of the grid.
// FIRST (round 0): the cross-tile flatten is folded into the load — a pointer that is not yet a tagged basin id
// aims at a perimeter cell k_perim has resolved, so one more hop gives the id, which is written back for the own
// cells (later kernels see a flat P) — and rep[] is the identity, so its look-up is skipped.
template <bool EMIT, bool FIRST>
__global__ void __launch_bounds__(NTHREADS)
k_minedge(const float *__restrict__ z, u32 *P, const u32 *__restrict__ rep,
          u64 *__restrict__ best, int ny, int nx, u32 *__restrict__ eA, u32 *__restrict__ eB,
          float *__restrict__ eW, u32 *__restrict__ ecount, u32 ecap, u32 nopen) {
  __shared__ float sz[TH * TH];
  __shared__ u32 sa[TH * TH];
  __shared__ u32 s_tot, s_base;
  const int r0 = blockIdx.y * T, c0 = blockIdx.x * T;
  const int tid = threadIdx.x;
  if (tid == 0) s_tot = 0;
#pragma unroll 4
  for (int i = tid; i < TH * TH; i += NTHREADS) {
    int lr = i / TH, lc = i - lr * TH;
    int r = r0 + lr - 1, c = c0 + lc - 1;
    bool inb = r >= 0 && r < ny && c >= 0 && c < nx;
    size_t g = inb ? (size_t)r * nx + c : 0;
    u32 p = __ldcg(&P[g]);
    float v = __ldg(&z[g]);
    if (FIRST && inb && !(p & HC_TAG)) {
      p = __ldcg(&P[p]);   // tagged: a perimeter cell (a stale read of a cell being flattened right now is still a valid pointer)
      if (lr >= 1 && lr <= T && lc >= 1 && lc <= T) P[g] = p;
    }
    u32 id = p & 0x7FFFFFFFu;
    u32 a = HC_NONE;
    if (inb && id != HC_LAB_NODATA) a = FIRST ? id : (id ? __ldg(&rep[id]) : 0u);
    sz[i] = v;
    sa[i] = a;
  }
  __syncthreads();
  u32 cnt = 0;
  for (int i = tid; i < T * T; i += NTHREADS) {
    int lr = i >> 6, lc = i & (T - 1);
    int ci = (lr + 1) * TH + lc + 1;
    u32 a = sa[ci];
    if (a < nopen || a == HC_NONE) continue;
    float v = sz[ci];
    u64 key = ~0ull;
#pragma unroll
    for (int dr = -1; dr <= 1; dr++)
#pragma unroll
      for (int dc = -1; dc <= 1; dc++) {
        if (dr == 0 && dc == 0) continue;
        int ni = ci + dr * TH + dc;
        u32 b = sa[ni];
        if (b == a || b == HC_NONE) continue;
        float w = fmaxf(v, sz[ni]);
        u64 k = ((u64)d_ford(w) << 32) | (u64)b;
        key = k < key ? k : key;
        cnt++;
      }
    if (key != ~0ull) {
      if (key < __ldcg(&best[a])) atomicMin(&best[a], key);
    }
  }
  if (!EMIT) return;
  u32 off = cnt ? atomicAdd(&s_tot, cnt) : 0u;
  __syncthreads();
  if (tid == 0) s_base = s_tot ? atomicAdd(ecount, s_tot) : 0u;
  __syncthreads();
  if (!cnt || (u64)s_base + s_tot > (u64)ecap) return;  // overflow: host falls back to grid rounds
  u32 o = s_base + off;
  for (int i = tid; i < T * T; i += NTHREADS) {
    int lr = i >> 6, lc = i & (T - 1);
    int ci = (lr + 1) * TH + lc + 1;
    u32 a = sa[ci];
    if (a < nopen || a == HC_NONE) continue;
    float v = sz[ci];
#pragma unroll
    for (int dr = -1; dr <= 1; dr++)
#pragma unroll
      for (int dc = -1; dc <= 1; dc++) {
        if (dr == 0 && dc == 0) continue;
        int ni = ci + dr * TH + dc;
        u32 b = sa[ni];
        if (b == a || b == HC_NONE) continue;
        eA[o] = a;
        eB[o] = b;
        eW[o] = fmaxf(v, sz[ni]);
        o++;
      }
  }
}


// ---------------------------------------------------------------------------------------------
// 4a. ONE grid pass that yields the whole basin graph (replaces the two k_minedge sweeps of round 0 / 1):
// every adjacent cell pair in different basins, at least one of them closed, is an edge (a, b, max(z_a, z_b));
// only the cheapest edge per basin pair matters for any Boruvka round, and a basin pair meets in few tiles, so
// the pairs are deduplicated per tile in a shared-memory hash table (key = the two basin ids, value = the
// ordered weight, atomicMin) and only the survivors are appended to the global list: ~100 entries per 64x64 tile
// on the C2 recipe instead of ~4700 differing cell pairs.  A thread walks 16 cells down one column and keeps the
// pair it saw last in registers (a boundary that runs along the walk produces the same pair again and again), so
// the table sees ~1 insert per boundary crossing.  Each unordered cell pair is looked at once, by its upper-left
// member (E, SW, S, SE neighbours; the member in the halo belongs to the neighbouring tile's sweep).  The load
// also folds the cross-tile flatten of P (see k_minedge<.., FIRST>).  Entries are UNDIRECTED: the list rounds let
// both ends pick.  A tile whose table overflows (noise: every cell its own basin) appends the pair unreduced.
// ---------------------------------------------------------------------------------------------
#define EH_CAP 2048
#define EH_PROBES 24
#define EDGES_SMEM (TH * TH * 8 + EH_CAP * 12 + 16)

__device__ __forceinline__ void eh_insert(u64 *keys, u32 *ws, u32 a, u32 b, u32 wf, u32 *__restrict__ eA,
                                          u32 *__restrict__ eB, float *__restrict__ eW, u32 *__restrict__ ecount, u32 ecap) {
  const u32 lo = a < b ? a : b, hi = a < b ? b : a;
  const u64 key = ((u64)lo << 32) | (u64)hi;
  u32 h = (u32)((lo * 0x9E3779B1u) ^ (hi * 0x85EBCA77u));
  h = (h ^ (h >> 15)) & (EH_CAP - 1);
  volatile u64 *vk = keys;
#pragma unroll 1
  for (int probe = 0; probe < EH_PROBES; probe++) {
    u64 cur = vk[h];
    if (cur == ~0ull) {
      cur = atomicCAS((unsigned long long *)&keys[h], ~0ull, (unsigned long long)key);
      if (cur == ~0ull) cur = key;
    }
    if (cur == key) { atomicMin(&ws[h], wf); return; }
    h = (h + 1) & (EH_CAP - 1);
  }
  u32 o = atomicAdd(ecount, 1u);   // table full around this slot: unreduced
  if (o < ecap) { eA[o] = lo; eB[o] = hi; eW[o] = d_funord(wf); }
}

// Fast path of k_edges: the distinct basin ids of the tile + halo get dense local ids (a small shared hash table
// keyed by the id, then a rank over its used slots); with L <= EL_MAX of them the cheapest weight per unordered
// pair lives in a triangular L x L table and a differing cell pair costs one pre-checked 32-bit shared atomicMin.
// Elevations are kept as order-preserving uints, so max(z_a, z_b) is one integer max.
// TMA = true: the two (tile + halo) windows of z and P arrive by two bulk-tensor copies issued by one thread
// (outside-raster elements: NaN / 0), see tma.cuh; otherwise a scalar loader.
#define EK_THREADS 512
#define WP 72                 // window pitch = TMA box width; the box starts at column c0 - 4 (16-byte aligned row start),
#define WXO 3                 // the logical window (column c0 - 1 first) sits WXO elements into it
#define WN (TH * WP)          // window elements (66 rows)
#define EL_SLOTS 512
#define EL_MAX 112
#define EL_NONE 0xFFFFu
#define EL_TAB (EL_MAX * (EL_MAX - 1) / 2 + 1)
// layout: sz u32[WN] | sl u16[WN] | gid u32[EL_MAX] | big: union { P staging u32[WN] (TMA), lkeys u32[EL_SLOTS] + ldense u16[EL_SLOTS], tab u32[EL_TAB], hash } | mbar, counters
#define EO_SL (WN * 4)
#define EO_GID (EO_SL + WN * 2)
#define EO_BIG ((EO_GID + EL_MAX * 4 + 127) & ~127)   // (a TMA destination: 128-byte aligned)
#define EO_LK (EO_BIG + WN * 4)          // the label hash lives behind the P staging window (both are live in step 1)
#define EO_LD (EO_LK + EL_SLOTS * 4)
#define EO_MAX3(a, b, c) ((a) > (b) ? ((a) > (c) ? (a) : (c)) : ((b) > (c) ? (b) : (c)))
#define EO_END (EO_BIG + ((EO_MAX3(WN * 4 + EL_SLOTS * 6, EL_TAB * 4, EH_CAP * 12) + 127) & ~127))
#undef EDGES_SMEM
#define EDGES_SMEM (EO_END + 128)

template <bool TMA>
__global__ void __launch_bounds__(EK_THREADS)
k_edges(const float *__restrict__ z, u32 *P, const __grid_constant__ CUtensorMap mz, const __grid_constant__ CUtensorMap mp,
        int ny, int nx, u32 *__restrict__ eA, u32 *__restrict__ eB, float *__restrict__ eW, u32 *__restrict__ ecount,
        u32 ecap, u32 nopen, unsigned long long *dbg) {
  extern __shared__ __align__(128) unsigned char smem_e[];
  u32 *sz_box = (u32 *)smem_e;                                 // float bits, then d_ford(z)
  u32 *sz = sz_box + WXO;                                      // logical window origin (row r0 - 1, column c0 - 1)
  unsigned short *sl = (unsigned short *)(smem_e + EO_SL);     // hash slot, later dense local id (EL_NONE: no cell)
  u32 *gid = (u32 *)(smem_e + EO_GID);
  u32 *stage_box = (u32 *)(smem_e + EO_BIG);                   // P window (TMA path only)
  u32 *stage = stage_box + WXO;
  u32 *tab = (u32 *)(smem_e + EO_BIG);                         // pair table: reuses the staging area after step 2
  u32 *lkeys = (u32 *)(smem_e + EO_LK);
  unsigned short *ldense = (unsigned short *)(smem_e + EO_LD);
  uint64_t *mbar = (uint64_t *)(smem_e + EO_END);
  u32 *s_cnt = (u32 *)(smem_e + EO_END + 16);   // [0] entries, [1] base in the global list, [2] label-table overflow, [3..18] warp sums
  const int r0 = blockIdx.y * T, c0 = blockIdx.x * T;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (TMA && tid == 0) {
    hc_mbar_init(mbar, 1);
    hc_mbar_expect_tx(mbar, 2u * WN * 4u);
    hc_tma_load_2d(sz_box, &mz, mbar, c0 - 1 - WXO, r0 - 1);
    hc_tma_load_2d(stage_box, &mp, mbar, c0 - 1 - WXO, r0 - 1);
  }
  if (tid < 3) s_cnt[tid] = 0;
  for (int i = tid; i < EL_SLOTS; i += EK_THREADS) lkeys[i] = HC_NONE;
  __syncthreads();
  if (TMA) hc_mbar_wait(mbar, 0);
  // 1. window -> ordered elevations + dense-id hash of the basin ids (cross-tile flatten of P folded in)
  for (int lr = warp; lr < TH; lr += EK_THREADS / 32) {
    const int r = r0 + lr - 1;
#pragma unroll
    for (int m = 0; m < 3; m++) {
      const int lc = lane + 32 * m;
      if (lc >= WP) break;
      const int c = c0 + lc - 1;
      const int i = lr * WP + lc;
      if (lc >= TH) { sl[i] = (unsigned short)EL_NONE; continue; }   // the window's pad columns
      const bool inb = r >= 0 && r < ny && c >= 0 && c < nx;
      const size_t g = inb ? (size_t)r * nx + c : 0;
      u32 p;
      float v;
      if (TMA) { p = stage[i]; v = __uint_as_float(sz[i]); }
      else { p = __ldcg(&P[g]); v = __ldg(&z[g]); }
      if (inb && !(p & HC_TAG)) {
        p = __ldcg(&P[p]);   // tagged: a perimeter cell k_perim resolved (a stale read of a cell being flattened right now is still a valid pointer)
        if (lr >= 1 && lr <= T && lc >= 1 && lc <= T) P[g] = p;
      }
      const u32 id = p & 0x7FFFFFFFu;
      u32 slot = EL_NONE;
      if (inb && id != HC_LAB_NODATA) {
        u32 h = (id * 0x9E3779B1u) >> (32 - 9);
        for (int probe = 0; probe < EL_SLOTS; probe++) {
          u32 cur = ((volatile u32 *)lkeys)[h];
          if (cur == HC_NONE) {
            cur = atomicCAS(&lkeys[h], HC_NONE, id);
            if (cur == HC_NONE) cur = id;
          }
          if (cur == id) { slot = h; break; }
          h = (h + 1) & (EL_SLOTS - 1);
        }
        if (slot == EL_NONE) s_cnt[2] = 1;   // more than EL_SLOTS distinct ids
      }
      sz[i] = d_ford(v);   // (NaN of a no-cell element maps to some uint: never used, its sl is EL_NONE)
      sl[i] = (unsigned short)slot;
    }
  }
  __syncthreads();
  // 2. rank the used slots -> dense ids; gid[dense] = basin id
  {
    const u32 k0 = lkeys[tid];
    const u32 v = k0 != HC_NONE;
    u32 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u32 t_ = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t_; }
    if (lane == 31) s_cnt[3 + warp] = inc;
    __syncthreads();
    u32 base = 0;
    for (int w = 0; w < warp; w++) base += s_cnt[3 + w];
    const u32 excl = base + inc - v;
    if (v) { ldense[tid] = (unsigned short)excl; if (excl < EL_MAX) gid[excl] = k0; }
    if (tid == EK_THREADS - 1) s_cnt[0] = base + inc;   // L
  }
  __syncthreads();
  const u32 L = s_cnt[0];
  const bool fast = L <= EL_MAX && !s_cnt[2];
  if (fast) {
    for (int i = tid; i < WN; i += EK_THREADS) { const u32 s_ = sl[i]; if (s_ != EL_NONE) sl[i] = ldense[s_]; }
  }
  __syncthreads();
  if (tid == 0) s_cnt[0] = 0;
  if (fast) {
    for (int i = tid; i < EL_TAB; i += EK_THREADS) tab[i] = 0xFFFFFFFFu;
    __syncthreads();
    // 3. column walk: thread -> column lc = tid % 64, rows [8 * (tid / 64), +8); E, SW, S, SE pairs
    const int lc = tid & (T - 1), lr0 = (tid >> 6) * (T / 8);
    int ci = (lr0 + 1) * WP + lc + 1;
    u32 a_m = sl[ci], a_r = sl[ci + 1];
#pragma unroll
    for (int k = 0; k < T / 8; k++, ci += WP) {
      const u32 b_l = sl[ci + WP - 1], b_m = sl[ci + WP], b_r = sl[ci + WP + 1];
      if (a_m != EL_NONE) {
        const u32 z_m = sz[ci];
#define HC_PAIR(B, OFF)                                                     \
        if ((B) != a_m && (B) != EL_NONE) {                                 \
          const u32 wf = max(z_m, sz[ci + (OFF)]);                          \
          const u32 lo = min(a_m, (B)), hi = max(a_m, (B));                 \
          u32 *slot_ = &tab[((hi * (hi - 1)) >> 1) + lo];                   \
          if (wf < *(volatile u32 *)slot_) atomicMin(slot_, wf);            \
        }
        HC_PAIR(a_r, 1) HC_PAIR(b_l, WP - 1) HC_PAIR(b_m, WP) HC_PAIR(b_r, WP + 1)
#undef HC_PAIR
      }
      a_m = b_m; a_r = b_r;
    }
    __syncthreads();
    // 4. emit the pairs that met in this tile (at least one end closed)
    u32 mine = 0;
    for (int i = tid; i < EL_TAB; i += EK_THREADS) mine += tab[i] != 0xFFFFFFFFu;
    u32 off = mine ? atomicAdd(&s_cnt[0], mine) : 0u;
    __syncthreads();
    if (tid == 0) s_cnt[1] = s_cnt[0] ? atomicAdd(ecount, s_cnt[0]) : 0u;
    __syncthreads();
    if (!mine || (u64)s_cnt[1] + s_cnt[0] > (u64)ecap) return;   // overflow: the host falls back to grid rounds
    u32 o = s_cnt[1] + off;
    for (int i = tid; i < EL_TAB; i += EK_THREADS) {
      const u32 w = tab[i];
      if (w == 0xFFFFFFFFu) continue;
      // i = hi * (hi - 1) / 2 + lo  ->  hi = floor((1 + sqrt(1 + 8 i)) / 2)
      u32 hi = (u32)((1.0f + sqrtf(1.0f + 8.0f * (float)i)) * 0.5f);
      while (hi * (hi - 1) / 2 > (u32)i) hi--;
      while ((hi + 1) * hi / 2 <= (u32)i) hi++;
      const u32 lo = (u32)i - hi * (hi - 1) / 2;
      const u32 ga = gid[lo], gb = gid[hi];
      const bool keep = ga >= nopen || gb >= nopen;
      eA[o] = keep ? ga : 0u;   // (an open-open pair keeps its slot as a harmless 0-0 entry: the count was taken before)
      eB[o] = keep ? gb : 0u;
      eW[o] = d_funord(w);
      o++;
    }
    return;
  }
  // ---- general path (too many basins meet in this tile): 64-bit-key hash, a thread keeps its last pair ----
  if (dbg && tid == 0) HC_DBG(8, 1);
  u64 *keys = (u64 *)(smem_e + EO_BIG);
  u32 *ws = (u32 *)(smem_e + EO_BIG + EH_CAP * 8);
  for (int i = tid; i < EH_CAP; i += EK_THREADS) { keys[i] = ~0ull; ws[i] = 0xFFFFFFFFu; }
  __syncthreads();
  {
    // (no room for a u32 id per cell beside the hash: ids are read through P on the fly; the own cells are flat by
    // now, a halo cell may still need the hop)
    const int lc = tid & (T - 1), lr0 = (tid >> 6) * (T / 8);
    u32 ka = HC_NONE, kb = HC_NONE, kw = 0xFFFFFFFFu;
    auto lab = [&](int lr, int lc2) -> u32 {   // basin id of window cell (lr, lc2), HC_NONE if no cell
      int r = r0 + lr - 1, c = c0 + lc2 - 1;
      if (r < 0 || r >= ny || c < 0 || c >= nx) return HC_NONE;
      u32 p = __ldcg(&P[(size_t)r * nx + c]);
      if (!(p & HC_TAG)) p = __ldcg(&P[p]);
      p &= 0x7FFFFFFFu;
      return p == HC_LAB_NODATA ? HC_NONE : p;
    };
    for (int k = 0; k < T / 8; k++) {
      const int lr = lr0 + 1 + k, lcc = lc + 1;
      const int ci = lr * WP + lcc;
      const u32 a_m = lab(lr, lcc);
      if (a_m == HC_NONE) continue;
      const u32 z_m = sz[ci];
      const int offs[4][2] = {{0, 1}, {1, -1}, {1, 0}, {1, 1}};
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const u32 b = lab(lr + offs[q][0], lcc + offs[q][1]);
        if (b == a_m || b == HC_NONE || (a_m < nopen && b < nopen)) continue;
        const u32 wf = max(z_m, sz[ci + offs[q][0] * WP + offs[q][1]]);
        const u32 lo = a_m < b ? a_m : b, hi = a_m < b ? b : a_m;
        if (lo == ka && hi == kb) kw = wf < kw ? wf : kw;
        else {
          if (ka != HC_NONE) eh_insert(keys, ws, ka, kb, kw, eA, eB, eW, ecount, ecap);
          ka = lo; kb = hi; kw = wf;
        }
      }
    }
    if (ka != HC_NONE) eh_insert(keys, ws, ka, kb, kw, eA, eB, eW, ecount, ecap);
  }
  __syncthreads();
  u32 mine = 0;
  for (int i = tid; i < EH_CAP; i += EK_THREADS) mine += keys[i] != ~0ull;
  u32 off = mine ? atomicAdd(&s_cnt[0], mine) : 0u;
  __syncthreads();
  if (tid == 0) s_cnt[1] = s_cnt[0] ? atomicAdd(ecount, s_cnt[0]) : 0u;
  __syncthreads();
  if (!mine || (u64)s_cnt[1] + s_cnt[0] > (u64)ecap) return;
  u32 o = s_cnt[1] + off;
  for (int i = tid; i < EH_CAP; i += EK_THREADS) {
    const u64 k = keys[i];
    if (k == ~0ull) continue;
    eA[o] = (u32)(k >> 32);
    eB[o] = (u32)(k & 0xFFFFFFFFu);
    eW[o] = d_funord(ws[i]);
    o++;
  }
}

// a Boruvka pick round on the compact edge list
__global__ void __launch_bounds__(256)
k_minedge_list(const u32 *__restrict__ eA, const u32 *__restrict__ eB, const float *__restrict__ eW, u32 ne,
               const u32 *__restrict__ rep, u64 *__restrict__ best, u32 nopen) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ne) return;
  u32 a = __ldg(&rep[eA[i]]);
  if (a < nopen) return;
  u32 b = __ldg(&rep[eB[i]]);
  if (a == b) return;
  u64 key = ((u64)d_ford(eW[i]) << 32) | (u64)b;
  if (key < __ldcg(&best[a])) atomicMin(&best[a], key);
}

__global__ void k_tab_init(u32 *rep, float *lvl, u32 *hp, u32 K1) {
  u32 b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= K1) return;
  rep[b] = b;
  lvl[b] = -INFINITY;
  hp[b] = 0;
}

// each active basin adopts its cheapest edge: ptr = target, lvl = max(lvl, weight)
__global__ void k_hook(const u32 *rep, const u64 *best, u32 *ptr, float *lvl, u32 K1, u32 nopen) {
  u32 b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= K1) return;
  if (b < nopen || rep[b] != b) { ptr[b] = b < nopen ? b : rep[b]; return; }
  u64 k = best[b];
  if (k == ~0ull) { ptr[b] = 0u; return; }  // no neighbour at all (walled in): never raised
  float w = d_funord((u32)(k >> 32));
  ptr[b] = (u32)(k & 0xFFFFFFFFu);
  lvl[b] = fmaxf(lvl[b], w);
}

// break mutual pairs: the smaller id becomes the root of the pair's tree
__global__ void k_mutual(const u32 *rep, const u32 *ptr, u32 *ptr2, u32 K1, u32 nopen) {
  u32 b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= K1) return;
  u32 p = ptr[b];
  if (b >= nopen && rep[b] == b && p >= nopen && ptr[p] == b && b < p) p = b;
  ptr2[b] = p;
}

__global__ void k_jump(u32 *ptr, u32 K1, u32 *changed) {
  u32 b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= K1) return;
  u32 p = ptr[b];
  u32 pp = ptr[p];
  if (pp != p) { ptr[b] = pp; *changed = 1u; }
}

// merge trees into their roots, count what is still closed
// n_active[0] = closed roots left, n_active[1] = roots merged this round
__global__ void k_merge(u32 *rep, u32 *hp, const u32 *ptr, u32 K1, u32 *n_active, u32 nopen) {
  u32 b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= K1 || b < nopen) return;
  if (rep[b] != b) return;
  u32 r = ptr[b];
  if (r != b) { hp[b] = r; rep[b] = r; atomicAdd(n_active + 1, 1u); }
  else atomicAdd(n_active, 1u);
}

__global__ void k_repflat(u32 *rep, u32 K1) {
  u32 b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= K1) return;
  u32 s = rep[b];
  u32 r = rep[s];
  if (r != s) rep[b] = r;
}

// final level of an original basin = max of pick weights up its contraction chain
// (troot = the open node the chain ends in: 0 = outside, or a band's seam terminal)
__global__ void k_levels(const u32 *hp, const float *lvl, float *level0, u32 *troot, u32 K1, u32 nopen) {
  u32 b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= K1) return;
  float m = -INFINITY;
  u32 x = b;
  while (x >= nopen) { m = fmaxf(m, lvl[x]); x = hp[x]; }
  level0[b] = m;
  if (troot) troot[b] = x;
}

// ---------------------------------------------------------------------------------------------
// 5. apply
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_apply(const float *__restrict__ z, const u32 *__restrict__ P, const float *__restrict__ level0,
        float *__restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float v = z[i];
    u32 id = P[i] & 0x7FFFFFFFu;
    if (id != 0u && id != HC_LAB_NODATA) v = fmaxf(v, __ldg(&level0[id]));
    out[i] = v;
  }
}

// ---------------------------------------------------------------------------------------------
size_t fill_workspace(int64_t ny, int64_t nx) {
  size_t n = (size_t)ny * nx;
  size_t tiles = (size_t)((ny + T - 1) / T) * ((nx + T - 1) / T);
  return hc_align(n * 4) + hc_align(n) + 2 * hc_align(tiles) + hc_align(256);
}

// exterior-nodata mask for the WBT seed rule; defined in label.cu
int ext_nodata_mask(hc_ctx *ctx, const float *z, uint8_t *ext, int64_t ny, int64_t nx, float nodata,
                    int *has_interior, cudaStream_t st);

// ---------------------------------------------------------------------------------------------
// generic Boruvka rounds on an explicit undirected edge list (both endpoints see every entry; ties
// broken by (weight, entry index), which is the same key from both ends => only 2-cycles).
// Nodes < nopen are "open" (never hook).  msf = 1: a closed component without outgoing edges
// stays a root (minimum spanning FOREST); msf = 0: it is attached to node 0 unraised.
// mark (optional) receives 1 for every entry picked: the MST/MSF edges.
// ---------------------------------------------------------------------------------------------
struct bor_tabs { u64 *best; u32 *rep, *hp, *ptr, *ptr2; float *lvl; };

static int bor_tabs_alloc(hc_ctx *ctx, bor_tabs *t, u32 K1) {
  t->best = (u64 *)arena_take(ctx, (size_t)K1 * 8);
  t->rep = (u32 *)arena_take(ctx, (size_t)K1 * 4);
  t->hp = (u32 *)arena_take(ctx, (size_t)K1 * 4);
  t->ptr = (u32 *)arena_take(ctx, (size_t)K1 * 4);
  t->ptr2 = (u32 *)arena_take(ctx, (size_t)K1 * 4);
  t->lvl = (float *)arena_take(ctx, (size_t)K1 * 4);
  if (!t->best || !t->rep || !t->hp || !t->ptr || !t->ptr2 || !t->lvl) return HC_ERR_NOMEM;
  return HC_OK;
}

// hook -> break mutual pairs -> pointer-double to the roots -> merge; cnt: [1]=changed [2]=active [3]=merged
static int bor_contract(hc_ctx *ctx, const bor_tabs &t, u32 K1, u32 nopen, u32 *cnt, u32 *active, u32 *merged,
                        cudaStream_t st) {
  const int tb = 256;
  const int tgK = (int)((K1 + tb - 1) / tb);
  HC_PROF(ctx, st, "k_mutual");
  k_mutual<<<tgK, tb, 0, st>>>(t.rep, t.ptr2, t.ptr, K1, nopen);
  HC_LAUNCH_CHECK(ctx);
  // pointer doubling to the roots: 8 doublings (paths up to 256 hops) per read-back
  for (int it = 0; it < 16; it++) {
    for (int j = 0; j < 7; j++) {
      HC_PROF(ctx, st, "k_jump");
      k_jump<<<tgK, tb, 0, st>>>(t.ptr, K1, cnt + 5);
      HC_LAUNCH_CHECK(ctx);
    }
    HC_CUDA(cudaMemsetAsync(cnt + 1, 0, 4, st));
    HC_PROF(ctx, st, "k_jump");
    k_jump<<<tgK, tb, 0, st>>>(t.ptr, K1, cnt + 1);
    HC_LAUNCH_CHECK(ctx);
    HC_CUDA(cudaMemcpyAsync(ctx->h_mail, cnt + 1, 4, cudaMemcpyDeviceToHost, st));
    HC_CUDA(cudaStreamSynchronize(st));
    if (!ctx->h_mail[0]) break;
  }
  HC_CUDA(cudaMemsetAsync(cnt + 2, 0, 8, st));
  HC_PROF(ctx, st, "k_merge");
  k_merge<<<tgK, tb, 0, st>>>(t.rep, t.hp, t.ptr, K1, cnt + 2, nopen);
  HC_LAUNCH_CHECK(ctx);
  HC_PROF(ctx, st, "k_repflat");
  k_repflat<<<tgK, tb, 0, st>>>(t.rep, K1);
  HC_LAUNCH_CHECK(ctx);
  HC_CUDA(cudaMemcpyAsync(ctx->h_mail, cnt + 2, 8, cudaMemcpyDeviceToHost, st));
  HC_CUDA(cudaStreamSynchronize(st));
  *active = ctx->h_mail[0];
  *merged = ctx->h_mail[1];
  return HC_OK;
}

// The remaining Boruvka rounds in ONE cooperative launch: every phase is a grid-stride loop, phases are
// separated by grid.sync(), loop decisions are read from counters every block sees after a sync.  Replaces
// ~10 tiny launches and 2-3 host read-backs per round (the rounds are latency, not throughput: the tables
// have 1e4..1e6 entries).  idx_keys = 1: undirected list, key (weight, entry index), both ends pick, picked
// entries are marked (MSF / seam graph); idx_keys = 0: the fill's directed list, key (weight, neighbour).
// ctl: [0..2] jump-changed flags (rotating: a flag is cleared two syncs after its last reader),
//      [3..6] active/merged counters (parity), [7] rounds done
struct bor_pargs {
  u64 *best; u32 *rep, *hp, *ptr, *ptr2; float *lvl;
  const u32 *eA, *eB; const float *eW; uint8_t *mark; u32 *ctl;
  u32 K1, nopen, ne; int msf, idx_keys, do_init, max_rounds, skip_pick, both;
};
__global__ void __launch_bounds__(512)
k_bor_persistent(bor_pargs a) {
  cg::grid_group grid = cg::this_grid();
  const u32 tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  if (a.do_init) {
    for (u32 b = tid; b < a.K1; b += nth) { a.rep[b] = b; a.lvl[b] = -INFINITY; a.hp[b] = 0; }
    if (a.mark) for (u32 i = tid; i < a.ne; i += nth) a.mark[i] = 0;
  }
  if (tid < 9) a.ctl[tid] = 0;
  grid.sync();
  int round = 0;
  // phase time stamps of thread 0 (ns), ctl[32 + 2 i]: diagnostic, read with HC_TRACE
  int ts_n = 0;
#define HC_TS() do { if (tid == 0 && ts_n < 96) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); ((unsigned long long *)(a.ctl + 32))[ts_n++] = t_; } } while (0)
  HC_TS();
  for (; round < a.max_rounds; round++) {
    // (skip_pick: best[] was filled by a grid pass over the raster; this launch only contracts)
    if (!a.skip_pick) {
    for (u32 b = tid; b < a.K1; b += nth) a.best[b] = ~0ull;
    grid.sync();
    // pick
    for (u32 i = tid; i < a.ne; i += nth) {
      u32 ra = __ldcg(&a.rep[a.eA[i]]);
      u32 rb = __ldcg(&a.rep[a.eB[i]]);
      if (ra == rb) continue;
      if (a.idx_keys) {
        u64 key = ((u64)d_ford(a.eW[i]) << 32) | (u64)i;
        if (ra >= a.nopen && key < __ldcg(&a.best[ra])) atomicMin(&a.best[ra], key);
        if (rb >= a.nopen && key < __ldcg(&a.best[rb])) atomicMin(&a.best[rb], key);
      } else {
        // key (weight, neighbour): the fill's lists; both = 1: undirected entries, either end may pick
        const u64 wk = (u64)d_ford(a.eW[i]) << 32;
        if (ra >= a.nopen) {
          u64 key = wk | (u64)rb;
          if (key < __ldcg(&a.best[ra])) atomicMin(&a.best[ra], key);
        }
        if (a.both && rb >= a.nopen) {
          u64 key = wk | (u64)ra;
          if (key < __ldcg(&a.best[rb])) atomicMin(&a.best[rb], key);
        }
      }
    }
    grid.sync();
    HC_TS();
    }
    // hook -> ptr2
    for (u32 b = tid; b < a.K1; b += nth) {
      u32 r = __ldcg(&a.rep[b]);
      if (b < a.nopen || r != b) { a.ptr2[b] = b < a.nopen ? b : r; continue; }
      u64 k = __ldcg(&a.best[b]);
      if (k == ~0ull) { a.ptr2[b] = a.msf ? b : 0u; continue; }
      if (a.idx_keys) {
        u32 idx = (u32)(k & 0xFFFFFFFFu);
        u32 ra = __ldcg(&a.rep[a.eA[idx]]), rb = __ldcg(&a.rep[a.eB[idx]]);
        a.ptr2[b] = ra == b ? rb : ra;
        a.lvl[b] = fmaxf(a.lvl[b], a.eW[idx]);
        if (a.mark) a.mark[idx] = 1;
      } else {
        a.ptr2[b] = (u32)(k & 0xFFFFFFFFu);
        a.lvl[b] = fmaxf(a.lvl[b], d_funord((u32)(k >> 32)));
      }
    }
    grid.sync();
    HC_TS();
    // break mutual pairs -> ptr
    for (u32 b = tid; b < a.K1; b += nth) {
      u32 p = __ldcg(&a.ptr2[b]);
      if (b >= a.nopen && __ldcg(&a.rep[b]) == b && p >= a.nopen && __ldcg(&a.ptr2[p]) == b && b < p) p = b;
      a.ptr[b] = p;
    }
    grid.sync();
    HC_TS();
    // pointer doubling to the roots
    for (int it = 0;; it++) {
      u32 *flag = &a.ctl[it % 3];
      if (tid == 0) a.ctl[(it + 1) % 3] = 0;   // last read in iteration it-2: every thread has passed a sync since
      int ch = 0;
      for (u32 b = tid; b < a.K1; b += nth) {
        u32 p = __ldcg(&a.ptr[b]);
        u32 pp = __ldcg(&a.ptr[p]);
        if (pp != p) {
          // two more hops per sweep (paths shrink 4x per grid barrier instead of 2x)
          u32 p3 = __ldcg(&a.ptr[pp]);
          p3 = __ldcg(&a.ptr[p3]);
          a.ptr[b] = p3; ch = 1;
        }
      }
      // one flag store per block: hundreds of thousands of stores to one address serialise in its L2 slice
      ch = __syncthreads_or(ch);
      if (ch && threadIdx.x == 0) *flag = 1u;
      grid.sync();
      if (!__ldcg(flag)) break;
    }
    HC_TS();
    // merge trees into their roots
    u32 *ctr = &a.ctl[3 + 2 * (round & 1)];
    if (tid == 0) { a.ctl[3 + 2 * ((round + 1) & 1)] = 0; a.ctl[4 + 2 * ((round + 1) & 1)] = 0; }
    u32 nact = 0, nmer = 0;
    for (u32 b = tid; b < a.K1; b += nth) {
      if (b < a.nopen || __ldcg(&a.rep[b]) != b) continue;
      u32 r = __ldcg(&a.ptr[b]);
      if (r != b) { a.hp[b] = r; a.rep[b] = r; nmer++; }
      else nact++;
    }
    // (block totals first: one atomic per block, not one per thread, on the two shared counters)
    nact = __reduce_add_sync(0xFFFFFFFFu, nact);
    nmer = __reduce_add_sync(0xFFFFFFFFu, nmer);
    __shared__ u32 s_act, s_mer;
    if (threadIdx.x == 0) { s_act = 0; s_mer = 0; }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { if (nact) atomicAdd(&s_act, nact); if (nmer) atomicAdd(&s_mer, nmer); }
    __syncthreads();
    if (threadIdx.x == 0) { if (s_act) atomicAdd(ctr, s_act); if (s_mer) atomicAdd(ctr + 1, s_mer); }
    grid.sync();
    for (u32 b = tid; b < a.K1; b += nth) {
      u32 s0 = __ldcg(&a.rep[b]);
      u32 r = __ldcg(&a.rep[s0]);
      if (r != s0) a.rep[b] = r;
    }
    grid.sync();
    HC_TS();
    const u32 active = __ldcg(ctr), merged = __ldcg(ctr + 1);
    if (active == 0 || merged == 0) { round++; break; }
  }
  if (tid == 0) { a.ctl[7] = (u32)round; a.ctl[8] = a.ctl[3 + 2 * ((round - 1) & 1)]; a.ctl[9] = (u32)ts_n; }   // rounds done, closed roots left
}

// launch helper: as many co-resident 512-thread blocks as the work needs (at most the device holds)
static int bor_persistent_launch(hc_ctx *ctx, bor_pargs &a, int *rounds, cudaStream_t st) {
  static int max_blocks = 0;
  if (!max_blocks) {
    int per_sm = 0;
    HC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bor_persistent, 512, 0));
    max_blocks = per_sm * ctx->sm_count;
    if (max_blocks < 1) { hc_set_error("cooperative launch: no resident block"); return HC_ERR_CUDA; }
  }
  size_t work = a.ne > a.K1 ? a.ne : a.K1;
  int blocks = (int)((work + 2047) / 2048);
  if (blocks < 1) blocks = 1;
  if (blocks > max_blocks) blocks = max_blocks;
  void *args[] = {&a};
  HC_PROF(ctx, st, "k_bor_persistent");
  HC_CUDA(cudaLaunchCooperativeKernel((void *)k_bor_persistent, dim3(blocks), dim3(512), args, 0, st));
  HC_LAUNCH_CHECK(ctx);
  if (rounds) {
    HC_CUDA(cudaMemcpyAsync(ctx->h_mail, a.ctl + 7, 4, cudaMemcpyDeviceToHost, st));
    HC_CUDA(cudaStreamSynchronize(st));
    *rounds = (int)ctx->h_mail[0];
  }
  return HC_OK;
}

static int bor_list_solve(hc_ctx *ctx, const bor_tabs &t, u32 K1, u32 nopen, const u32 *eA, const u32 *eB,
                          const float *eW, u32 ne, int msf, uint8_t *mark, u32 *cnt, int *rounds, cudaStream_t st) {
  bor_pargs a;
  a.best = t.best; a.rep = t.rep; a.hp = t.hp; a.ptr = t.ptr; a.ptr2 = t.ptr2; a.lvl = t.lvl;
  a.eA = eA; a.eB = eB; a.eW = eW; a.mark = mark; a.ctl = cnt + 16;
  a.K1 = K1; a.nopen = nopen; a.ne = ne; a.msf = msf; a.idx_keys = 1; a.do_init = 1; a.max_rounds = 64; a.skip_pick = 0; a.both = 0;
  if (!getenv("HC_BAND_STATS")) rounds = nullptr;   // (the diagnostic read-back costs a host sync per solve)
  return bor_persistent_launch(ctx, a, rounds, st);
}

// ---------------------------------------------------------------------------------------------
// phase 1 (shared by the whole-raster fill and the band-local fill): watershed labels P, local
// levels level0 (relative to the open nodes), troot (which open node each basin hangs under)
// ---------------------------------------------------------------------------------------------
struct fill_state {
  u32 *P; float *level0; u32 *troot; u32 K1, nopen; u32 *cnt; uint8_t *seam_seed;
  u32 *eA, *eB; float *eW; u32 ecap;
};

static int fill_local(hc_ctx *ctx, const float *z, int64_t ny, int64_t nx, float nodata, int seed_mode, hc_rows rw,
                      fill_state *fs, cudaStream_t st, const uint8_t *ext_in = nullptr) {
  size_t n = (size_t)ny * nx;
  if (n >= 0x7FFFFFF0ull) { hc_set_error("raster too large for 31-bit cell indices"); return HC_ERR_TOO_LARGE; }
  const bool banded = rw.top || rw.bot;
  if (banded && ny < 2) { hc_set_error("a band needs at least 2 rows"); return HC_ERR_ARG; }
  // (a band cannot tell exterior from interior nodata by itself: under the WBT rule the caller passes the
  //  exterior-nodata raster it got from the banded labelling, halo rows included; NULL = all nodata is exterior)
  const int tx = (int)((nx + T - 1) / T), ty = (int)((ny + T - 1) / T);
  const size_t tiles = (size_t)tx * ty;
  HC_TRY(arena_reset(ctx));

  const uint8_t *ext = nullptr;
  if (seed_mode == HC_SEED_WBT && banded) {
    ext = ext_in;
  } else if (seed_mode == HC_SEED_WBT) {
    uint8_t *e = (uint8_t *)arena_take(ctx, n);
    if (!e) return HC_ERR_NOMEM;
    int has_interior = 0;
    HC_TRY(ext_nodata_mask(ctx, z, e, ny, nx, nodata, &has_interior, st));
    ext = has_interior ? e : nullptr;  // every nodata cell sits on the raster border: all exterior
  }
  u32 *P = (u32 *)arena_take(ctx, n * 4);
  u32 *cnt = (u32 *)arena_take(ctx, 2048);  // [0]=pit counter [1]=changed [2]=n_active [3]=n_merged [4]=edges
  uint8_t *seam_seed = (uint8_t *)arena_take(ctx, banded ? (size_t)2 * nx : 1);
  if (!P || !cnt || !seam_seed) return HC_ERR_NOMEM;
  (void)tiles;
  const u32 nopen = banded ? (u32)(2 * nx + 1) : 1u;

  HC_CUDA(cudaMemsetAsync(cnt, 0, 2048, st));
  if (banded) HC_CUDA(cudaMemsetAsync(seam_seed, 0, (size_t)2 * nx, st));
  dim3 tg(tx, ty);
  {
    // window map over every readable row (a band's halo rows included): box 68 x 66, origin row rw.rlo
    hc_tmap mzp;
    hc_tmap_2d(&mzp, z + (long long)rw.rlo * nx, 4, 1, (uint64_t)nx, (uint64_t)(rw.rhi - rw.rlo), (uint64_t)nx * 4, KP_WP, TH);
    HC_PROF(ctx, st, "k_ptr");
    if (mzp.ok) k_ptr<true><<<tg, NTHREADS, 0, st>>>(z, mzp.map, ext, P, (int)ny, (int)nx, nodata, cnt, rw, nopen - 1, seam_seed);
    else k_ptr<false><<<tg, NTHREADS, 0, st>>>(z, mzp.map, ext, P, (int)ny, (int)nx, nodata, cnt, rw, nopen - 1, seam_seed);
    HC_LAUNCH_CHECK(ctx);
  }
  HC_PROF(ctx, st, "k_perim");
  k_perim<<<tg, 256, 0, st>>>(P, (int)ny, (int)nx);
  HC_LAUNCH_CHECK(ctx);
  int gblocks = ctx->sm_count * 8;

  HC_CUDA(cudaMemcpyAsync(ctx->h_mail, cnt, 4, cudaMemcpyDeviceToHost, st));
  HC_CUDA(cudaStreamSynchronize(st));
  const u32 K = ctx->h_mail[0];        // pits
  const u32 K1 = nopen + K;            // open nodes + pits
  ctx->fill_basins = K;
  ctx->fill_rounds = 0;

  bor_tabs t;
  HC_TRY(bor_tabs_alloc(ctx, &t, K1));
  float *level0 = (float *)t.ptr2;  // reused after the rounds
  u32 *troot = t.ptr;
  // compact edge list for the rounds after the second grid pass
  // (the emit round lists ~0.2 entries per cell on the C2 recipe; if it overflows, the rounds fall back to grid
  // passes, which is correct but costs a full sweep per round: 6 B/cell of headroom is cheap on a 180 GB part)
  size_t ecap_sz = n / 2 > (size_t)(1u << 20) ? n / 2 : (size_t)(1u << 20);
  if (ecap_sz > (size_t)(1u << 30)) ecap_sz = (size_t)(1u << 30);
  const u32 ecap = (u32)ecap_sz;
  u32 *eA = (u32 *)arena_take(ctx, (size_t)ecap * 4);
  u32 *eB = (u32 *)arena_take(ctx, (size_t)ecap * 4);
  float *eW = (float *)arena_take(ctx, (size_t)ecap * 4);
  if (!eA || !eB || !eW) return HC_ERR_NOMEM;
  u32 ne = 0;
  bool have_list = false;

  const int tb = 256;
  const int tgK = (int)((K1 + tb - 1) / tb);
  HC_PROF(ctx, st, "k_tab_init");
  k_tab_init<<<tgK, tb, 0, st>>>(t.rep, t.lvl, t.hp, K1);
  HC_LAUNCH_CHECK(ctx);

  u32 active = K, merged = 0;
  int round = 0;
  if (K == 0) {   // no pit at all: no round will fold the flatten into its load
    HC_PROF(ctx, st, "k_flatten");
    k_flatten<<<gblocks, 256, 0, st>>>(P, n);
    HC_LAUNCH_CHECK(ctx);
  }
  if (K > 0 && !getenv("HC_FILL_GRID_ROUNDS")) {
    // one grid pass lists the basin graph (per-tile deduplicated), then every Boruvka round runs on that list in
    // one cooperative launch; only if the list overflows (every cell its own basin) the grid rounds below take over
    static bool attr_set = false;
    if (!attr_set) {
      HC_CUDA(cudaFuncSetAttribute(k_edges<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, EDGES_SMEM));
      HC_CUDA(cudaFuncSetAttribute(k_edges<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, EDGES_SMEM));
      attr_set = true;
    }
    HC_CUDA(cudaMemsetAsync(cnt + 4, 0, 4, st));
    hc_tmap mz, mp;
    hc_tmap_2d(&mz, z, 4, 1, (uint64_t)nx, (uint64_t)ny, (uint64_t)nx * 4, WP, TH);
    hc_tmap_2d(&mp, P, 4, 0, (uint64_t)nx, (uint64_t)ny, (uint64_t)nx * 4, WP, TH);
    HC_PROF(ctx, st, "k_edges");
    if (mz.ok && mp.ok)
      k_edges<true><<<tg, EK_THREADS, EDGES_SMEM, st>>>(z, P, mz.map, mp.map, (int)ny, (int)nx, eA, eB, eW, cnt + 4, ecap, nopen, ctx->d_dbg);
    else
      k_edges<false><<<tg, EK_THREADS, EDGES_SMEM, st>>>(z, P, mz.map, mp.map, (int)ny, (int)nx, eA, eB, eW, cnt + 4, ecap, nopen, ctx->d_dbg);
    HC_LAUNCH_CHECK(ctx);
    HC_CUDA(cudaMemcpyAsync(ctx->h_mail + 4, cnt + 4, 4, cudaMemcpyDeviceToHost, st));
    HC_CUDA(cudaStreamSynchronize(st));
    ne = ctx->h_mail[4];
    if (getenv("HC_TRACE")) fprintf(stderr, "[fill] %u basins, edge list %u of %u\n", K, ne, ecap);
    if (ne <= ecap) {
      bor_pargs a;
      a.best = t.best; a.rep = t.rep; a.hp = t.hp; a.ptr = t.ptr; a.ptr2 = t.ptr2; a.lvl = t.lvl;
      a.eA = eA; a.eB = eB; a.eW = eW; a.mark = nullptr; a.ctl = cnt + 16;
      a.K1 = K1; a.nopen = nopen; a.ne = ne; a.msf = 0; a.idx_keys = 0; a.do_init = 0; a.max_rounds = 64; a.skip_pick = 0; a.both = 1;
      HC_TRY(bor_persistent_launch(ctx, a, nullptr, st));
      // rounds: diagnostic only, fetched without a synchronisation (hc_fill_stats reads it after the caller's own)
      HC_CUDA(cudaMemcpyAsync(ctx->h_mail + 32, cnt + 16 + 7, 4, cudaMemcpyDeviceToHost, st));
      ctx->fill_rounds = -1;
      if (getenv("HC_TRACE")) {
        static unsigned long long ts[100];
        u32 c3[3];
        HC_CUDA(cudaStreamSynchronize(st));
        HC_CUDA(cudaMemcpy(c3, cnt + 16 + 7, 12, cudaMemcpyDeviceToHost));
        HC_CUDA(cudaMemcpy(ts, cnt + 16 + 32, 96 * 8, cudaMemcpyDeviceToHost));
        fprintf(stderr, "[fill] list rounds %u, closed roots left %u; phase us:", c3[0], c3[1]);
        for (u32 i = 1; i < c3[2] && i < 96; i++) fprintf(stderr, " %.1f", (double)(ts[i] - ts[i - 1]) * 1e-3);
        fprintf(stderr, "\n");
      }
      active = 0;
    }
  }
  while (active > 0) {
    if (round > 64) { hc_set_error("fill: Boruvka did not converge"); return HC_ERR_INTERNAL; }
    HC_CUDA(cudaMemsetAsync(t.best, 0xFF, (size_t)K1 * 8, st));
    if (have_list) {
      HC_PROF(ctx, st, "k_minedge_list");
      k_minedge_list<<<(ne + 255) / 256, 256, 0, st>>>(eA, eB, eW, ne, t.rep, t.best, nopen);
      HC_LAUNCH_CHECK(ctx);
    } else if (round == 1) {
      HC_CUDA(cudaMemsetAsync(cnt + 4, 0, 4, st));
      HC_PROF(ctx, st, "k_minedge_emit");
      k_minedge<true, false><<<tg, NTHREADS, 0, st>>>(z, P, t.rep, t.best, (int)ny, (int)nx, eA, eB, eW, cnt + 4, ecap, nopen);
      HC_LAUNCH_CHECK(ctx);
      HC_CUDA(cudaMemcpyAsync(ctx->h_mail + 4, cnt + 4, 4, cudaMemcpyDeviceToHost, st));
    } else {
      HC_PROF(ctx, st, "k_minedge");
      if (round == 0) k_minedge<false, true><<<tg, NTHREADS, 0, st>>>(z, P, t.rep, t.best, (int)ny, (int)nx, eA, eB, eW, cnt + 4, ecap, nopen);
      else k_minedge<false, false><<<tg, NTHREADS, 0, st>>>(z, P, t.rep, t.best, (int)ny, (int)nx, eA, eB, eW, cnt + 4, ecap, nopen);
      HC_LAUNCH_CHECK(ctx);
    }
    if (!have_list) {
      // contraction of a grid round in one cooperative launch (hook, mutual pairs, pointer doubling to convergence,
      // merge): one host read-back per round instead of one per four doublings plus one for the counters
      bor_pargs a;
      a.best = t.best; a.rep = t.rep; a.hp = t.hp; a.ptr = t.ptr; a.ptr2 = t.ptr2; a.lvl = t.lvl;
      a.eA = nullptr; a.eB = nullptr; a.eW = nullptr; a.mark = nullptr; a.ctl = cnt + 16;
      a.K1 = K1; a.nopen = nopen; a.ne = 0; a.msf = 0; a.idx_keys = 0; a.do_init = 0; a.max_rounds = 1; a.skip_pick = 1; a.both = 0;
      HC_TRY(bor_persistent_launch(ctx, a, nullptr, st));
      HC_CUDA(cudaMemcpyAsync(ctx->h_mail, cnt + 16 + 8, 4, cudaMemcpyDeviceToHost, st));
      HC_CUDA(cudaStreamSynchronize(st));
      active = ctx->h_mail[0];
      merged = 1;
    } else {
      HC_PROF(ctx, st, "k_hook");
      k_hook<<<tgK, tb, 0, st>>>(t.rep, t.best, t.ptr2, t.lvl, K1, nopen);
      HC_LAUNCH_CHECK(ctx);
      HC_TRY(bor_contract(ctx, t, K1, nopen, cnt, &active, &merged, st));
    }
    if (round == 1 && !have_list) {  // the emit round's count arrived with the syncs above
      ne = ctx->h_mail[4];
      have_list = ne <= ecap;
    }
    if (getenv("HC_TRACE")) fprintf(stderr, "[fill] round %d: %u closed supernodes left of %u basins (edge list %u%s)\n", round, active, K, ne, have_list ? "" : " unused");
    round++;
    if (have_list && active > 0 && (size_t)K1 + ne < (size_t)(2u << 20)) {
      // small tables are latency bound: all remaining rounds in one cooperative launch (big ones keep the
      // per-round launches, whose kernels fill the machine)
      bor_pargs a;
      a.best = t.best; a.rep = t.rep; a.hp = t.hp; a.ptr = t.ptr; a.ptr2 = t.ptr2; a.lvl = t.lvl;
      a.eA = eA; a.eB = eB; a.eW = eW; a.mark = nullptr; a.ctl = cnt + 16;
      a.K1 = K1; a.nopen = nopen; a.ne = ne; a.msf = 0; a.idx_keys = 0; a.do_init = 0; a.max_rounds = 64; a.skip_pick = 0; a.both = 0;
      int more = 0;
      HC_TRY(bor_persistent_launch(ctx, a, &more, st));
      round += more;
      active = 0;
    }
  }
  if (ctx->fill_rounds != -1) ctx->fill_rounds = round;

  HC_PROF(ctx, st, "k_levels");
  k_levels<<<tgK, tb, 0, st>>>(t.hp, t.lvl, level0, banded ? troot : nullptr, K1, nopen);
  HC_LAUNCH_CHECK(ctx);
  fs->P = P; fs->level0 = level0; fs->troot = troot; fs->K1 = K1; fs->nopen = nopen; fs->cnt = cnt;
  fs->seam_seed = seam_seed; fs->eA = eA; fs->eB = eB; fs->eW = eW; fs->ecap = ecap;
  return HC_OK;
}

int fill_run(hc_ctx *ctx, const float *z, float *out, int64_t ny, int64_t nx, float nodata, int seed_mode,
             cudaStream_t st) {
  if (ny <= 0 || nx <= 0) return HC_OK;
  fill_state fs;
  hc_rows rw = {0, (int)ny, 0, 0};
  HC_TRY(fill_local(ctx, z, ny, nx, nodata, seed_mode, rw, &fs, st));
  HC_PROF(ctx, st, "k_apply");
  k_apply<<<ctx->sm_count * 8, 256, 0, st>>>(z, fs.P, fs.level0, out, (size_t)ny * nx);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

// ---------------------------------------------------------------------------------------------
// 5b. apply + D8 in ONE tile pass (the whole-raster chain): filled = max(z, level[basin]) is formed for the tile and
// its halo in shared memory and the D8 stencil (same arithmetic as d8.cu / oracle orc_d8) reads it from there, so the
// filled raster is written once and never read back for the directions: z + P in (8 B/cell), filled + 2 x dir out
// (6 B/cell) instead of apply (12) + D8 (6).  Windows of z and P arrive by TMA on pitched rasters (tma.cuh).
// ---------------------------------------------------------------------------------------------
struct fd8_params { double len[8]; float fact[8]; };

template <int TABLE, bool SLOPE, bool TMA>
__global__ void __launch_bounds__(NTHREADS)
k_apply_d8(const float *__restrict__ z, const u32 *__restrict__ P, const __grid_constant__ CUtensorMap mz,
           const __grid_constant__ CUtensorMap mp, const float *__restrict__ level0, float *__restrict__ filled,
           uint8_t *__restrict__ dir, uint8_t *__restrict__ dir2, float *__restrict__ slope, int ny, int nx, float nodata,
           fd8_params prm, hc_nf_lists nfl) {
  __shared__ __align__(128) float sf_box[TH * WP];
  __shared__ __align__(128) u32 sp_box[TH * WP];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ u32 s_nf;   // NOFLOW cells of the tile (listed for the flat pass while their 3 x 3 window is in registers)
  if (threadIdx.x == 0) s_nf = 0;
  float *sf = sf_box + WXO;
  u32 *sp = sp_box + WXO;
  const int r0 = blockIdx.y * T, c0 = blockIdx.x * T;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float qnan = __int_as_float(0x7fc00000);
  if (TMA) {
    if (tid == 0) {
      hc_mbar_init(&mbar, 1);
      hc_mbar_expect_tx(&mbar, 2u * TH * WP * 4u);
      hc_tma_load_2d(sf_box, &mz, &mbar, c0 - 1 - WXO, r0 - 1);
      hc_tma_load_2d(sp_box, &mp, &mbar, c0 - 1 - WXO, r0 - 1);
    }
    __syncthreads();
    hc_mbar_wait(&mbar, 0);
  }
  // window of FILLED elevations (NaN: no cell)
  for (int lr = warp; lr < TH; lr += NTHREADS / 32) {
    const int r = r0 + lr - 1;
#pragma unroll
    for (int m = 0; m < 3; m++) {
      const int lc = lane + 32 * m;
      if (lc >= TH) break;
      const int c = c0 + lc - 1;
      const int i = lr * WP + lc;
      const bool inb = r >= 0 && r < ny && c >= 0 && c < nx;
      float v;
      u32 p;
      if (TMA) { v = sf[i]; p = sp[i]; }
      else {
        const size_t g = inb ? (size_t)r * nx + c : 0;
        v = __ldg(&z[g]);
        p = __ldg(&P[g]);
      }
      const u32 id = p & 0x7FFFFFFFu;
      if (!inb || v == nodata || v != v) v = qnan;
      else if (id != 0u && id != HC_LAB_NODATA) v = fmaxf(v, __ldg(&level0[id]));
      sf[i] = v;
    }
  }
  __syncthreads();
  // column walk: thread -> column lc = tid % 64, rows [16 * (tid / 64), +16), 3x3 window in registers
  const int lc = tid & (T - 1), lr0 = (tid >> 6) * (T / 4);
  const int c = c0 + lc;
  const size_t tile = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
  if (c < nx) {
  float w[3][3];
  {
    const float *row = &sf[lr0 * WP + lc];
    w[0][0] = row[0]; w[0][1] = row[1]; w[0][2] = row[2];
    w[1][0] = row[WP]; w[1][1] = row[WP + 1]; w[1][2] = row[WP + 2];
  }
#pragma unroll 4
  for (int k = 0; k < T / 4; k++) {
    const int lr = lr0 + k, r = r0 + lr;
    if (r >= ny) break;
    const float *row = &sf[(lr + 2) * WP + lc];
    w[2][0] = row[0]; w[2][1] = row[1]; w[2][2] = row[2];
    const size_t g = (size_t)r * nx + c;
    const float v = w[1][1];
    uint8_t code;
    float sl;
    if (v != v) {
      code = HC_DIR_NODATA;
      sl = -1.0f;
      filled[g] = __ldg(&z[g]);   // a nodata cell keeps its own value
    } else {
      filled[g] = v;
      int best = -1;
      bool contaminated = false;
      double smax = 0.0;
      float smaxf = 0.0f;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const float zn = w[1 + c_dr[TABLE][j]][1 + c_dc[TABLE][j]];
        contaminated |= zn != zn;   // (a NaN neighbour never wins below: every comparison with NaN is false)
        if (TABLE == HC_TABLE_WBT) {
          if (zn < v) {
            double s_ = ((double)v - (double)zn) / prm.len[j];
            if (s_ > smax) { smax = s_; best = j; }
          }
        } else {
          float s_ = __fmul_rn(prm.fact[j], __fsub_rn(v, zn));
          if (s_ > smaxf) { smaxf = s_; best = j; }
        }
      }
      if (TABLE == HC_TABLE_TAUDEM && contaminated) {
        code = HC_DIR_NODATA;
        sl = -1.0f;
      } else {
        code = best < 0 ? (uint8_t)HC_DIR_NOFLOW : (uint8_t)c_code[TABLE][best];
        sl = TABLE == HC_TABLE_WBT ? (float)smax : smaxf;
      }
    }
    dir[g] = code;
    if (dir2) dir2[g] = code;
    if (SLOPE) slope[g] = sl;
    if (code == HC_DIR_NOFLOW && nfl.entries) {
      // the flat pass's record of this cell: equal-elevation neighbours, "touches higher ground"
      u32 eq = 0, hi = 0;
      int kb = 0;
#pragma unroll
      for (int dr = 0; dr < 3; dr++)
#pragma unroll
        for (int dc = 0; dc < 3; dc++) {
          if (dr == 1 && dc == 1) continue;
          const float zn = w[dr][dc];
          if (zn == v) eq |= 1u << kb;
          else if (zn > v) hi = HC_NF_HI;
          kb++;
        }
      const u32 pos = atomicAdd(&s_nf, 1u);
      if (pos < HC_NFCAP) {
        nfl.entries[tile * HC_NFCAP + pos] = (unsigned short)((u32)(lr * T + lc) | hi);
        nfl.eq[tile * HC_NFCAP + pos] = (uint8_t)eq;
      }
    }
#pragma unroll
    for (int j = 0; j < 3; j++) { w[0][j] = w[1][j]; w[1][j] = w[2][j]; }
  }
  }
  if (nfl.counts) {
    __syncthreads();
    if (tid == 0) nfl.counts[tile] = s_nf;
  }
}

template <int TABLE, bool SLOPE>
static int apply_d8_launch(hc_ctx *ctx, const float *z, const fill_state &fs, float *filled, uint8_t *dir, uint8_t *dir2,
                           float *slope, int64_t ny, int64_t nx, float nodata, const fd8_params &prm, cudaStream_t st,
                           hc_nf_lists nfl) {
  dim3 tg((unsigned)((nx + T - 1) / T), (unsigned)((ny + T - 1) / T));
  hc_tmap mz, mp;
  hc_tmap_2d(&mz, z, 4, 1, (uint64_t)nx, (uint64_t)ny, (uint64_t)nx * 4, WP, TH);
  hc_tmap_2d(&mp, fs.P, 4, 0, (uint64_t)nx, (uint64_t)ny, (uint64_t)nx * 4, WP, TH);
  HC_PROF(ctx, st, "k_apply_d8");
  if (mz.ok && mp.ok)
    k_apply_d8<TABLE, SLOPE, true><<<tg, NTHREADS, 0, st>>>(z, fs.P, mz.map, mp.map, fs.level0, filled, dir, dir2, slope, (int)ny, (int)nx, nodata, prm, nfl);
  else
    k_apply_d8<TABLE, SLOPE, false><<<tg, NTHREADS, 0, st>>>(z, fs.P, mz.map, mp.map, fs.level0, filled, dir, dir2, slope, (int)ny, (int)nx, nodata, prm, nfl);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

int flats_core(hc_ctx *ctx, const float *z, const uint8_t *dirA, uint8_t *dirB, int64_t ny, int64_t nx, float nodata,
               int table, cudaStream_t st, const hc_nf_lists *pre = nullptr);

// fill -> (apply + D8 fused) -> flat resolution, one arena generation: the head of hc_fill_d8_accum_*
int fill_d8_flats_run(hc_ctx *ctx, const float *z, float *filled, uint8_t *dir, float *slope, int64_t ny, int64_t nx,
                      float nodata, double dx, double dy, int seed_mode, int table, int resolve_flats, cudaStream_t st) {
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (table != HC_TABLE_WBT && table != HC_TABLE_TAUDEM) { hc_set_error("d8: bad table %d", table); return HC_ERR_ARG; }
  if (!(dx > 0) || !(dy > 0)) { hc_set_error("d8: dx, dy must be > 0"); return HC_ERR_ARG; }
  fill_state fs;
  hc_rows rw = {0, (int)ny, 0, 0};
  HC_TRY(fill_local(ctx, z, ny, nx, nodata, seed_mode, rw, &fs, st));
  uint8_t *dirA = nullptr;
  if (resolve_flats) {
    dirA = (uint8_t *)arena_take(ctx, (size_t)ny * nx);
    if (!dirA) return HC_ERR_NOMEM;
  }
  static const int DR[2][8] = {{-1, 0, 1, 1, 1, 0, -1, -1}, {0, -1, -1, -1, 0, 1, 1, 1}};
  static const int DC[2][8] = {{1, 1, 1, 0, -1, -1, -1, 0}, {1, 1, 0, -1, -1, -1, 0, 1}};
  fd8_params prm;
  for (int k = 0; k < 8; k++) {
    double l = sqrt((double)(DC[table][k] * DC[table][k]) * dx * dx + (double)(DR[table][k] * DR[table][k]) * dy * dy);
    prm.len[k] = l;
    prm.fact[k] = (float)(1.0 / l);
  }
  // with flats: D8 writes dirA (kept as the flat pass's read-only input) and dir; the flat pass rewrites NOFLOW cells of dir
  uint8_t *d1 = resolve_flats ? dirA : dir, *d2 = resolve_flats ? dir : nullptr;
  // ... and lists the NOFLOW cells per tile for the flat pass while it has their windows in registers
  hc_nf_lists nfl = {nullptr, nullptr, nullptr};
  if (resolve_flats) {
    const size_t tiles = (size_t)((nx + T - 1) / T) * ((ny + T - 1) / T);
    nfl.entries = (unsigned short *)arena_take(ctx, tiles * HC_NFCAP * 2);
    nfl.eq = (uint8_t *)arena_take(ctx, tiles * HC_NFCAP);
    nfl.counts = (u32 *)arena_take(ctx, tiles * 4);
    if (!nfl.entries || !nfl.eq || !nfl.counts) return HC_ERR_NOMEM;
  }
  if (table == HC_TABLE_WBT) {
    if (slope) HC_TRY((apply_d8_launch<HC_TABLE_WBT, true>(ctx, z, fs, filled, d1, d2, slope, ny, nx, nodata, prm, st, nfl)));
    else HC_TRY((apply_d8_launch<HC_TABLE_WBT, false>(ctx, z, fs, filled, d1, d2, slope, ny, nx, nodata, prm, st, nfl)));
  } else {
    if (slope) HC_TRY((apply_d8_launch<HC_TABLE_TAUDEM, true>(ctx, z, fs, filled, d1, d2, slope, ny, nx, nodata, prm, st, nfl)));
    else HC_TRY((apply_d8_launch<HC_TABLE_TAUDEM, false>(ctx, z, fs, filled, d1, d2, slope, ny, nx, nodata, prm, st, nfl)));
  }
  if (!resolve_flats) return HC_OK;
  if (resolve_flats != 2) return flats_core(ctx, filled, dirA, dir, ny, nx, nodata, table, st, &nfl);
  // chain form: flats in two halves with the accumulation's tile phase split around the slow tier
  const uint8_t *tile_pending = nullptr;
  int n_pending = 0;
  HC_TRY(flats_whole_begin(ctx, filled, dirA, dir, ny, nx, nodata, table, st, &tile_pending, &n_pending, &nfl));
  HC_TRY(accum_split_prepare(ctx, dir, ctx->chain_acc, ny, nx));
  if (!n_pending) {
    HC_TRY(accum_split_tiles(ctx, table, nullptr, 0, st));
    return accum_split_finish(ctx, table, st);
  }
  // the tile phase of the settled tiles is queued first, on the caller's stream; the slow tier follows on the context's
  // HIGH-PRIORITY stream, so its small kernels overtake the queued CTAs of the big one instead of waiting behind them
  HC_CUDA(cudaEventRecord(ctx->ev_fork2, st));
  HC_CUDA(cudaStreamWaitEvent(ctx->s_aux2, ctx->ev_fork2, 0));
  HC_TRY(accum_split_tiles(ctx, table, tile_pending, 0, st));
  HC_TRY(flats_whole_finish(ctx, ctx->s_aux2));
  HC_CUDA(cudaEventRecord(ctx->ev_join2, ctx->s_aux2));
  HC_CUDA(cudaStreamWaitEvent(st, ctx->ev_join2, 0));
  HC_TRY(accum_split_tiles(ctx, table, tile_pending, 1, st));
  return accum_split_finish(ctx, table, st);
}

// The whole chain (hc_fill_d8_accum_*): as above, but the accumulation's tile-local phase of every tile WITHOUT pending
// flat cells (84 % of the tiles on the C2 recipe) runs on a second stream beside the flats' slow tier -- whose rounds are
// latency bound and leave most SMs idle -- and only the remaining tiles wait for the final directions.  A pending cell
// holds code 254 until then, which the tile-local phase of a neighbouring tile reads only as "not nodata".
int fill_d8_flats_accum_run(hc_ctx *ctx, const float *z, float *filled, uint8_t *dir, float *slope, uint32_t *acc, int64_t ny,
                            int64_t nx, float nodata, double dx, double dy, int seed_mode, int table, int resolve_flats,
                            cudaStream_t st) {
  if (ny <= 0 || nx <= 0) return HC_OK;
  // (measured on B200: the slow tier's persistent CTAs hold most of every SM's shared memory, so the accumulation's tile
  // phase gains nothing beside it -- C2 6.91 ms in sequence, 7.04 ms overlapped; the split form stays behind HC_ACC_OVERLAP)
  if (!resolve_flats || !getenv("HC_ACC_OVERLAP")) {
    HC_TRY(fill_d8_flats_run(ctx, z, filled, dir, slope, ny, nx, nodata, dx, dy, seed_mode, table, resolve_flats, st));
    return accum_run(ctx, dir, acc, ny, nx, table, st);
  }
  ctx->chain_acc = acc;
  return fill_d8_flats_run(ctx, z, filled, dir, slope, ny, nx, nodata, dx, dy, seed_mode, table, 2, st);
}

// =============================================================================================
// Row-band sharding (Barnes 2016 "Parallel priority-flood depression filling for
// trillion cell DEMs" with a band as the tile).
//
//   local   every seam cell (first / last own row next to another band) is the root of its own
//           terminal basin; phase 1 contracts the closed basins into the open nodes {outside,
//           terminals} and yields level0 = minimax level relative to the nearest open node and
//           troot = that node.  The graph among the open nodes (edge = lowest locally-filled pass
//           between two terminal watersheds) is reduced to its minimum spanning forest, which
//           preserves every pairwise minimax distance; the band publishes those <= 2nx edges plus
//           the <= 3nx cross-seam edges of its bottom seam.
//   finish  every rank solves the union of all bands' edges for the terminals' true levels
//           (redundantly; the graph is tiny) and applies out = max(z, level0, level[troot]).
// Exactness: level(B) = max(level0(B), level(troot(B))) because a closed component's outgoing
// edges are all band-local (cross-seam edges touch terminals only) — see DESIGN.md §6.
// =============================================================================================
__global__ void __launch_bounds__(NTHREADS)
k_term_edges(const float *__restrict__ z, const u32 *__restrict__ P, const u32 *__restrict__ troot,
             const float *__restrict__ level0, int ny, int nx, u32 *__restrict__ eA, u32 *__restrict__ eB,
             float *__restrict__ eW, u32 *__restrict__ ecount, u32 ecap) {
  __shared__ float sz[TH * TH];
  __shared__ u32 sa[TH * TH];
  __shared__ u32 s_tot, s_base;
  const int r0 = blockIdx.y * T, c0 = blockIdx.x * T;
  const int tid = threadIdx.x;
  if (tid == 0) s_tot = 0;
  for (int i = tid; i < TH * TH; i += NTHREADS) {
    int lr = i / TH, lc = i - lr * TH;
    int r = r0 + lr - 1, c = c0 + lc - 1;
    bool inb = r >= 0 && r < ny && c >= 0 && c < nx;
    size_t g = inb ? (size_t)r * nx + c : 0;
    u32 id = __ldg(&P[g]) & 0x7FFFFFFFu;
    float v = __ldg(&z[g]);
    u32 a = HC_NONE;
    if (inb && id != HC_LAB_NODATA) {
      a = id ? __ldg(&troot[id]) : 0u;
      if (id) v = fmaxf(v, __ldg(&level0[id]));  // locally filled elevation
    }
    sz[i] = v;
    sa[i] = a;
  }
  __syncthreads();
  // an adjacent pair in different terminal watersheds is emitted once, by the cell with the smaller node
  u32 cnt = 0;
  for (int i = tid; i < T * T; i += NTHREADS) {
    int lr = i >> 6, lc = i & (T - 1);
    int ci = (lr + 1) * TH + lc + 1;
    u32 a = sa[ci];
    if (a == HC_NONE) continue;
#pragma unroll
    for (int dr = -1; dr <= 1; dr++)
#pragma unroll
      for (int dc = -1; dc <= 1; dc++) {
        if (dr == 0 && dc == 0) continue;
        u32 b = sa[ci + dr * TH + dc];
        if (b != HC_NONE && a < b) cnt++;
      }
  }
  u32 off = cnt ? atomicAdd(&s_tot, cnt) : 0u;
  __syncthreads();
  if (tid == 0) s_base = s_tot ? atomicAdd(ecount, s_tot) : 0u;
  __syncthreads();
  if (!cnt || (u64)s_base + s_tot > (u64)ecap) return;  // overflow: host reports it
  u32 o = s_base + off;
  for (int i = tid; i < T * T; i += NTHREADS) {
    int lr = i >> 6, lc = i & (T - 1);
    int ci = (lr + 1) * TH + lc + 1;
    u32 a = sa[ci];
    if (a == HC_NONE) continue;
    float v = sz[ci];
#pragma unroll
    for (int dr = -1; dr <= 1; dr++)
#pragma unroll
      for (int dc = -1; dc <= 1; dc++) {
        if (dr == 0 && dc == 0) continue;
        int ni = ci + dr * TH + dc;
        u32 b = sa[ni];
        if (b == HC_NONE || !(a < b)) continue;
        eA[o] = a;
        eB[o] = b;
        eW[o] = fmaxf(v, sz[ni]);
        o++;
      }
  }
}

// seam cells that are seeds themselves: edge (outside, terminal, z)
__global__ void k_seam_seed_edges(const float *__restrict__ z, const uint8_t *__restrict__ seam_seed, int ny, int nx,
                                  u32 *eA, u32 *eB, float *eW, u32 *ecount, u32 ecap) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * nx || !seam_seed[i]) return;
  int side = i / nx, c = i - side * nx;
  float v = z[(size_t)(side ? ny - 1 : 0) * nx + c];
  u32 o = atomicAdd(ecount, 1u);
  if (o >= ecap) return;
  eA[o] = 0u;
  eB[o] = 1u + (u32)i;
  eW[o] = v;
}

// out triples (a, b, w bits) in GLOBAL node ids: 0 = outside, 1 + (band*2 + side)*nx + c = seam cell
__global__ void k_compact_marked(const u32 *__restrict__ eA, const u32 *__restrict__ eB, const float *__restrict__ eW,
                                 const uint8_t *__restrict__ mark, u32 ne, u32 gbase, u32 *out, u32 *ocount, u32 ocap) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ne || !mark[i]) return;
  u32 o = atomicAdd(ocount, 1u);
  if (o >= ocap) return;
  u32 a = eA[i], b = eB[i];
  out[3 * (size_t)o] = a ? gbase + a : 0u;
  out[3 * (size_t)o + 1] = b ? gbase + b : 0u;
  out[3 * (size_t)o + 2] = __float_as_uint(eW[i]);
}

// cross-seam edges of the bottom seam: own cell (ny-1, c) -- halo cell (ny, c+d), weight max(z, z)
__global__ void k_cross_edges(const float *__restrict__ z, int ny, int nx, float nodata, u32 g_own, u32 g_next,
                              u32 *out, u32 *ocount, u32 ocap) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nx) return;
  float v = z[(size_t)(ny - 1) * nx + c];
  if (d_is_nodata(v, nodata)) return;
  for (int d = -1; d <= 1; d++) {
    int cc = c + d;
    if (cc < 0 || cc >= nx) continue;
    float w = z[(size_t)ny * nx + cc];
    if (d_is_nodata(w, nodata)) continue;
    u32 o = atomicAdd(ocount, 1u);
    if (o >= ocap) return;
    out[3 * (size_t)o] = g_own + (u32)c;
    out[3 * (size_t)o + 1] = g_next + (u32)cc;
    out[3 * (size_t)o + 2] = __float_as_uint(fmaxf(v, w));
  }
}

__global__ void k_unpack_edges(const u32 *__restrict__ in, u32 ne, u32 *eA, u32 *eB, float *eW) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ne) return;
  eA[i] = in[3 * (size_t)i];
  eB[i] = in[3 * (size_t)i + 1];
  eW[i] = __uint_as_float(in[3 * (size_t)i + 2]);
}

// final level of a local basin = max(local level, true level of the terminal it hangs under)
__global__ void k_level_final(const float *__restrict__ level0, const u32 *__restrict__ troot,
                              const float *__restrict__ tlevel, u32 gbase, float *__restrict__ lf, u32 K1) {
  u32 b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= K1) return;
  float m = level0[b];
  u32 t = troot[b];
  if (t) m = fmaxf(m, tlevel[gbase + t]);
  lf[b] = m;
}

int64_t band_fill_edge_cap(int64_t nx) { return 5 * nx + 16; }

int band_fill_local(hc_ctx *ctx, const float *z, int64_t nb, int64_t nx, float nodata, int seed_mode, int flags,
                    int64_t band_index, uint32_t *edges_out, cudaStream_t st, const uint8_t *ext) {
  hc_band_state *bs = &ctx->band;
  hc_rows rw = {(flags & 1) ? -1 : 0, (int)nb + ((flags & 2) ? 1 : 0), flags & 1, (flags >> 1) & 1};
  const int64_t cap = band_fill_edge_cap(nx);
  fill_state fs;
  if ((double)(band_index + 2) * 2.0 * (double)nx >= 4294967000.0) { hc_set_error("too many seam cells"); return HC_ERR_TOO_LARGE; }
  if (!(flags & 3)) {
    // a single band: no terminals at all, nothing to publish
    HC_TRY(fill_local(ctx, z, nb, nx, nodata, seed_mode, rw, &fs, st, ext));
    HC_CUDA(cudaMemsetAsync(edges_out, 0, (size_t)cap * 12, st));
  } else {
    HC_TRY(fill_local(ctx, z, nb, nx, nodata, seed_mode, rw, &fs, st, ext));
    HC_CUDA(cudaMemsetAsync(edges_out, 0, (size_t)cap * 12, st));
    const int tx = (int)((nx + T - 1) / T), ty = (int)((nb + T - 1) / T);
    dim3 tg(tx, ty);
    u32 *cnt = fs.cnt;
    HC_CUDA(cudaMemsetAsync(cnt + 4, 0, 4, st));
    HC_PROF(ctx, st, "k_term_edges");
    k_term_edges<<<tg, NTHREADS, 0, st>>>(z, fs.P, fs.troot, fs.level0, (int)nb, (int)nx, fs.eA, fs.eB, fs.eW, cnt + 4, fs.ecap);
    HC_LAUNCH_CHECK(ctx);
    HC_PROF(ctx, st, "k_seam_seed_edges");
    k_seam_seed_edges<<<(unsigned)((2 * nx + 255) / 256), 256, 0, st>>>(z, fs.seam_seed, (int)nb, (int)nx, fs.eA, fs.eB, fs.eW, cnt + 4, fs.ecap);
    HC_LAUNCH_CHECK(ctx);
    HC_CUDA(cudaMemcpyAsync(ctx->h_mail, cnt + 4, 4, cudaMemcpyDeviceToHost, st));
    HC_CUDA(cudaStreamSynchronize(st));
    const u32 ne = ctx->h_mail[0];
    ctx->band_term_edges = ne;
    if (ne > fs.ecap) { hc_set_error("banded fill: %u terminal edges exceed the list capacity %u", ne, fs.ecap); return HC_ERR_TOO_LARGE; }
    // minimum spanning forest of the open-node graph
    bor_tabs t2;
    HC_TRY(bor_tabs_alloc(ctx, &t2, fs.nopen));
    uint8_t *mark = (uint8_t *)arena_take(ctx, ne ? ne : 1);
    if (!mark) return HC_ERR_NOMEM;
    int rounds = 0;
    HC_TRY(bor_list_solve(ctx, t2, fs.nopen, 0u, fs.eA, fs.eB, fs.eW, ne, 1, mark, cnt, &rounds, st));
    ctx->band_msf_rounds = rounds;
    HC_CUDA(cudaMemsetAsync(cnt + 6, 0, 4, st));
    const u32 gbase = (u32)(band_index * 2 * nx);  // global id of local node t >= 1 is gbase + t
    if (ne) {
      HC_PROF(ctx, st, "k_compact_marked");
      k_compact_marked<<<(ne + 255) / 256, 256, 0, st>>>(fs.eA, fs.eB, fs.eW, mark, ne, gbase, edges_out, cnt + 6, (u32)cap);
      HC_LAUNCH_CHECK(ctx);
    }
    if (flags & 2) {
      HC_PROF(ctx, st, "k_cross_edges");
      k_cross_edges<<<(unsigned)((nx + 255) / 256), 256, 0, st>>>(z, (int)nb, (int)nx, nodata, gbase + 1u + (u32)nx,
                                                                   (u32)((band_index + 1) * 2 * nx) + 1u, edges_out, cnt + 6, (u32)cap);
      HC_LAUNCH_CHECK(ctx);
    }
  }
  bs->flags = flags; bs->nb = nb; bs->nx = nx; bs->band_index = band_index;
  bs->z = z; bs->P = fs.P; bs->level0 = fs.level0; bs->troot = fs.troot; bs->K1 = fs.K1; bs->nopen = fs.nopen;
  bs->cnt = fs.cnt;
  bs->fill_ready = 1;
  return HC_OK;
}

int band_fill_finish(hc_ctx *ctx, const uint32_t *all_edges, int64_t nbands, float *out, cudaStream_t st) {
  hc_band_state *bs = &ctx->band;
  if (!bs->fill_ready) { hc_set_error("hc_band_fill_finish without hc_band_fill_local"); return HC_ERR_ARG; }
  bs->fill_ready = 0;
  const int64_t nx = bs->nx, nb = bs->nb;
  const size_t n = (size_t)nb * nx;
  const float *level = bs->level0;
  if (bs->flags & 3) {
    const int64_t cap = band_fill_edge_cap(nx);
    const u32 ne = (u32)(nbands * cap);
    const u32 K1g = (u32)(1 + nbands * 2 * nx);
    u32 *eA = (u32 *)arena_take(ctx, (size_t)ne * 4);
    u32 *eB = (u32 *)arena_take(ctx, (size_t)ne * 4);
    float *eW = (float *)arena_take(ctx, (size_t)ne * 4);
    float *lf = (float *)arena_take(ctx, (size_t)bs->K1 * 4);
    if (!eA || !eB || !eW || !lf) return HC_ERR_NOMEM;
    HC_PROF(ctx, st, "k_unpack_edges");
    k_unpack_edges<<<(ne + 255) / 256, 256, 0, st>>>(all_edges, ne, eA, eB, eW);
    HC_LAUNCH_CHECK(ctx);
    bor_tabs tg;
    HC_TRY(bor_tabs_alloc(ctx, &tg, K1g));
    int rounds = 0;
    HC_TRY(bor_list_solve(ctx, tg, K1g, 1u, eA, eB, eW, ne, 0, nullptr, bs->cnt, &rounds, st));
    ctx->band_global_rounds = rounds;
    float *tlevel = (float *)tg.ptr2;
    const int tb = 256;
    HC_PROF(ctx, st, "k_levels");
    k_levels<<<(K1g + tb - 1) / tb, tb, 0, st>>>(tg.hp, tg.lvl, tlevel, nullptr, K1g, 1u);
    HC_LAUNCH_CHECK(ctx);
    HC_PROF(ctx, st, "k_level_final");
    k_level_final<<<(bs->K1 + tb - 1) / tb, tb, 0, st>>>(bs->level0, bs->troot, tlevel, (u32)(bs->band_index * 2 * nx), lf, bs->K1);
    HC_LAUNCH_CHECK(ctx);
    level = lf;
  }
  HC_PROF(ctx, st, "k_apply");
  k_apply<<<ctx->sm_count * 8, 256, 0, st>>>(bs->z, bs->P, level, out, n);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}
