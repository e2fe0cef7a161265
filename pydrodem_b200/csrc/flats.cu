// flats.cu — drainage directions over flats (Barnes, Lehman & Mulla 2014; the two-gradient idea of
// Garbrecht & Martz 1997 that TauDEM D8FlowDir applies to filled pits, run_aws.py:594-603).
//
// Serial form (oracle/hydro_oracle.c orc_resolve_flats): label flats, BFS "away from higher"
// from the high edges, BFS "towards lower" from the low edges, combine into a flat mask
// FlatHeight - away + 2*towards, then D8 on the mask inside each flat.
//
// Parallel form used here.  Both BFS fields are geodesic (8-connected, unit-step) distances
// through the NOFLOW cells of one elevation, i.e. the unique fixed point of
//      d(c) = 1 + min over same-flat neighbours d(n),  sources fixed at 1,
// so they can be computed by chaotic relaxation.  No flat labels are needed: two 8-adjacent cells
// of equal elevation are always in the same flat, FlatHeight is a per-flat constant that cancels
// in every comparison the masked D8 makes, and a flat without a low edge is exactly one whose
// cells the "towards lower" field never reaches.
//
// Two tiers, both driven by per-tile RECORDS of the NOFLOW cells (hc_nf_lists: local index, which of the 8 neighbours
// share the elevation, "touches higher ground") that the fused apply + D8 pass writes while it has each cell's 3 x 3
// window in registers (k_flat_list builds them for a D8 raster handed in by a caller / a row band):
//   fast  k_flat_fast_sparse: one CTA per 64x64 tile gathers the records of its tile plus an 8-cell apron (80x80
//         region), relaxes to the fixed point in shared memory and writes final directions for every flat that lies
//         completely inside the apron (on a filled DEM nearly all of them).  No elevation is read.
//   slow  flats that reach the apron's rim are marked pending (code 254) and solved by tile relaxation against two
//         global distance rasters that hold state only at NOFLOW cells (k_flat_init_lists, k_flat_relax,
//         k_flat_dirs_lists) on the tiles that hold pending cells and their neighbours.  A cell WITH a direction next to an
//         equal NOFLOW cell is a low edge: constant 1, asked of the D8 raster, never stored.  A device queue holds the
//         dirty tiles of a round (one cooperative launch runs all rounds); tiles with <= 1024 NOFLOW cells relax from
//         their records in shared memory (relax_light), lake tiles through a shared-memory window with directional
//         cone sweeps (relax_tile).  A tile is started only if it holds part of an open flat and re-queued only if
//         one of its rim cells can improve.
#include <stdlib.h>
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

#define T HC_TILE
#define TH (T + 2)
#define NTHREADS 256
#define HC_DIR_PENDING 254
#define FL_LCAP 1024  // NOFLOW cells a tile may hold to take the light relaxation path

// ================================================================================================
// fast tier
// ================================================================================================
#define FH 8                 // apron
#define FR (T + 2 * FH)      // region edge (80)
#define F_NF 0x4000u         // flags live in the top bits of the 16-bit "lo" word
#define F_OPEN 0x8000u
#define F_LO 0x3FFFu
#define FLCAP 1536           // NOFLOW cells a tile may hold before it is left to the slow tier
#define FSWEEPS 40           // sweep cap: a flat that fits the apron converges well inside it
#define FS_BATCH 1           // relaxation passes between two block-wide votes (4 measured slower: 0.78 vs 0.63 ms, most flats settle in 2-3 passes)

// ------------------------------------------------------------------------------------------------
// Sparse form of the fast tier.  On a filled DEM only a few percent of the cells are NOFLOW (3.4 % on the C2
// recipe), yet k_flat_fast stages and classifies all 80 x 80 cells of its region to find them.  k_flat_list
// compacts the NOFLOW cells of every tile once (one byte read per cell, raster order inside the tile);
// k_flat_fast_sparse then builds its region's work list from the lists of the 3 x 3 tiles around it, reads the
// 3 x 3 elevations of just those cells, and keeps the per-cell state in the same 80 x 80 shared grid — same
// rules, same result, a small fraction of the instructions.
// ------------------------------------------------------------------------------------------------
#define NFCAP HC_NFCAP   // listed NOFLOW cells per tile; a tile with more is "dense": its region is left to the slow tier

__constant__ const int c_k_of_j[2][8] = {{2, 4, 7, 6, 5, 3, 0, 1}, {4, 2, 1, 0, 3, 5, 6, 7}};   // 3 x 3 slot of table neighbour j

// generic entry points (a D8 raster handed in by the caller, row bands): the lists k_apply_d8 builds on the fly in the
// whole-raster chain.  One CTA per tile: compact the NOFLOW cells (one byte read per cell), then look at the 3 x 3
// elevations of just those cells for the record (equal-elevation neighbours, "touches higher ground").
__global__ void __launch_bounds__(256)
k_flat_list(const float *__restrict__ z, const uint8_t *__restrict__ dir, int ny, int nx, float nodata, hc_nf_lists nfl, hc_rows rw) {
  __shared__ u32 s_w[8];
  __shared__ unsigned short s_li[NFCAP];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = tid >> 2, lc0 = (tid & 3) * 16;
  const int r = blockIdx.y * T + lr, c0 = blockIdx.x * T + lc0;
  u32 m = 0;
  if (r < ny && c0 < nx) {
    const uint8_t *row = dir + (size_t)r * nx + c0;
    if (c0 + 16 <= nx && (((uintptr_t)row) & 15u) == 0) {
      const uint4 q = __ldg((const uint4 *)row);
      const u32 w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; j++)
#pragma unroll
        for (int b = 0; b < 4; b++)
          if (((w4[j] >> (8 * b)) & 0xFFu) == HC_DIR_NOFLOW) m |= 1u << (4 * j + b);
    } else {
      for (int j = 0; j < 16 && c0 + j < nx; j++)
        if (__ldg(row + j) == HC_DIR_NOFLOW) m |= 1u << j;
    }
  }
  const u32 cnt = __popc(m);
  u32 inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { u32 t_ = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t_; }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  u32 base = 0, total = 0;
  for (int w = 0; w < 8; w++) { if (w < warp) base += s_w[w]; total += s_w[w]; }
  u32 off = base + inc - cnt;
  const size_t tile = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
  while (m) {
    const int j = __ffs(m) - 1;
    m &= m - 1;
    if (off < NFCAP) s_li[off] = (unsigned short)(lr * T + lc0 + j);
    off++;
  }
  if (tid == 0) nfl.counts[tile] = total;
  __syncthreads();
  const u32 nl = total < NFCAP ? total : NFCAP;
  for (u32 k = tid; k < nl; k += 256) {
    const int li = s_li[k];
    const int rr = blockIdx.y * T + (li >> 6), cc = blockIdx.x * T + (li & (T - 1));
    const float v = __ldg(z + (long long)rr * nx + cc);
    u32 eq = 0, hi = 0;
    int kb = 0;
#pragma unroll
    for (int dr = -1; dr <= 1; dr++)
#pragma unroll
      for (int dc = -1; dc <= 1; dc++) {
        if (!dr && !dc) continue;
        const int nr = rr + dr, nc = cc + dc;
        if (nr >= rw.rlo && nr < rw.rhi && nc >= 0 && nc < nx) {
          const float zn = __ldg(z + (long long)nr * nx + nc);
          if (zn != nodata) {
            if (zn == v) eq |= 1u << kb;
            else if (zn > v) hi = HC_NF_HI;
          }
        }
        kb++;
      }
    nfl.entries[tile * NFCAP + k] = (unsigned short)((u32)li | hi);
    nfl.eq[tile * NFCAP + k] = (uint8_t)eq;
  }
}

#define FS_THREADS 128
#define FS_SMEM (FR * FR * 4 + FLCAP * 4 + 16)

template <int TABLE>
__global__ void __launch_bounds__(FS_THREADS)
k_flat_fast_sparse(const uint8_t *__restrict__ dirA, uint8_t *__restrict__ dirB, const hc_nf_lists nfl,
                   uint8_t *__restrict__ tile_pending, int ny, int nx, unsigned long long *dbg, hc_rows rw) {
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned short *slo = (unsigned short *)smem;                          // [OPEN | NF | lo:14] per region cell
  unsigned short *shi = (unsigned short *)(smem + FR * FR * 2);
  unsigned short *list = (unsigned short *)(smem + FR * FR * 4);         // region index of the work cells
  unsigned short *lmask = (unsigned short *)(smem + FR * FR * 4 + FLCAP * 2);   // [low-edge neighbours : 8 | NOFLOW neighbours : 8], 3 x 3 slots
  u32 *s_n = (u32 *)(smem + FR * FR * 4 + FLCAP * 4);                    // [0] work cells, [1] give up
  const int tid = threadIdx.x;
  const int by = blockIdx.y, bx = blockIdx.x, ty = gridDim.y, tx = gridDim.x;
  const int r0 = by * T - FH, c0 = bx * T - FH;
  // the nine tiles' counts and every thread's first two records of each tile are requested before anything is looked
  // at: the gather is a handful of dependent round trips to L2 / HBM, and with one tile after the other (count ->
  // records -> next tile's count ...) those round trips were most of the kernel's time
  u32 cnt9[9];
#pragma unroll
  for (int t9 = 0; t9 < 9; t9++) {
    const int yy = by + t9 / 3 - 1, xx = bx + t9 % 3 - 1;
    cnt9[t9] = (yy >= 0 && yy < ty && xx >= 0 && xx < tx) ? __ldg(nfl.counts + (size_t)yy * tx + xx) : 0u;
  }
  if (tid < 2) s_n[tid] = 0;
  {
    uint4 *q = (uint4 *)smem;
    for (int i = tid; i < FR * FR * 4 / 16; i += FS_THREADS) q[i] = make_uint4(0, 0, 0, 0);
  }
  bool dense = false;   // (every thread sees the same counts: no shared flag needed)
#pragma unroll
  for (int t9 = 0; t9 < 9; t9++)
    if (cnt9[t9] > NFCAP) { dense = true; cnt9[t9] = 0; }
  u32 rec9[9][2];
#pragma unroll
  for (int t9 = 0; t9 < 9; t9++) {
    const size_t base = ((size_t)(by + t9 / 3 - 1) * tx + (bx + t9 % 3 - 1)) * NFCAP;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const u32 k = tid + h * FS_THREADS;
      rec9[t9][h] = k < cnt9[t9] ? ((u32)__ldg(nfl.entries + base + k) | ((u32)__ldg(nfl.eq + base + k) << 16)) : 0xFFFFFFFFu;
    }
  }
  __syncthreads();
  // 1. the region's NOFLOW cells from the records of the 3 x 3 tiles around: flags into the grid, the record (equal-
  //    elevation neighbours, high edge) beside the work list -- no elevation is read here at all
  auto put = [&](const int t9, const u32 rec) {
    const int dy = t9 / 3 - 1, dx = t9 % 3 - 1;
    const int li = (int)(rec & 0x0FFFu);
    const int rr = (li >> 6) + T * dy + FH, cc = (li & (T - 1)) + T * dx + FH;
    if (rr < 0 || rr >= FR || cc < 0 || cc >= FR) return;
    const int ri = rr * FR + cc;
    // a NOFLOW cell on the region's rim may continue outside the region: its flat is "open"
    if (rr == 0 || cc == 0 || rr == FR - 1 || cc == FR - 1) slo[ri] = (unsigned short)(F_NF | F_OPEN);
    else {
      slo[ri] = (unsigned short)F_NF;
      const u32 pos = atomicAdd(&s_n[0], 1u);
      if (pos < FLCAP) {
        list[pos] = (unsigned short)ri;
        lmask[pos] = (unsigned short)(((rec >> 16) & 0xFFu) | ((rec & HC_NF_HI) ? 0x100u : 0u));   // (record, resolved below)
      }
    }
  };
#pragma unroll
  for (int t9 = 0; t9 < 9; t9++) {
#pragma unroll
    for (int h = 0; h < 2; h++)
      if (rec9[t9][h] != 0xFFFFFFFFu) put(t9, rec9[t9][h]);
    if (cnt9[t9] > 2 * FS_THREADS) {
      const size_t base = ((size_t)(by + t9 / 3 - 1) * tx + (bx + t9 % 3 - 1)) * NFCAP;
      for (u32 k = tid + 2 * FS_THREADS; k < cnt9[t9]; k += FS_THREADS)
        put(t9, (u32)__ldg(nfl.entries + base + k) | ((u32)__ldg(nfl.eq + base + k) << 16));
    }
  }
  // ... and a NOFLOW cell of a band's halo row belongs to the neighbouring band: open too (no tile lists them)
  if (rw.top && by == 0) {
    const int rr = -1 - r0;   // region row of raster row -1
    for (int cc = tid; cc < FR; cc += FS_THREADS) {
      const int c = c0 + cc;
      if (c >= 0 && c < nx && __ldg(dirA - (long long)nx + c) == HC_DIR_NOFLOW) slo[rr * FR + cc] = (unsigned short)(F_NF | F_OPEN);
    }
  }
  if (rw.bot) {
    const int rr = ny - r0;   // region row of raster row ny
    if (rr >= 0 && rr < FR)
      for (int cc = tid; cc < FR; cc += FS_THREADS) {
        const int c = c0 + cc;
        if (c >= 0 && c < nx && __ldg(dirA + (long long)ny * nx + c) == HC_DIR_NOFLOW) slo[rr * FR + cc] = (unsigned short)(F_NF | F_OPEN);
      }
  }
  __syncthreads();
  const u32 nlist = s_n[0];
  bool giveup = nlist > FLCAP || dense;
  if (!giveup && nlist) {
    // 2. edges: a NOFLOW cell next to higher ground is a high edge (hi = 1); next to an equal cell that HAS a
    //    direction (a low edge) it starts at lo = 2; equal NOFLOW neighbours are remembered as a mask
    for (u32 k = tid; k < nlist; k += FS_THREADS) {
      const int ri = list[k];
      const u32 rec = lmask[k];
      u32 mnf = 0;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int kk = j + (j >= 4);
        if (((rec >> j) & 1u) && (slo[ri + (kk / 3 - 1) * FR + (kk % 3 - 1)] & F_NF)) mnf |= 1u << j;
      }
      const u32 mlow = (rec & 0xFFu) & ~mnf;
      lmask[k] = (unsigned short)((mlow << 8) | mnf);
      shi[ri] = (unsigned short)((rec >> 8) & 1u);
      if (mlow) slo[ri] = (unsigned short)(F_NF | 2u);
    }
    __syncthreads();
    // 3. relax the two distance fields and the "open" flag to their fixed point
    //    (FS_BATCH passes between two block-wide votes: every update is a monotone step -- distances only fall, the
    //    open flag is only ever set -- so the passes need no barrier between them, and a batch without any change is a
    //    full pass over stable values)
    int sweeps = 0;
    volatile unsigned short *vslo = slo, *vshi = shi;
    for (;;) {
      int changed = 0;
#pragma unroll 1
      for (int it = 0; it < FS_BATCH; it++)
      for (u32 k = tid; k < nlist; k += FS_THREADS) {
        const int ri = list[k];
        u32 mnf = lmask[k] & 0xFFu;
        if (!mnf) continue;
        const u32 w = vslo[ri];
        const u32 lo = w & F_LO, hi = vshi[ri];
        u32 blo = lo ? lo : 0xFFFFu, bhi = hi ? hi : 0xFFFFu;
        u32 open = w & F_OPEN;
        while (mnf) {
          const int j = __ffs(mnf) - 1;
          mnf &= mnf - 1;
          const int kk = j + (j >= 4);
          const int ni = ri + (kk / 3 - 1) * FR + (kk % 3 - 1);
          const u32 nw = vslo[ni];
          const u32 nlo = nw & F_LO;
          if (nlo && nlo + 1 < blo) blo = nlo + 1;
          const u32 nhi = vshi[ni];
          if (nhi && nhi + 1 < bhi) bhi = nhi + 1;
          open |= nw & F_OPEN;
        }
        if (blo == 0xFFFFu) blo = 0;
        if (bhi == 0xFFFFu) bhi = 0;
        const u32 nw2 = F_NF | open | blo;
        if (nw2 != w) { vslo[ri] = (unsigned short)nw2; changed = 1; }
        if (bhi != hi) { vshi[ri] = (unsigned short)bhi; changed = 1; }
      }
      sweeps += FS_BATCH;
      if (!__syncthreads_or(changed)) break;
      if (sweeps >= FSWEEPS) { giveup = true; break; }
    }
    if (tid == 0) { HC_DBG(0, sweeps); HC_DBG(1, 1); }
  }
  // 4. write: dirB already holds dirA's codes, so only the NOFLOW cells of the own tile are touched — closed flats
  //    get their direction, open ones are marked pending
  int pend = 0;
  if (!giveup) {
    for (u32 k = tid; k < nlist; k += FS_THREADS) {
      const int ri = list[k];
      const int rr = ri / FR, cc = ri - rr * FR;
      if (rr < FH || rr >= FH + T || cc < FH || cc >= FH + T) continue;   // not an own cell
      const int r = r0 + rr, c = c0 + cc;
      const u32 w = slo[ri];
      uint8_t d = HC_DIR_NOFLOW;
      if (w & F_OPEN) { d = HC_DIR_PENDING; pend = 1; }
      else if (w & F_LO) {
        const u32 mk = lmask[k];
        const int own = 2 * (int)(w & F_LO) - (int)shi[ri];
        int mn = own, best = -1;
        bool bdiag = false;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const int ks = c_k_of_j[TABLE][j];
          int key;
          if ((mk >> ks) & 1u) {
            const int ni = ri + c_dr[TABLE][j] * FR + c_dc[TABLE][j];
            const u32 nw = slo[ni];
            if (!(nw & F_LO)) continue;
            key = 2 * (int)(nw & F_LO) - (int)shi[ni];
          } else if ((mk >> (8 + ks)) & 1u) {
            key = -0x40000000;
          } else continue;
          const bool kdiag = (c_dr[TABLE][j] != 0) && (c_dc[TABLE][j] != 0);
          if (key < mn || (key == mn && best >= 0 && bdiag && !kdiag)) { mn = key; best = j; bdiag = kdiag; }
        }
        if (best >= 0) d = (uint8_t)c_code[TABLE][best];
      }
      if (d != HC_DIR_NOFLOW) dirB[(size_t)r * nx + c] = d;
    }
  } else {
    // dense / overflowing region: every NOFLOW cell of the own tile goes to the slow tier
    const size_t tile = (size_t)by * tx + bx;
    const u32 cnt = nfl.counts[tile];
    if (cnt > NFCAP) {
      for (int i = tid; i < T * T; i += FS_THREADS) {
        const int r = by * T + i / T, c = bx * T + i % T;
        if (r < ny && c < nx && dirA[(size_t)r * nx + c] == HC_DIR_NOFLOW) { dirB[(size_t)r * nx + c] = HC_DIR_PENDING; pend = 1; }
      }
    } else {
      for (u32 k = tid; k < cnt; k += FS_THREADS) {
        const int li = nfl.entries[tile * NFCAP + k] & 0x0FFFu;
        dirB[(size_t)(by * T + (li >> 6)) * nx + (bx * T + (li & (T - 1)))] = HC_DIR_PENDING;
        pend = 1;
      }
    }
  }
  pend = __syncthreads_or(pend);
  if (tid == 0 && pend) HC_DBG(2, 1);  // pending tiles
  if (tid == 0) tile_pending[(size_t)by * tx + bx] = (uint8_t)(pend ? 1 : 0);
}

// active = pending tiles and their 8 neighbours; also counts pending tiles
__global__ void k_flat_dilate(const uint8_t *__restrict__ pending, uint8_t *__restrict__ active, int tx, int ty,
                              u32 *__restrict__ n_pending, hc_rows rw) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= tx * ty) return;
  int y = t / tx, x = t - y * tx;
  int a = 0;
  for (int dy = -1; dy <= 1; dy++)
    for (int dx = -1; dx <= 1; dx++) {
      int yy = y + dy, xx = x + dx;
      if (yy >= 0 && yy < ty && xx >= 0 && xx < tx) a |= pending[yy * tx + xx];
    }
  // a band's seam tile rows always take part: the neighbouring band's flats may end on them
  if ((rw.top && y == 0) || (rw.bot && y == ty - 1)) a = 1;
  active[t] = (uint8_t)a;
  if (pending[t]) atomicAdd(n_pending, 1u);
}

// ================================================================================================
// slow tier (general: any flat size), restricted to `active` tiles
// ================================================================================================
// window form for the dense tiles (more than NFCAP NOFLOW cells: lakes): same initial state from the staged tile
__device__ __noinline__ void
flat_init_window(const float *__restrict__ z, const uint8_t *__restrict__ dir, const uint8_t *__restrict__ dirB,
                 u32 *__restrict__ dlo, u32 *__restrict__ dhi, const uint8_t *__restrict__ pending,
                 uint8_t *__restrict__ startq, int ny, int nx, float nodata, hc_rows rw, int start_all) {
  __shared__ float sz[TH * TH];
  __shared__ uint8_t sd[TH * TH];
  __shared__ uint8_t sp[TH * TH];   // dirB: the fast tier's PENDING marks, looked at on the halo ring only
  const int tile = blockIdx.y * gridDim.x + blockIdx.x;
  const int r0 = blockIdx.y * T, c0 = blockIdx.x * T;
  const int tid = threadIdx.x;
  const float qnan = __int_as_float(0x7fc00000);
  for (int i = tid; i < TH * TH; i += NTHREADS) {
    int lr = i / TH, lc = i - lr * TH;
    int r = r0 + lr - 1, c = c0 + lc - 1;
    float v = qnan;
    uint8_t d = HC_DIR_NODATA, pb = 0;
    if (r >= rw.rlo && r < rw.rhi && c >= 0 && c < nx) {
      long long g = (long long)r * nx + c;
      v = __ldg(z + g);
      if (v == nodata) v = qnan;
      d = __ldg(dir + g);
      pb = __ldcg(dirB + g);
    }
    sz[i] = v;
    sd[i] = d;
    sp[i] = pb;
  }
  __syncthreads();
  int touch = 0;
  for (int i = tid; i < T * T; i += NTHREADS) {
    int lr = i / T, lc = i - lr * T;
    int r = r0 + lr, c = c0 + lc;
    if (r >= ny || c >= nx) continue;
    int ci = (lr + 1) * TH + lc + 1;
    float v = sz[ci];
    if (!(v == v) || sd[ci] != HC_DIR_NOFLOW) continue;
    u32 lo = 0, hi = 0;
#pragma unroll
    for (int dr = -1; dr <= 1; dr++)
#pragma unroll
      for (int dc = -1; dc <= 1; dc++) {
        if (!dr && !dc) continue;
        const int hr = lr + 1 + dr, hc = lc + 1 + dc;
        const int ni = hr * TH + hc;
        const float zn = sz[ni];
        if (zn != zn) continue;
        if (zn > v) hi = 1;
        else if (zn == v) {
          if (sd[ni] != HC_DIR_NOFLOW) lo = 2;
          else if (!(hr >= 1 && hr <= T && hc >= 1 && hc <= T) && r + dr >= 0 && r + dr < ny && sp[ni] == HC_DIR_PENDING) touch = 1;
        }
      }
    const size_t g = (size_t)r * nx + c;
    dlo[g] = lo;
    dhi[g] = hi;
  }
  touch = __syncthreads_or(touch);
  if (tid == 0) {
    const bool seam = (rw.top && blockIdx.y == 0) || (rw.bot && blockIdx.y == gridDim.y - 1);
    startq[tile] = (uint8_t)((pending[tile] || touch || seam || start_all) ? 1 : 0);
  }
}

// State lives in two rasters that only ever mean something AT NOFLOW CELLS of active tiles:
//   dlo[c] = 1 + geodesic distance to the nearest low edge through the flat (a cell next to a low edge starts at 2)
//   dhi[c] = 1 + geodesic distance to the nearest high edge (a NOFLOW cell touching higher ground is 1); 0 = not reached
// A valid cell WITH a direction that touches an equal NOFLOW cell is a low edge: constant 1, never stored -- whoever
// looks at an equal-elevation neighbour asks the D8 raster (dirA) whether it is NOFLOW first.
//
// k_flat_init_lists: tiles with at most NFCAP NOFLOW cells start from their records (no elevation, no window): the
// record says which neighbours share the elevation and whether the cell is a high edge; dirA says which of those
// neighbours have a direction.  Also decides whether the tile must be relaxed before anyone asks for it: yes if it
// holds pending cells itself, or if one of its NOFLOW cells touches a PENDING cell of the same elevation in a
// neighbouring tile (it then holds part of a flat that is open over there, possibly with that flat's only sources).
__global__ void __launch_bounds__(256)
k_flat_init_lists(const float *__restrict__ z, const uint8_t *__restrict__ dirA, const uint8_t *__restrict__ dirB,
                  u32 *__restrict__ dlo, u32 *__restrict__ dhi, const uint8_t *__restrict__ active,
                  const uint8_t *__restrict__ pending, uint8_t *__restrict__ startq, const hc_nf_lists nfl, int ny, int nx,
                  float nodata, hc_rows rw, int start_all) {
  const int tile = blockIdx.y * gridDim.x + blockIdx.x;
  if (!active[tile]) { if (threadIdx.x == 0) startq[tile] = 0; return; }
  const u32 cnt = nfl.counts[tile];
  if (cnt > NFCAP) {   // a dense tile (lake): the window form
    flat_init_window(z, dirA, dirB, dlo, dhi, pending, startq, ny, nx, nodata, rw, start_all);
    return;
  }
  const int r0 = blockIdx.y * T, c0 = blockIdx.x * T;
  int touch = 0;
  for (u32 k = threadIdx.x; k < cnt; k += 256) {
    const u32 e = nfl.entries[(size_t)tile * NFCAP + k];
    const u32 eq = nfl.eq[(size_t)tile * NFCAP + k];
    const int lr = (int)((e & 0x0FFFu) >> 6), lc = (int)(e & 63u);
    const long long g = (long long)(r0 + lr) * nx + (c0 + lc);
    u32 lo = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (!((eq >> j) & 1u)) continue;
      const int kk = j + (j >= 4);
      const int dr = kk / 3 - 1, dc = kk % 3 - 1;
      const long long gn = g + (long long)dr * nx + dc;
      if (__ldg(dirA + gn) != HC_DIR_NOFLOW) lo = 2;
      else {
        const int nr = lr + dr, nc = lc + dc;
        if ((nr < 0 || nr >= T || nc < 0 || nc >= T) && r0 + nr >= 0 && r0 + nr < ny && __ldcg(dirB + gn) == HC_DIR_PENDING) touch = 1;
      }
    }
    dlo[g] = lo;
    dhi[g] = (e & HC_NF_HI) ? 1u : 0u;
  }
  touch = __syncthreads_or(touch);
  if (threadIdx.x == 0) {
    // a band cannot see the neighbouring band's pending marks: its seam tile rows always start
    const bool seam = (rw.top && blockIdx.y == 0) || (rw.bot && blockIdx.y == gridDim.y - 1);
    startq[tile] = (uint8_t)((pending[tile] || touch || seam || start_all) ? 1 : 0);
  }
}

// Slow-tier visits.  Shared-memory window of a lake tile: z f32 | lo u32 | hi u32 | mask u16, pitch TP (odd).  Inside a
// visit "not reached" is R_INF instead of 0, so that a distance can be lowered with one shared atomicMin.
#define TP (TH + 1)
#define R_INF 0x3FFFFFFFu
#define M_NF 0x100u    // mask bits 0..7: equal-elevation neighbour k of the 3x3 window (row-major, centre skipped)
#define M_CHG 0x200u   // set when this visit lowered the cell (write-back then needs no global re-read)
#define FLAT_SMEM (TH * TP * 14 + 16)

// One directional sweep of one distance field over the tile window (one warp, two lines per lane).  Direction D walks
// the 64 lines of the tile (0: rows downwards, 1: rows upwards, 2: columns rightwards, 3: columns leftwards) and lowers
// every cell of a line from its three neighbours on the PREVIOUS line, whose values the warp still holds in registers
// (neighbouring lanes by shuffle, the two halo ends from shared memory).  The four directions together cover all eight
// neighbours, so an iteration of all four that lowers nothing proves the fixed point; on an obstacle-free tile a
// geodesic between two cells is a chessboard path that stays inside one direction's cone, so ONE iteration already
// carries every halo value to every cell (the old whole-line Gauss-Seidel needed the same iterations at 27 shared
// loads per cell and step).  Concurrent sweeps only ever lower a value (atomicMin), so none is lost.
template <int D>
__device__ __forceinline__ int cone_sweep(u32 *val, unsigned short *smk, const int lane) {
  constexpr bool vert = D < 2;
  constexpr int step = (D == 0 || D == 2) ? 1 : -1;
  constexpr int bA = D == 0 ? 0 : (D == 1 ? 5 : (D == 2 ? 0 : 2));
  constexpr int bB = D == 0 ? 1 : (D == 1 ? 6 : (D == 2 ? 3 : 4));
  constexpr int bC = D == 0 ? 2 : (D == 1 ? 7 : (D == 2 ? 5 : 7));
  const int j0 = 1 + 2 * lane, j1 = j0 + 1;
  volatile u32 *v = val;
  volatile unsigned short *mk = smk;
#define CS_IDX(s_, j_) (vert ? (s_) * TP + (j_) : (j_) * TP + (s_))
  int s = step > 0 ? 1 : T;
  u32 p0 = v[CS_IDX(s - step, j0)], p1 = v[CS_IDX(s - step, j1)];
  int changed = 0;
  // the line's own values, masks and the previous line's two halo ends are fetched one step ahead, so that their
  // shared-memory latency hides behind the shuffle / min chain of the step before (a value another sweep lowers in
  // between is only seen a step late, which costs at most one more iteration)
  u32 n_c0 = v[CS_IDX(s, j0)], n_c1 = v[CS_IDX(s, j1)], n_m0 = mk[CS_IDX(s, j0)], n_m1 = mk[CS_IDX(s, j1)];
  u32 n_e0 = v[CS_IDX(s - step, 0)], n_e1 = v[CS_IDX(s - step, T + 1)];
  for (int k = 0; k < T; k++, s += step) {
    const int i0 = CS_IDX(s, j0), i1 = CS_IDX(s, j1);
    u32 c0 = n_c0, c1 = n_c1;
    const u32 m0 = n_m0, m1 = n_m1, e0 = n_e0, e1 = n_e1;
    if (k + 1 < T) {
      const int sn = s + step;
      n_c0 = v[CS_IDX(sn, j0)]; n_c1 = v[CS_IDX(sn, j1)];
      n_m0 = mk[CS_IDX(sn, j0)]; n_m1 = mk[CS_IDX(sn, j1)];
      n_e0 = v[CS_IDX(s, 0)]; n_e1 = v[CS_IDX(s, T + 1)];
    }
    u32 lft = __shfl_up_sync(0xFFFFFFFFu, p1, 1), rgt = __shfl_down_sync(0xFFFFFFFFu, p0, 1);
    if (lane == 0) lft = e0;
    if (lane == 31) rgt = e1;
    const u32 a = min(min(((m0 >> bA) & 1u) ? lft : R_INF, ((m0 >> bB) & 1u) ? p0 : R_INF), ((m0 >> bC) & 1u) ? p1 : R_INF) + 1u;
    const u32 b = min(min(((m1 >> bA) & 1u) ? p0 : R_INF, ((m1 >> bB) & 1u) ? p1 : R_INF), ((m1 >> bC) & 1u) ? rgt : R_INF) + 1u;
    if ((m0 & M_NF) && a < c0) { atomicMin(&val[i0], a); mk[i0] = (unsigned short)(m0 | M_CHG); c0 = a; changed = 1; }
    if ((m1 & M_NF) && b < c1) { atomicMin(&val[i1], b); mk[i1] = (unsigned short)(m1 | M_CHG); c1 = b; changed = 1; }
    p0 = c0;
    p1 = c1;
  }
#undef CS_IDX
  return changed;
}

// one visit of a lake tile (by, bx): load the window, relax to the local fixed point, write back, queue the neighbours
// that can improve
__device__ __forceinline__ void relax_tile(const float *__restrict__ z, const uint8_t *__restrict__ dirA, u32 *dlo, u32 *dhi,
                                           const uint8_t *__restrict__ active, u32 *__restrict__ list_out,
                                           u32 *__restrict__ n_out, u32 *__restrict__ queued, const u32 stamp,
                                           uint8_t *__restrict__ visited, int ny, int nx, float nodata,
                                           unsigned long long *dbg, hc_rows rw, int by, int bx, int tx, int ty,
                                           unsigned char *smem) {
  float *sz = (float *)smem;
  u32 *slo = (u32 *)(smem + TH * TP * 4);
  u32 *shi = (u32 *)(smem + TH * TP * 8);
  unsigned short *smk = (unsigned short *)(smem + TH * TP * 12);
  __shared__ int s_side;
  const int r0 = by * T, c0 = bx * T;
  const int tid = threadIdx.x;
  const float qnan = __int_as_float(0x7fc00000);
  __shared__ uint8_t s_act[9], s_vis[9];
  __syncthreads();  // the previous visit of this CTA is completely done with the shared arrays
  const long long t_begin = clock64();
  if (tid == 0) s_side = 0;
  if (tid < 9) {
    int yy = by + tid / 3 - 1, xx = bx + tid % 3 - 1;
    bool in = yy >= 0 && yy < ty && xx >= 0 && xx < tx;
    s_act[tid] = in ? active[yy * tx + xx] : 0;
    s_vis[tid] = in ? __ldcg(visited + yy * tx + xx) : 1;   // L2: a persistent CTA's L1 may hold a stale 0
  }
  __syncthreads();
  // first visit of this tile: it must wake the not-yet-visited neighbours that hold more of its flats (they may
  // own the flat's only sources, and nothing else would ever ask for them)
  const bool first = s_vis[4] == 0;
  if (tid == 0) visited[by * tx + bx] = 1;
#pragma unroll 4
  for (int i = tid; i < TH * TH; i += NTHREADS) {
    const int lr = i / TH, lc = i - lr * TH;
    const int r = r0 + lr - 1, c = c0 + lc - 1;
    const int ay = lr == 0 ? 0 : (lr == TH - 1 ? 2 : 1), ax = lc == 0 ? 0 : (lc == TH - 1 ? 2 : 1);
    // (a band's halo rows hold the neighbouring band's state, refreshed between relaxations)
    const bool halo_row = (r == -1 && rw.top) || (r == ny && rw.bot);
    const bool inr = c >= 0 && c < nx && ((r >= 0 && r < ny) || halo_row);
    const long long g = inr ? (long long)r * nx + c : 0;
    const float v = __ldg(z + g);
    const bool nfc = __ldg(dirA + g) == HC_DIR_NOFLOW;
    // NOFLOW cells of tiles outside the active set hold no state: not in any flat (they cannot share a flat with a pending
    // cell); a cell WITH a direction needs no state -- it only matters as the low edge of an equal NOFLOW neighbour: constant 1
    const bool f = inr && nfc && (halo_row || s_act[ay * 3 + ax]);
    const bool inb = inr && (f || !nfc);
    u32 lo = 1, hi = 0;
    if (f) { lo = __ldcg(dlo + g); hi = __ldcg(dhi + g); }
    sz[lr * TP + lc] = (inb && v != nodata) ? v : qnan;
    slo[lr * TP + lc] = lo ? lo : R_INF;
    shi[lr * TP + lc] = hi ? hi : R_INF;
    smk[lr * TP + lc] = (unsigned short)(f ? M_NF : 0u);
  }
  __syncthreads();
  // equal-elevation links of the own cells (NaN never compares equal: nodata and absent cells link to nothing)
  for (int i = tid; i < T * T; i += NTHREADS) {
    const int ci = ((i >> 6) + 1) * TP + (i & (T - 1)) + 1;
    const float v = sz[ci];
    u32 m = 0;
    int k = 0;
#pragma unroll
    for (int dr = -1; dr <= 1; dr++)
#pragma unroll
      for (int dc = -1; dc <= 1; dc++) {
        if (!dr && !dc) continue;
        if (sz[ci + dr * TP + dc] == v) m |= 1u << k;
        k++;
      }
    smk[ci] = (unsigned short)(smk[ci] | m);
  }
  __syncthreads();
  // warp w: field w & 1 (lo / hi), direction w >> 1
  const int warp = tid >> 5, lane = tid & 31;
  u32 *val = (warp & 1) ? shi : slo;
  int ever = 0, nsweeps = 0;
  for (;;) {
    int changed;
    switch (warp >> 1) {
      case 0: changed = cone_sweep<0>(val, smk, lane); break;
      case 1: changed = cone_sweep<1>(val, smk, lane); break;
      case 2: changed = cone_sweep<2>(val, smk, lane); break;
      default: changed = cone_sweep<3>(val, smk, lane); break;
    }
    nsweeps += 4;
    if (!__syncthreads_or(changed)) break;
    ever = 1;
  }
  if (tid == 0) { HC_DBG(3, nsweeps); HC_DBG(4, 1); HC_DBG(5, 1); HC_DBG(10, clock64() - t_begin); HC_DBG(11, nsweeps / 4); }
  if (!ever && !first) return;
  // write back; a neighbouring tile is re-queued only if one of ITS rim cells (my halo) can actually improve
  // from a rim cell of mine that changed (distances only ever decrease, so a stale halo value can cause a
  // spurious visit but never a missed one).  Without this test a tile is revisited every time the front
  // moves on in any neighbour.
  int nb9 = 0;  // bit (dy+1)*3 + (dx+1)
  for (int i = tid; i < T * T; i += NTHREADS) {
    const int lr = i / T, lc = i - lr * T;
    const int r = r0 + lr, c = c0 + lc;
    if (r >= ny || c >= nx) continue;
    const int ci = (lr + 1) * TP + lc + 1;
    const u32 m = smk[ci];
    if (!(m & M_NF)) continue;
    const bool chg = (m & M_CHG) != 0;
    if (!chg && !first) continue;
    const size_t g = (size_t)r * nx + c;
    const u32 lo = slo[ci], hi = shi[ci];
    if (chg) {
      dlo[g] = lo == R_INF ? 0u : lo;
      dhi[g] = hi == R_INF ? 0u : hi;
    }
    if (lr == 0 || lc == 0 || lr == T - 1 || lc == T - 1) {
      int k = 0;
#pragma unroll
      for (int dr = -1; dr <= 1; dr++)
#pragma unroll
        for (int dc = -1; dc <= 1; dc++) {
          if (!dr && !dc) continue;
          const int kk = k++;
          const int hr = lr + 1 + dr, hc = lc + 1 + dc;
          if (!(hr == 0 || hr == TH - 1 || hc == 0 || hc == TH - 1)) continue;  // not a halo cell
          if (!((m >> kk) & 1u)) continue;                                      // not the same elevation
          const int ni = hr * TP + hc;
          if (!(smk[ni] & M_NF)) continue;
          const u32 nlo = slo[ni], nhi = shi[ni];
          const int nb = (hr == 0 ? 0 : (hr == TH - 1 ? 2 : 1)) * 3 + (hc == 0 ? 0 : (hc == TH - 1 ? 2 : 1));
          if ((chg && (lo + 1u < nlo || hi + 1u < nhi)) || (first && !s_vis[nb])) nb9 |= 1 << nb;
        }
    }
  }
  if (nb9) atomicOr(&s_side, nb9);
  __syncthreads();
  nb9 = s_side;
  if (tid < 9 && tid != 4 && (nb9 & (1 << tid))) {
    int yy = by + tid / 3 - 1, xx = bx + tid % 3 - 1;
    if (yy >= 0 && yy < ty && xx >= 0 && xx < tx && active[yy * tx + xx]) {
      // the first one to flag a tile appends it to the next round's queue
      const u32 t = (u32)(yy * tx + xx);
      if (atomicExch(&queued[t], stamp) != stamp) list_out[atomicAdd(n_out, 1u)] = t;
    }
  }
}

// Light visit for tiles with at most FL_LCAP NOFLOW cells (most visits): no window.  The tile's compact cell list and
// neighbour masks were built at init; the visit copies those cells' two distances into shared memory, resolves every
// in-tile neighbour to its list position through a 64 x 64 position grid (an equal-elevation neighbour that is not
// listed has a direction, i.e. it is a low edge: constant 1), folds the neighbours in other tiles -- fixed for the
// visit -- into one constant per cell, and then runs the Jacobi sweeps entirely on shared memory (the first version
// swept on the L2-resident rasters: eight dependent L2 round trips per cell and sweep).  Same re-queue / wake rules as
// the window path.
#define FL_PER (FL_LCAP / NTHREADS)
#ifndef HC_LIGHT_ASYNC
#define HC_LIGHT_ASYNC 1
#endif
#define FL_BATCH 3            // relaxation passes between two block-wide votes (light visits)
#define FL_RING (4 * T + 4)   // ring cells: row -1 (T + 2, corners included), row T (T + 2), column -1 (T), column T (T)
#define RING_IDX(nr_, nc_) ((nr_) < 0 ? (nc_) + 1 : ((nr_) >= T ? (T + 2) + (nc_) + 1 : ((nc_) < 0 ? 2 * (T + 2) + (nr_) : 2 * (T + 2) + T + (nr_))))
#define RING_TILE(nr_, nc_) (((nr_) < 0 ? 0 : ((nr_) >= T ? 2 : 1)) * 3 + ((nc_) < 0 ? 0 : ((nc_) >= T ? 2 : 1)))
__device__ __forceinline__ void relax_light(const uint8_t *__restrict__ dirA, u32 *dlo, u32 *dhi, const uint8_t *__restrict__ active,
                                            const unsigned short *__restrict__ tlist, const uint8_t *__restrict__ tmask,
                                            const u32 cnt, u32 *__restrict__ list_out, u32 *__restrict__ n_out,
                                            u32 *__restrict__ queued, const u32 stamp, uint8_t *visited, int ny, int nx, int tx,
                                            int ty, unsigned long long *dbg, const u32 t, const hc_rows rw, unsigned char *smem) {
  unsigned short *grid = (unsigned short *)smem;        // [T * T] list position + 1 of a NOFLOW cell, else 0
  volatile u32 *vlo = (volatile u32 *)(smem + T * T * 2);   // [FL_LCAP]
  volatile u32 *vhi = vlo + FL_LCAP;
  // snapshot of the one-cell ring around the tile (4 x 64 + 4 corners; RING_IDX), taken once at the start of the visit
  // with one load per thread: 0 = no cell there, 1 = a cell with a direction, 2 = a NOFLOW cell (+ its two distances)
  u32 *h_lo = (u32 *)(smem + T * T * 2 + FL_LCAP * 8);
  u32 *h_hi = h_lo + FL_RING;
  uint8_t *h_nf = (uint8_t *)(h_hi + FL_RING);
  __shared__ u32 s_want;
  __shared__ uint8_t s_lact[9], s_lvis[9];   // the 3 x 3 tiles around: takes part in the slow tier / has been visited
  const int tid = threadIdx.x;
  const int by = (int)(t / tx), bx = (int)(t % tx);
  const int r0 = by * T, c0 = bx * T;
  __syncthreads();  // the previous visit of this CTA is completely done with the shared arrays
  const long long t_begin = clock64();
  // every global read of the visit that does not depend on another one is issued here, side by side: the cell list,
  // the ring, the neighbour tiles' flags (chains of dependent reads per rim cell were most of a visit's time)
  for (int i = tid; i < FL_RING; i += NTHREADS) {
    int nr, nc;
    if (i < 2 * (T + 2)) { nr = i < T + 2 ? -1 : T; nc = (i < T + 2 ? i : i - (T + 2)) - 1; }
    else { const int k = i - 2 * (T + 2); nc = k < T ? -1 : T; nr = k < T ? k : k - T; }
    const int r = r0 + nr, c = c0 + nc;
    u32 f = 0, a = 0, b = 0;
    if (c >= 0 && c < nx && r >= rw.rlo && r < rw.rhi) {
      const long long g = (long long)r * nx + c;
      f = __ldg(dirA + g) == HC_DIR_NOFLOW ? 2u : 1u;
      a = __ldcg(dlo + g);    // (meaningless unless f == 2 and the cell's tile takes part: looked at under those tests only)
      b = __ldcg(dhi + g);
    }
    h_nf[i] = (uint8_t)f; h_lo[i] = a; h_hi[i] = b;
  }
  if (tid < 9) {
    const int yy = by + tid / 3 - 1, xx = bx + tid % 3 - 1;
    uint8_t a = 0, v = 1;
    if (xx >= 0 && xx < tx) {
      if (yy >= 0 && yy < ty) { a = active[yy * tx + xx]; v = __ldcg(visited + yy * tx + xx); }   // (L2: a persistent CTA's L1 may hold a stale 0)
      else a = (yy < 0 ? rw.top : rw.bot) ? 1 : 0;   // a band's halo row: the exchange maintains its state
    }
    s_lact[tid] = a; s_lvis[tid] = v;
  }
  {
    uint4 *q = (uint4 *)grid;
    for (int i = tid; i < T * T * 2 / 16; i += NTHREADS) q[i] = make_uint4(0, 0, 0, 0);
  }
  if (tid == 0) s_want = 0;
  __syncthreads();
  const bool first = s_lvis[4] == 0;
  long long g[FL_PER];
  u32 lo[FL_PER], hi[FL_PER], msk[FL_PER], ref[FL_PER][4];
#pragma unroll
  for (int j = 0; j < FL_PER; j++) {
    const u32 k = tid + j * NTHREADS;
    msk[j] = 0; g[j] = 0; lo[j] = hi[j] = R_INF;
    if (k < cnt) {
      const u32 li = tlist[(size_t)t * FL_LCAP + k] & 0x0FFFu;
      msk[j] = tmask[(size_t)t * FL_LCAP + k] | 0x100u | (li << 16);   // bit 8 = "this slot holds a cell"
      g[j] = (long long)(r0 + (int)(li >> 6)) * nx + (c0 + (int)(li & 63u));
      const u32 a = __ldcg(dlo + g[j]), b = __ldcg(dhi + g[j]);
      lo[j] = a ? a : R_INF;
      hi[j] = b ? b : R_INF;
      grid[li] = (unsigned short)(k + 1);
      vlo[k] = lo[j];
      vhi[k] = hi[j];
    }
  }
  __syncthreads();
  const long long t_l1 = clock64();
  // neighbours: listed cells of the tile go to a packed queue of list positions (eight 16-bit slots, first free slot
  // = 0xFFFF); everything else is constant during the visit and is folded into the cell's own value right away
  u32 chg = 0;
#pragma unroll
  for (int j = 0; j < FL_PER; j++) {
    ref[j][0] = ref[j][1] = ref[j][2] = ref[j][3] = 0xFFFFFFFFu;
    if (!(msk[j] & 0x100u)) continue;
    u32 elo = R_INF, ehi = R_INF;
    const int lr = (int)(msk[j] >> 22), lc = (int)((msk[j] >> 16) & 63u);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (!((msk[j] >> k) & 1u)) continue;
      const int kk = k + (k >= 4);                               // position in the 3x3 window
      const int nr = lr + kk / 3 - 1, nc = lc + kk % 3 - 1;
      // (a band's halo row can lie inside the tile's footprint: that cell belongs to the other band, it is not listed)
      if (nr >= 0 && nr < T && nc >= 0 && nc < T && r0 + nr < ny) {
        const u32 pos = grid[nr * T + nc];
        if (pos) {
          ref[j][3] = __funnelshift_l(ref[j][2], ref[j][3], 16);
          ref[j][2] = __funnelshift_l(ref[j][1], ref[j][2], 16);
          ref[j][1] = __funnelshift_l(ref[j][0], ref[j][1], 16);
          ref[j][0] = (ref[j][0] << 16) | (pos - 1u);
        } else elo = min(elo, 2u);                               // a low edge (lo = 1) of the same elevation
      } else if (nr >= 0 && nr < T && nc >= 0 && nc < T) {
        // a band's halo row inside the tile's footprint (ny not a multiple of the tile size): not on the ring
        const long long gn = g[j] + (long long)(kk / 3 - 1) * nx + (kk % 3 - 1);
        if (__ldg(dirA + gn) != HC_DIR_NOFLOW) { elo = min(elo, 2u); continue; }
        const u32 a = __ldcg(dlo + gn), b = __ldcg(dhi + gn);
        if (a) elo = min(elo, a + 1u);
        if (b) ehi = min(ehi, b + 1u);
      } else {
        const int ri = RING_IDX(nr, nc);
        const u32 f = h_nf[ri];
        if (f == 1u) { elo = min(elo, 2u); continue; }   // a low edge of the same elevation (needs no state)
        // a NOFLOW cell of a tile outside the active set holds no state (its rasters were never initialised): not in
        // any flat, as in the window path
        if (f != 2u || !s_lact[RING_TILE(nr, nc)]) continue;
        const u32 a = h_lo[ri], b = h_hi[ri];
        if (a) elo = min(elo, a + 1u);
        if (b) ehi = min(ehi, b + 1u);
      }
    }
    const u32 k = tid + j * NTHREADS;
    if (elo < lo[j]) { lo[j] = elo; vlo[k] = elo; chg |= 1u << j; }
    if (ehi < hi[j]) { hi[j] = ehi; vhi[k] = ehi; chg |= 1u << j; }
    if ((ref[j][0] & 0xFFFFu) == 0xFFFFu) msk[j] |= 0x200u;     // bit 9: no listed neighbour, nothing left to relax
  }
  int nsweeps = 0;
  const long long t_l2 = clock64();
  __syncthreads();
#if HC_LIGHT_ASYNC
  // Asynchronous pull relaxation: every thread lowers ITS cells from their listed neighbours, FL_BATCH times in a row
  // without any barrier in between -- chaotic iteration of a monotone operator converges to the same fixed point in any
  // interleaving, so warps need not wait for each other -- and one block-wide vote per batch decides whether anything
  // still moved.  A batch in which no thread changed anything is a full pass over stable values: the fixed point.  (The
  // barrier-per-sweep forms, pull or push, spent half of a visit waiting at ~14 block barriers with 7 of 8 warps idle.)
  for (;;) {
    int changed = 0;
#pragma unroll 1
    for (int it = 0; it < FL_BATCH; it++) {
#pragma unroll
      for (int j = 0; j < FL_PER; j++) {
        if ((msk[j] & 0x300u) != 0x100u) continue;
        u32 bl = lo[j], bh = hi[j];
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const u32 rk = (ref[j][i >> 1] >> (16 * (i & 1))) & 0xFFFFu;
          if (rk == 0xFFFFu) break;   // (slots fill from the low end)
          bl = min(bl, vlo[rk] + 1u);
          bh = min(bh, vhi[rk] + 1u);
        }
        const u32 k = tid + j * NTHREADS;
        if (bl < lo[j]) { lo[j] = bl; vlo[k] = bl; chg |= 1u << j; changed = 1; }
        if (bh < hi[j]) { hi[j] = bh; vhi[k] = bh; chg |= 1u << j; changed = 1; }
      }
    }
    nsweeps += FL_BATCH;
    if (!__syncthreads_or(changed)) break;
  }
#else
  // Push relaxation: a cell whose value dropped offers value + 1 to its listed neighbours (shared atomicMin, only when it
  // lowers them), so a sweep costs work only where the front is; the first sweep offers every reached cell once.  (The
  // pull form re-read all eight neighbours of every cell in every sweep: 16 shared loads per cell and sweep.)  A sweep
  // that lowers nothing ends the visit: cells that did not offer have not changed since their last offer, and their
  // neighbours only ever decrease.
  u32 push = 0;
#pragma unroll
  for (int j = 0; j < FL_PER; j++)
    if ((msk[j] & 0x300u) == 0x100u && (lo[j] != R_INF || hi[j] != R_INF)) push |= 1u << j;
  for (;;) {
    int changed = 0;
#pragma unroll
    for (int j = 0; j < FL_PER; j++) {
      if (!((push >> j) & 1u)) continue;
      const u32 nl = lo[j] + 1u, nh = hi[j] + 1u;
      const u32 own = tid + j * NTHREADS;
      // all sixteen loads are issued before any is looked at (absent slots read the cell's own slot)
      u32 rk[8], cl[8], ch[8];
#pragma unroll
      for (int i = 0; i < 8; i++) {
        rk[i] = (ref[j][i >> 1] >> (16 * (i & 1))) & 0xFFFFu;
        const u32 at = rk[i] == 0xFFFFu ? own : rk[i];
        cl[i] = vlo[at];
        ch[i] = vhi[at];
      }
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (rk[i] == 0xFFFFu) continue;
        if (cl[i] > nl) { atomicMin((u32 *)&vlo[rk[i]], nl); changed = 1; }
        if (ch[i] > nh) { atomicMin((u32 *)&vhi[rk[i]], nh); changed = 1; }
      }
    }
    nsweeps++;
    if (!__syncthreads_or(changed)) break;
    push = 0;
#pragma unroll
    for (int j = 0; j < FL_PER; j++) {
      if (!(msk[j] & 0x100u)) continue;
      const u32 k = tid + j * NTHREADS;
      const u32 a = vlo[k], b = vhi[k];
      if (a < lo[j] || b < hi[j]) {
        lo[j] = a; hi[j] = b;
        chg |= 1u << j;
        if (!(msk[j] & 0x200u)) push |= 1u << j;
      }
    }
  }
#endif
  const long long t_l3 = clock64();
  if (tid == 0) { HC_DBG(3, nsweeps); HC_DBG(4, 1); HC_DBG(6, 1); HC_DBG(12, t_l1 - t_begin); HC_DBG(13, t_l2 - t_l1); HC_DBG(14, t_l3 - t_l2); HC_DBG(15, cnt); }
#pragma unroll
  for (int j = 0; j < FL_PER; j++)
    if ((chg >> j) & 1u) {
      __stcg(dlo + g[j], lo[j] == R_INF ? 0u : lo[j]);
      __stcg(dhi + g[j], hi[j] == R_INF ? 0u : hi[j]);
    }
  // re-queue neighbours that can improve from a changed rim cell; on the first visit also wake the unvisited
  // neighbours that hold more of this tile's flats (votes are collected in shared memory, one thread per neighbour
  // tile then does the global queue operation)
  u32 want = 0;
#pragma unroll
  for (int j = 0; j < FL_PER; j++) {
    if (!(msk[j] & 0x100u)) continue;
    const bool c = (chg >> j) & 1u;
    if (!c && !first) continue;
    const int lr = (int)(msk[j] >> 22), lc = (int)((msk[j] >> 16) & 63u);
    if (lr != 0 && lc != 0 && lr != T - 1 && lc != T - 1) continue;
    u32 m = msk[j] & 0xFFu;
    while (m) {
      const int k = __ffs(m) - 1;
      m &= m - 1;
      const int kk = k + (k >= 4);
      const int nr = lr + kk / 3 - 1, nc = lc + kk % 3 - 1;
      if (nr >= 0 && nr < T && nc >= 0 && nc < T) continue;   // inside this tile
      const int dy = nr < 0 ? -1 : (nr >= T ? 1 : 0), dx = nc < 0 ? -1 : (nc >= T ? 1 : 0);
      const int yy = by + dy, xx = bx + dx;
      if (yy < 0 || yy >= ty || xx < 0 || xx >= tx) continue;  // a band's halo row: the exchange handles it
      const int t9 = (dy + 1) * 3 + dx + 1;
      if (!s_lact[t9]) continue;
      // (the ring is the snapshot from the start of the visit: distances only ever decrease, so a stale value can cause a
      // spurious visit, never a missed one -- as in the window path)
      const int ri = RING_IDX(nr, nc);
      if (h_nf[ri] != 2u) continue;
      const u32 a = h_lo[ri], b = h_hi[ri];
      const u32 nlo = a ? a : R_INF, nhi = b ? b : R_INF;
      const bool benefit = c && (lo[j] + 1u < nlo || hi[j] + 1u < nhi);
      if (benefit || (first && !s_lvis[t9])) want |= 1u << t9;
    }
  }
  if (want) atomicOr(&s_want, want);
  __syncthreads();
  if (tid < 9 && ((s_want >> tid) & 1u)) {
    const u32 tt = (u32)((by + tid / 3 - 1) * tx + bx + tid % 3 - 1);
    if (atomicExch(&queued[tt], stamp) != stamp) list_out[atomicAdd(n_out, 1u)] = tt;
  }
  if (tid == 0) { visited[t] = 1; HC_DBG(9, clock64() - t_begin); }
}

// The slow tier's round loop, on the device: persistent CTAs (cooperative launch) walk the queue of dirty tiles with
// dynamic fetch (tile costs differ by orders of magnitude), a grid-wide barrier ends the round, the queue the visits
// filled becomes the next round's input, until a round queues nothing.  ctl: [0] n(qA) [1] n(qB) [2..3] fetch counters.
// A tile is appended to the next queue by the first visit that stamps it with the round's number (`queued` holds 0 or 1
// on entry -- 1 for tiles the caller queued -- and rounds stamp from 2 on, so no per-round clearing is needed).
__global__ void __launch_bounds__(NTHREADS, 3)
k_flat_relax(const float *__restrict__ z, const uint8_t *__restrict__ dirA, u32 *dlo, u32 *dhi,
             const uint8_t *__restrict__ active, u32 *qA, u32 *qB, u32 *ctl, u32 *__restrict__ queued,
             uint8_t *visited, const unsigned short *__restrict__ tlist, const uint8_t *__restrict__ tmask,
             const u32 *__restrict__ tcount, int ny, int nx, float nodata, unsigned long long *dbg, hc_rows rw, int tx, int ty) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ u32 s_q;
  cg::grid_group grid = cg::this_grid();
  u32 *lin = qA, *lout = qB;
  for (u32 it = 0;; it++) {
    // three rotating counter slots: this round's input count, the count the visits build for the next round, and the
    // slot after that, which nobody touches during this round and which is cleared here for the round after
    volatile u32 *nin = ctl + it % 3u;
    u32 *nout = ctl + (it + 1u) % 3u, *fetch = ctl + 3 + it % 3u;
    if (blockIdx.x == 0 && threadIdx.x == 0) { ctl[(it + 2u) % 3u] = 0u; ctl[3 + (it + 2u) % 3u] = 0u; }
    const u32 n = *nin;
    if (blockIdx.x == 0 && threadIdx.x == 0 && it < 23u) {   // trace (HC_TRACE): tiles and start time of every round
      unsigned long long ns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
      ctl[8 + 2 * it] = n;
      ctl[9 + 2 * it] = (u32)ns;
      if (it >= 14u && n == 1u) { const u32 t0 = __ldcg(lin); ctl[9 + 2 * it] = t0 | (tcount[t0] > FL_LCAP ? 0x80000000u : 0u); }
    }
    if (!n) break;
    if (it >= (1u << 16)) {   // cannot happen on a consistent state (distances only decrease); never spin forever on a broken one
      if (blockIdx.x == 0 && threadIdx.x == 0) { HC_DBG(7, 1); *nin = 0u; }
      break;
    }
    const u32 stamp = it + 2u;
    for (;;) {
      __syncthreads();
      if (threadIdx.x == 0) s_q = atomicAdd(fetch, 1u);
      __syncthreads();
      const u32 q = s_q;
      if (q >= n) break;
      const u32 t = __ldcg(lin + q);
      const u32 cnt = tcount[t];
      if (!cnt) continue;
      if (cnt <= FL_LCAP)
        relax_light(dirA, dlo, dhi, active, tlist, tmask, cnt, lout, nout, queued, stamp, visited, ny, nx, tx, ty, dbg, t, rw, smem);
      else
        relax_tile(z, dirA, dlo, dhi, active, lout, nout, queued, stamp, visited, ny, nx, nodata, dbg, rw, (int)(t / tx), (int)(t % tx), tx, ty, smem);
    }
    grid.sync();
    u32 *sw = lin; lin = lout; lout = sw;
  }
}

// queue builders: every active tile (first round) / the active tiles of one tile row (after a halo row changed)
__global__ void k_flat_queue_tiles(const uint8_t *__restrict__ active, u32 first, u32 count, u32 *queued, u32 *list, u32 *n) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  u32 t = first + i;
  if (active[t] && atomicExch(&queued[t], 1u) == 0u) list[atomicAdd(n, 1u)] = t;
}

// masked D8 for the pending cells (window form: the dense tiles)
template <int TABLE>
__device__ __noinline__ void
flat_dirs_window(const float *__restrict__ z, const uint8_t *__restrict__ dirA, const u32 *__restrict__ dlo,
                 const u32 *__restrict__ dhi, uint8_t *__restrict__ dirB, int ny, int nx, float nodata, hc_rows rw) {
  const int r0 = blockIdx.y * T, c0 = blockIdx.x * T;
  for (int i = threadIdx.x; i < T * T; i += 256) {
    int lr = i / T, lc = i - lr * T;
    int r = r0 + lr, c = c0 + lc;
    if (r >= ny || c >= nx) continue;
    size_t g = (size_t)r * nx + c;
    if (dirB[g] != HC_DIR_PENDING) continue;
    uint8_t out = HC_DIR_NOFLOW;
    u32 lo = dlo[g];
    if (lo) {
      float v = z[g];
      long long own = 2ll * (long long)lo - (long long)dhi[g];
      long long mn = own;
      int best = -1;
      bool bdiag = false;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        int rr = r + c_dr[TABLE][k], cc = c + c_dc[TABLE][k];
        if (rr < rw.rlo || rr >= rw.rhi || cc < 0 || cc >= nx) continue;
        long long gn = (long long)rr * nx + cc;
        float zn = z[gn];
        if (zn != v) continue;  // also drops nodata / NaN
        long long key;
        if (dirA[gn] == HC_DIR_NOFLOW) {
          u32 nlo = dlo[gn];
          if (!nlo) continue;
          key = 2ll * (long long)nlo - (long long)dhi[gn];
        } else {
          key = -0x4000000000000000ll;  // low edge: below every mask value
        }
        bool kdiag = (c_dr[TABLE][k] != 0) && (c_dc[TABLE][k] != 0);
        if (key < mn || (key == mn && best >= 0 && bdiag && !kdiag)) { mn = key; best = k; bdiag = kdiag; }
      }
      if (best >= 0) out = (uint8_t)c_code[TABLE][best];
    }
    dirB[g] = out;
  }
}

// masked D8 for the pending cells of the tiles that have records: only the NOFLOW cells are visited, and of their
// neighbours only the equal-elevation ones the record names (no elevation is read)
template <int TABLE>
__global__ void __launch_bounds__(256)
k_flat_dirs_lists(const float *__restrict__ z, const uint8_t *__restrict__ dirA, const u32 *__restrict__ dlo,
                  const u32 *__restrict__ dhi, uint8_t *__restrict__ dirB, const uint8_t *__restrict__ pending,
                  const hc_nf_lists nfl, int ny, int nx, float nodata, hc_rows rw) {
  const int tile = blockIdx.y * gridDim.x + blockIdx.x;
  if (!pending[tile]) return;
  const u32 cnt = nfl.counts[tile];
  if (cnt > NFCAP) { flat_dirs_window<TABLE>(z, dirA, dlo, dhi, dirB, ny, nx, nodata, rw); return; }   // a dense tile
  const int r0 = blockIdx.y * T, c0 = blockIdx.x * T;
  for (u32 k = threadIdx.x; k < cnt; k += 256) {
    const u32 li = nfl.entries[(size_t)tile * NFCAP + k] & 0x0FFFu;
    const u32 eq = nfl.eq[(size_t)tile * NFCAP + k];
    const long long g = (long long)(r0 + (int)(li >> 6)) * nx + (c0 + (int)(li & 63u));
    if (dirB[g] != HC_DIR_PENDING) continue;
    uint8_t out = HC_DIR_NOFLOW;
    const u32 lo = dlo[g];
    if (lo) {
      long long mn = 2ll * (long long)lo - (long long)dhi[g];
      int best = -1;
      bool bdiag = false;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        if (!((eq >> c_k_of_j[TABLE][j]) & 1u)) continue;
        const long long gn = g + (long long)c_dr[TABLE][j] * nx + c_dc[TABLE][j];
        long long key;
        if (dirA[gn] == HC_DIR_NOFLOW) {
          const u32 nlo = dlo[gn];
          if (!nlo) continue;
          key = 2ll * (long long)nlo - (long long)dhi[gn];
        } else {
          key = -0x4000000000000000ll;  // low edge: below every mask value
        }
        const bool kdiag = (c_dr[TABLE][j] != 0) && (c_dc[TABLE][j] != 0);
        if (key < mn || (key == mn && best >= 0 && bdiag && !kdiag)) { mn = key; best = j; bdiag = kdiag; }
      }
      if (best >= 0) out = (uint8_t)c_code[TABLE][best];
    }
    dirB[g] = out;
  }
}

size_t flats_workspace(int64_t ny, int64_t nx) {
  size_t n = (size_t)(ny + 2) * nx;
  size_t tiles = (size_t)((ny + T - 1) / T) * ((nx + T - 1) / T);
  // NOFLOW records (3 B x NFCAP per tile), two distance rasters, per-tile maps and queues
  return hc_align(tiles * NFCAP * 2) + hc_align(tiles * NFCAP) + 2 * hc_align(n * 4) + 4 * hc_align(tiles) + 4 * hc_align(tiles * 4) + hc_align(256);
}

// ---- the pass in pieces (the whole-raster call runs them back to back; a band interleaves the
// ---- relaxation with seam exchanges).  State lives in ctx->band.
static int flats_fast_tier(hc_ctx *ctx, const float *z, const uint8_t *dirA, uint8_t *dirB, int64_t ny, int64_t nx,
                           float nodata, int table, hc_rows rw, cudaStream_t st, const hc_nf_lists *pre) {
  hc_band_state *bs = &ctx->band;
  const int tx = (int)((nx + T - 1) / T), ty = (int)((ny + T - 1) / T);
  const size_t tiles = (size_t)tx * ty;
  bs->pending = (uint8_t *)arena_take(ctx, tiles);
  bs->active = (uint8_t *)arena_take(ctx, tiles);
  bs->qA = (u32 *)arena_take(ctx, tiles * 4);
  bs->qB = (u32 *)arena_take(ctx, tiles * 4);
  bs->queued = (u32 *)arena_take(ctx, tiles * 4);
  bs->flag = (u32 *)arena_take(ctx, 256);   // [1] pending tiles, [2..3] halo changed, [8] n(qA), [9] n(qB)
  if (!bs->pending || !bs->active || !bs->qA || !bs->qB || !bs->queued || !bs->flag) return HC_ERR_NOMEM;
  bs->fz = z; bs->dirA = dirA; bs->dirB = dirB; bs->nb = ny; bs->nx = nx; bs->nodata = nodata; bs->table = table;

  static bool attr_set = false;
  if (!attr_set) {
    HC_CUDA(cudaFuncSetAttribute(k_flat_relax, cudaFuncAttributeMaxDynamicSharedMemorySize, FLAT_SMEM));
    attr_set = true;
  }
  dim3 tg(tx, ty);
  // per-tile records of the NOFLOW cells: handed in by the fused apply + D8 pass, or listed here from the D8 raster
  hc_nf_lists nfl;
  if (pre && pre->entries) nfl = *pre;
  else {
    nfl.entries = (unsigned short *)arena_take(ctx, tiles * NFCAP * 2);
    nfl.eq = (uint8_t *)arena_take(ctx, tiles * NFCAP);
    nfl.counts = (u32 *)arena_take(ctx, tiles * 4);
    if (!nfl.entries || !nfl.eq || !nfl.counts) return HC_ERR_NOMEM;
    HC_PROF(ctx, st, "k_flat_list");
    k_flat_list<<<tg, 256, 0, st>>>(z, dirA, (int)ny, (int)nx, nodata, nfl, rw);
    HC_LAUNCH_CHECK(ctx);
  }
  bs->tlist = nfl.entries; bs->tmask = nfl.eq; bs->tcount = nfl.counts;
  HC_PROF(ctx, st, "k_flat_fast_sparse");
  if (table == HC_TABLE_WBT)
    k_flat_fast_sparse<HC_TABLE_WBT><<<tg, FS_THREADS, FS_SMEM, st>>>(dirA, dirB, nfl, bs->pending, (int)ny, (int)nx, ctx->d_dbg, rw);
  else
    k_flat_fast_sparse<HC_TABLE_TAUDEM><<<tg, FS_THREADS, FS_SMEM, st>>>(dirA, dirB, nfl, bs->pending, (int)ny, (int)nx, ctx->d_dbg, rw);
  HC_LAUNCH_CHECK(ctx);
  HC_CUDA(cudaMemsetAsync(bs->flag, 0, 64, st));
  HC_PROF(ctx, st, "k_flat_dilate");
  k_flat_dilate<<<(unsigned)((tiles + 255) / 256), 256, 0, st>>>(bs->pending, bs->active, tx, ty, bs->flag + 1, rw);
  HC_LAUNCH_CHECK(ctx);
  HC_CUDA(cudaMemcpyAsync(ctx->h_mail, bs->flag + 1, 4, cudaMemcpyDeviceToHost, st));
  HC_CUDA(cudaStreamSynchronize(st));
  bs->n_pending = (int)ctx->h_mail[0];
  return HC_OK;
}

static int flats_slow_init(hc_ctx *ctx, hc_rows rw, cudaStream_t st) {
  hc_band_state *bs = &ctx->band;
  const int64_t ny = bs->nb, nx = bs->nx;
  const int tx = (int)((nx + T - 1) / T), ty = (int)((ny + T - 1) / T);
  const size_t tiles = (size_t)tx * ty;
  // state rasters carry the band's halo rows (row -1 and row ny) like z and dir do
  const size_t nh = (size_t)(ny + 2) * nx;
  u32 *dlo = (u32 *)arena_take(ctx, nh * 4);
  u32 *dhi = (u32 *)arena_take(ctx, nh * 4);
  if (!dlo || !dhi) return HC_ERR_NOMEM;
  HC_CUDA(cudaMemsetAsync(dlo, 0, (size_t)nx * 4, st));
  HC_CUDA(cudaMemsetAsync(dlo + (size_t)(ny + 1) * nx, 0, (size_t)nx * 4, st));
  HC_CUDA(cudaMemsetAsync(dhi, 0, (size_t)nx * 4, st));
  HC_CUDA(cudaMemsetAsync(dhi + (size_t)(ny + 1) * nx, 0, (size_t)nx * 4, st));
  bs->dlo = dlo + nx; bs->dhi = dhi + nx;
  dim3 tg(tx, ty);
  uint8_t *startq = (uint8_t *)arena_take(ctx, tiles);
  bs->visited = (uint8_t *)arena_take(ctx, tiles);   // per-tile "visited" map of the slow tier
  if (!startq || !bs->visited) return HC_ERR_NOMEM;
  HC_CUDA(cudaMemsetAsync(bs->visited, 0, tiles, st));
  const int start_all = getenv("HC_FLAT_START_ALL") ? 1 : 0;
  const hc_nf_lists nfl = {bs->tlist, bs->tmask, bs->tcount};
  HC_PROF(ctx, st, "k_flat_init_lists");
  k_flat_init_lists<<<tg, 256, 0, st>>>(bs->fz, bs->dirA, bs->dirB, bs->dlo, bs->dhi, bs->active, bs->pending, startq, nfl, (int)ny,
                                        (int)nx, bs->nodata, rw, start_all);
  HC_LAUNCH_CHECK(ctx);
  // first relaxation visits every active tile: queue A
  HC_CUDA(cudaMemsetAsync(bs->queued, 0, tiles * 4, st));
  HC_PROF(ctx, st, "k_flat_queue_tiles");
  k_flat_queue_tiles<<<(unsigned)((tiles + 255) / 256), 256, 0, st>>>(startq, 0u, (u32)tiles, bs->queued, bs->qA, bs->flag + 8);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

// relax until the queue runs dry (all rounds in ONE cooperative launch, no host read-back); on return queue A (the next
// call's input) is empty and clean
static int flats_relax_loop(hc_ctx *ctx, hc_rows rw, cudaStream_t st) {
  hc_band_state *bs = &ctx->band;
  const int64_t ny = bs->nb, nx = bs->nx;
  int tx = (int)((nx + T - 1) / T), ty = (int)((ny + T - 1) / T);
  const size_t tiles = (size_t)tx * ty;
  static int max_blocks = 0;
  if (!max_blocks) {
    int per_sm = 0;
    HC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_flat_relax, NTHREADS, FLAT_SMEM));
    if (const char *e = getenv("HC_FLAT_CTAS_PER_SM")) { const int v = atoi(e); if (v >= 1 && v < per_sm) per_sm = v; }   // (experiments)
    max_blocks = per_sm * ctx->sm_count;
    if (max_blocks < 1) { hc_set_error("flats: cooperative launch: no resident block"); return HC_ERR_CUDA; }
  }
  const unsigned grid = (unsigned)(tiles < (size_t)max_blocks ? tiles : (size_t)max_blocks);
  int iny = (int)ny, inx = (int)nx;
  u32 *ctl = bs->flag + 8;
  void *args[] = {(void *)&bs->fz, (void *)&bs->dirA, (void *)&bs->dlo, (void *)&bs->dhi, (void *)&bs->active, (void *)&bs->qA,
                  (void *)&bs->qB, (void *)&ctl, (void *)&bs->queued, (void *)&bs->visited, (void *)&bs->tlist,
                  (void *)&bs->tmask, (void *)&bs->tcount, (void *)&iny, (void *)&inx, (void *)&bs->nodata, (void *)&ctx->d_dbg,
                  (void *)&rw, (void *)&tx, (void *)&ty};
  HC_PROF(ctx, st, "k_flat_relax");
  HC_CUDA(cudaLaunchCooperativeKernel((void *)k_flat_relax, dim3(grid), dim3(NTHREADS), args, FLAT_SMEM, st));
  HC_LAUNCH_CHECK(ctx);
  if (getenv("HC_TRACE")) {
    u32 h[48];
    HC_CUDA(cudaMemcpyAsync(h, ctl + 8, sizeof h, cudaMemcpyDeviceToHost, st));
    HC_CUDA(cudaStreamSynchronize(st));
    fprintf(stderr, "[flats] rounds (tiles @ us):");
    for (int i = 0; i < 24; i++) {
      if (i >= 14 && h[2 * i] == 1) fprintf(stderr, " 1@tile(%u,%u)%s", (h[2 * i + 1] & 0x7FFFFFFFu) / (u32)tx, (h[2 * i + 1] & 0x7FFFFFFFu) % (u32)tx, (h[2 * i + 1] >> 31) ? "L" : "l");
      else fprintf(stderr, " %u@%.0f", h[2 * i], (double)(h[2 * i + 1] - h[1]) * 1e-3);
      if (!h[2 * i]) break;
    }
    fprintf(stderr, "\n");
  }
  HC_CUDA(cudaMemsetAsync(bs->queued, 0, tiles * 4, st));  // put_halos may queue seam tiles for the next call
  return HC_OK;
}

static int flats_dirs(hc_ctx *ctx, hc_rows rw, cudaStream_t st) {
  hc_band_state *bs = &ctx->band;
  const int64_t ny = bs->nb, nx = bs->nx;
  dim3 tg((unsigned)((nx + T - 1) / T), (unsigned)((ny + T - 1) / T));
  const hc_nf_lists nfl = {bs->tlist, bs->tmask, bs->tcount};
  HC_PROF(ctx, st, "k_flat_dirs_lists");
  if (bs->table == HC_TABLE_WBT)
    k_flat_dirs_lists<HC_TABLE_WBT><<<tg, 256, 0, st>>>(bs->fz, bs->dirA, bs->dlo, bs->dhi, bs->dirB, bs->pending, nfl, (int)ny, (int)nx, bs->nodata, rw);
  else
    k_flat_dirs_lists<HC_TABLE_TAUDEM><<<tg, 256, 0, st>>>(bs->fz, bs->dirA, bs->dlo, bs->dhi, bs->dirB, bs->pending, nfl, (int)ny, (int)nx, bs->nodata, rw);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

// dirA: D8 codes with NOFLOW on flats (read only); dirB: output (may not alias dirA).
int flats_core(hc_ctx *ctx, const float *z, const uint8_t *dirA, uint8_t *dirB, int64_t ny, int64_t nx,
               float nodata, int table, cudaStream_t st, const hc_nf_lists *pre = nullptr) {
  hc_rows rw = hc_make_rows(ny, 0);
  HC_TRY(flats_fast_tier(ctx, z, dirA, dirB, ny, nx, nodata, table, rw, st, pre));
  if (!ctx->band.n_pending) return HC_OK;  // every flat was resolved inside its apron
  HC_TRY(flats_slow_init(ctx, rw, st));
  HC_TRY(flats_relax_loop(ctx, rw, st));
  return flats_dirs(ctx, rw, st);
}

int flats_whole_begin(hc_ctx *ctx, const float *z, const uint8_t *dirA, uint8_t *dirB, int64_t ny, int64_t nx, float nodata,
                      int table, cudaStream_t st, const uint8_t **tile_pending, int *n_pending, const hc_nf_lists *pre) {
  HC_TRY(flats_fast_tier(ctx, z, dirA, dirB, ny, nx, nodata, table, hc_make_rows(ny, 0), st, pre));
  *tile_pending = ctx->band.pending;
  *n_pending = ctx->band.n_pending;
  return HC_OK;
}

int flats_whole_finish(hc_ctx *ctx, cudaStream_t st) {
  hc_rows rw = hc_make_rows(ctx->band.nb, 0);
  if (!ctx->band.n_pending) return HC_OK;
  HC_TRY(flats_slow_init(ctx, rw, st));
  HC_TRY(flats_relax_loop(ctx, rw, st));
  return flats_dirs(ctx, rw, st);
}

// in-place public form: dir is copied aside first
int flats_run(hc_ctx *ctx, const float *z, uint8_t *dir, int64_t ny, int64_t nx, float nodata, int table,
              cudaStream_t st) {
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (table != HC_TABLE_WBT && table != HC_TABLE_TAUDEM) { hc_set_error("flats: bad table %d", table); return HC_ERR_ARG; }
  size_t n = (size_t)ny * nx;
  HC_TRY(arena_reset(ctx));
  uint8_t *dirA = (uint8_t *)arena_take(ctx, n);
  if (!dirA) return HC_ERR_NOMEM;
  HC_CUDA(cudaMemcpyAsync(dirA, dir, n, cudaMemcpyDeviceToDevice, st));
  return flats_core(ctx, z, dirA, dir, ny, nx, nodata, table, st);
}

// D8 into an arena temporary, then flats straight into the caller's raster (no copy)
int d8_flats_run(hc_ctx *ctx, const float *z, uint8_t *dir, float *slope, int64_t ny, int64_t nx, float nodata,
                 double dx, double dy, int table, cudaStream_t st) {
  if (ny <= 0 || nx <= 0) return HC_OK;
  size_t n = (size_t)ny * nx;
  HC_TRY(arena_reset(ctx));
  uint8_t *dirA = (uint8_t *)arena_take(ctx, n);
  if (!dirA) return HC_ERR_NOMEM;
  // D8 writes both rasters; the flat pass then only rewrites the NOFLOW cells of `dir`
  HC_TRY(d8_run_rows(ctx, z, dirA, slope, ny, nx, nodata, dx, dy, table, hc_make_rows(ny, 0), st, dir));
  return flats_core(ctx, z, dirA, dir, ny, nx, nodata, table, st);
}

// ================================================================================================
// row bands: the fast tier leaves every flat that touches a seam pending; the slow
// tier's distance fields are relaxed band-locally, seam rows are exchanged, and the relaxation is
// repeated until no halo row changes anywhere (chaotic relaxation converges to the same geodesic
// distances in any order, so the result equals the whole-raster one bit for bit).
// ================================================================================================
int band_flats_begin(hc_ctx *ctx, const float *z, const uint8_t *dirA, uint8_t *dirB, int64_t nb, int64_t nx,
                     float nodata, int table, int flags, cudaStream_t st) {
  if (table != HC_TABLE_WBT && table != HC_TABLE_TAUDEM) { hc_set_error("flats: bad table %d", table); return HC_ERR_ARG; }
  hc_rows rw = hc_make_rows(nb, flags);
  HC_TRY(arena_reset(ctx));
  ctx->band.flags = flags;
  HC_CUDA(cudaMemcpyAsync(dirB, dirA, (size_t)nb * nx, cudaMemcpyDeviceToDevice, st));  // the fast tier only rewrites NOFLOW cells
  HC_TRY(flats_fast_tier(ctx, z, dirA, dirB, nb, nx, nodata, table, rw, st, nullptr));
  HC_TRY(flats_slow_init(ctx, rw, st));
  ctx->band.flats_ready = 1;
  return HC_OK;
}

// out[side][field][nx] (u32): side 0 = first own row, 1 = last own row; fields NOFLOW flag (informational: the
// receiver has the row's D8 codes in its halo already), dlo, dhi
__global__ void k_flat_get_seams(const uint8_t *__restrict__ dirA, const u32 *__restrict__ dlo, const u32 *__restrict__ dhi,
                                 int ny, int nx, u32 *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * nx) return;
  int side = i / nx, c = i - side * nx;
  size_t g = (size_t)(side ? ny - 1 : 0) * nx + c;
  u32 *o = out + (size_t)side * 3 * nx;
  const u32 f = dirA[g] == HC_DIR_NOFLOW ? 1u : 0u;
  o[c] = f;
  o[nx + c] = f ? dlo[g] : 0u;
  o[2 * (size_t)nx + c] = f ? dhi[g] : 0u;
}

// write a halo row from the neighbour's seam record; flag[2] != 0 if anything differed
__global__ void k_flat_put_halo(u32 *dlo, u32 *dhi, long long row, int nx, const u32 *__restrict__ in, u32 *flag) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nx) return;
  long long g = row * nx + c;
  // (only NOFLOW cells carry state; a cell with a direction is a constant low edge on both sides of the seam)
  u32 f = in[c], lo = f ? in[nx + c] : 0u, hi = f ? in[2 * (size_t)nx + c] : 0u;
  if (dlo[g] != lo || dhi[g] != hi) {
    dlo[g] = lo; dhi[g] = hi;
    *flag = 1u;
  }
}

int band_flats_get_seams(hc_ctx *ctx, uint32_t *out, cudaStream_t st) {
  hc_band_state *bs = &ctx->band;
  if (!bs->flats_ready) { hc_set_error("band flats: begin first"); return HC_ERR_ARG; }
  HC_PROF(ctx, st, "k_flat_get_seams");
  k_flat_get_seams<<<(unsigned)((2 * bs->nx + 255) / 256), 256, 0, st>>>(bs->dirA, bs->dlo, bs->dhi, (int)bs->nb, (int)bs->nx, out);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

int band_flats_put_halos(hc_ctx *ctx, const uint32_t *top, const uint32_t *bot, int *changed, cudaStream_t st) {
  hc_band_state *bs = &ctx->band;
  if (!bs->flats_ready) { hc_set_error("band flats: begin first"); return HC_ERR_ARG; }
  const int64_t ny = bs->nb, nx = bs->nx;
  const int tx = (int)((nx + T - 1) / T), ty = (int)((ny + T - 1) / T);
  HC_CUDA(cudaMemsetAsync(bs->flag + 2, 0, 8, st));
  unsigned gb = (unsigned)((nx + 255) / 256);
  if (top && (bs->flags & 1)) {
    HC_PROF(ctx, st, "k_flat_put_halo");
    k_flat_put_halo<<<gb, 256, 0, st>>>(bs->dlo, bs->dhi, -1ll, (int)nx, top, bs->flag + 2);
    HC_LAUNCH_CHECK(ctx);
  }
  if (bot && (bs->flags & 2)) {
    HC_PROF(ctx, st, "k_flat_put_halo");
    k_flat_put_halo<<<gb, 256, 0, st>>>(bs->dlo, bs->dhi, (long long)ny, (int)nx, bot, bs->flag + 3);
    HC_LAUNCH_CHECK(ctx);
  }
  HC_CUDA(cudaMemcpyAsync(ctx->h_mail, bs->flag + 2, 8, cudaMemcpyDeviceToHost, st));
  HC_CUDA(cudaStreamSynchronize(st));
  // tiles next to a halo row that changed are relaxed again: into queue A (the next relax call's input)
  if (ctx->h_mail[0]) {
    k_flat_queue_tiles<<<(unsigned)((tx + 255) / 256), 256, 0, st>>>(bs->active, 0u, (u32)tx, bs->queued, bs->qA, bs->flag + 8);
    HC_LAUNCH_CHECK(ctx);
  }
  if (ctx->h_mail[1]) {
    k_flat_queue_tiles<<<(unsigned)((tx + 255) / 256), 256, 0, st>>>(bs->active, (u32)((ty - 1) * tx), (u32)tx, bs->queued, bs->qA, bs->flag + 8);
    HC_LAUNCH_CHECK(ctx);
  }
  if (changed) *changed = (ctx->h_mail[0] || ctx->h_mail[1]) ? 1 : 0;
  return HC_OK;
}

// device-flag form of put_halos: nothing is read back.  `flag_dev[0]` is OR-ed with "some halo row changed"; the seam tile
// rows are re-queued by a kernel that looks at the change flags itself, so the host learns about convergence from ONE
// all_reduce of the ranks' flag words per exchange round instead of a read-back per band plus that all_reduce.
__global__ void k_flat_queue_tiles_if(const uint8_t *__restrict__ active, u32 first, u32 count, u32 *queued, u32 *list, u32 *n,
                                      const u32 *__restrict__ changed, int *__restrict__ flag_out) {
  if (!*changed) return;
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && flag_out) *flag_out = 1;
  if (i >= count) return;
  u32 t = first + i;
  if (active[t] && atomicExch(&queued[t], 1u) == 0u) list[atomicAdd(n, 1u)] = t;
}

int band_flats_put_halos_flag(hc_ctx *ctx, const uint32_t *top, const uint32_t *bot, int *flag_dev, cudaStream_t st) {
  hc_band_state *bs = &ctx->band;
  if (!bs->flats_ready) { hc_set_error("band flats: begin first"); return HC_ERR_ARG; }
  const int64_t ny = bs->nb, nx = bs->nx;
  const int tx = (int)((nx + T - 1) / T), ty = (int)((ny + T - 1) / T);
  HC_CUDA(cudaMemsetAsync(bs->flag + 2, 0, 8, st));
  unsigned gb = (unsigned)((nx + 255) / 256), gt = (unsigned)((tx + 255) / 256);
  if (top && (bs->flags & 1)) {
    HC_PROF(ctx, st, "k_flat_put_halo");
    k_flat_put_halo<<<gb, 256, 0, st>>>(bs->dlo, bs->dhi, -1ll, (int)nx, top, bs->flag + 2);
    HC_LAUNCH_CHECK(ctx);
    k_flat_queue_tiles_if<<<gt, 256, 0, st>>>(bs->active, 0u, (u32)tx, bs->queued, bs->qA, bs->flag + 8, bs->flag + 2, flag_dev);
    HC_LAUNCH_CHECK(ctx);
  }
  if (bot && (bs->flags & 2)) {
    HC_PROF(ctx, st, "k_flat_put_halo");
    k_flat_put_halo<<<gb, 256, 0, st>>>(bs->dlo, bs->dhi, (long long)ny, (int)nx, bot, bs->flag + 3);
    HC_LAUNCH_CHECK(ctx);
    k_flat_queue_tiles_if<<<gt, 256, 0, st>>>(bs->active, (u32)((ty - 1) * tx), (u32)tx, bs->queued, bs->qA, bs->flag + 8, bs->flag + 3, flag_dev);
    HC_LAUNCH_CHECK(ctx);
  }
  return HC_OK;
}

int band_flats_relax(hc_ctx *ctx, cudaStream_t st) {
  hc_band_state *bs = &ctx->band;
  if (!bs->flats_ready) { hc_set_error("band flats: begin first"); return HC_ERR_ARG; }
  return flats_relax_loop(ctx, hc_make_rows(bs->nb, bs->flags), st);
}

int band_flats_finish(hc_ctx *ctx, cudaStream_t st) {
  hc_band_state *bs = &ctx->band;
  if (!bs->flats_ready) { hc_set_error("band flats: begin first"); return HC_ERR_ARG; }
  bs->flats_ready = 0;
  return flats_dirs(ctx, hc_make_rows(bs->nb, bs->flags), st);
}
