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
//   3. k_flatten  every other cell takes its (now resolved) perimeter target's id.
//   4. Boruvka rounds on the basin graph.  Edge weight between adjacent basins = the lowest pass
//                 min over adjacent cell pairs of max(z_a, z_b).  Each still-closed basin picks
//                 its cheapest outgoing edge straight from the grid (k_minedge, one 64-bit
//                 atomicMin per basin-boundary cell at most), the pointer forest of those picks
//                 is rooted and compressed on the K-entry tables, trees that reach basin 0 are
//                 resolved, the others contract into their root.  A basin's final level is the
//                 max of the pick weights up its contraction chain (see DESIGN.md for the proof
//                 that this equals the minimax level).
//   5. k_apply    out = max(z, level[basin]).
//
// Only float comparisons/min/max touch elevations, so the result is bit-identical to the serial
// priority-flood (the eps=0 surface is unique).
#include <stdlib.h>

#include "common.cuh"

#define T HC_TILE
#define TH (T + 2)
#define NTHREADS 256

// ---------------------------------------------------------------------------------------------
// 1. pointers + in-tile compression
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS)
k_ptr(const float *__restrict__ z, const uint8_t *__restrict__ ext, u32 *__restrict__ P, int ny, int nx,
      float nodata, u32 *__restrict__ pit_counter) {
  __shared__ float sz[TH * TH];
  __shared__ unsigned short snxt[T * T];
  __shared__ u32 stv[T * T];
  __shared__ u32 s_npit, s_base;

  const int r0 = blockIdx.y * T, c0 = blockIdx.x * T;
  const int tid = threadIdx.x;
  if (tid == 0) s_npit = 0;

  for (int i = tid; i < TH * TH; i += NTHREADS) {
    int lr = i / TH, lc = i - lr * TH;
    int r = r0 + lr - 1, c = c0 + lc - 1;
    float v = __int_as_float(0x7fc00000);
    if (r >= 0 && r < ny && c >= 0 && c < nx) {
      v = __ldg(&z[(size_t)r * nx + c]);
      if (v == nodata) v = __int_as_float(0x7fc00000);
    }
    sz[i] = v;
  }
  __syncthreads();

  // pass A: classify, pick lowest neighbour, count pits
  for (int i = tid; i < T * T; i += NTHREADS) {
    int lr = i / T, lc = i - lr * T;
    int r = r0 + lr, c = c0 + lc;
    unsigned short nxt = (unsigned short)i;
    u32 tv = HC_NONE;  // nodata / outside raster
    if (r < ny && c < nx) {
      float v = sz[(lr + 1) * TH + lc + 1];
      if (v == v) {
        bool seed = (r == 0 || c == 0 || r == ny - 1 || c == nx - 1);
        float bz = v;
        int bdr = 0, bdc = 0;
#pragma unroll
        for (int dr = -1; dr <= 1; dr++)
#pragma unroll
          for (int dc = -1; dc <= 1; dc++) {
            if (dr == 0 && dc == 0) continue;
            float zn = sz[(lr + 1 + dr) * TH + lc + 1 + dc];
            if (zn != zn) {
              // nodata neighbour: seed if that nodata is exterior (ext == NULL: all nodata is)
              int rr = r + dr, cc = c + dc;
              if (rr >= 0 && rr < ny && cc >= 0 && cc < nx) {
                if (ext == nullptr || ext[(size_t)rr * nx + cc]) seed = true;
              }
              continue;
            }
            // total order (z, index): row-major index grows with (dr, dc) lexicographically
            bool lower = zn < bz || (zn == bz && (dr < bdr || (dr == bdr && dc < bdc)));
            if (lower) { bz = zn; bdr = dr; bdc = dc; }
          }
        if (seed) {
          tv = HC_TAG;  // basin 0
        } else if (bdr == 0 && bdc == 0) {
          tv = HC_TAG | 0x7FFFFFFEu;  // pit, id patched below
          atomicAdd(&s_npit, 1u);
        } else {
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
      stv[i] = HC_TAG | (s_base + k + 1u);
    }
  }
  __syncthreads();

  // pass B: pointer jumping inside the tile until every cell points at a terminal
  for (;;) {
    int changed = 0;
    for (int i = tid; i < T * T; i += NTHREADS) {
      unsigned short p = snxt[i];
      unsigned short pp = snxt[p];
      if (pp != p) { snxt[i] = pp; changed = 1; }
    }
    if (!__syncthreads_or(changed)) break;
  }

  for (int i = tid; i < T * T; i += NTHREADS) {
    int lr = i / T, lc = i - lr * T;
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
template <bool EMIT>
__global__ void __launch_bounds__(NTHREADS)
k_minedge(const float *__restrict__ z, const u32 *__restrict__ P, const u32 *__restrict__ rep,
          u64 *__restrict__ best, int ny, int nx, u32 *__restrict__ eA, u32 *__restrict__ eB,
          float *__restrict__ eW, u32 *__restrict__ ecount, u32 ecap) {
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
    u32 id = __ldg(&P[g]) & 0x7FFFFFFFu;
    float v = __ldg(&z[g]);
    u32 a = HC_NONE;
    if (inb && id != HC_LAB_NODATA) a = id ? __ldg(&rep[id]) : 0u;
    sz[i] = v;
    sa[i] = a;
  }
  __syncthreads();
  u32 cnt = 0;
  for (int i = tid; i < T * T; i += NTHREADS) {
    int lr = i >> 6, lc = i & (T - 1);
    int ci = (lr + 1) * TH + lc + 1;
    u32 a = sa[ci];
    if (a == 0u || a == HC_NONE) continue;
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
    if (a == 0u || a == HC_NONE) continue;
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

// a Boruvka pick round on the compact edge list
__global__ void __launch_bounds__(256)
k_minedge_list(const u32 *__restrict__ eA, const u32 *__restrict__ eB, const float *__restrict__ eW, u32 ne,
               const u32 *__restrict__ rep, u64 *__restrict__ best) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ne) return;
  u32 a = __ldg(&rep[eA[i]]);
  if (a == 0u) return;
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
__global__ void k_hook(const u32 *rep, const u64 *best, u32 *ptr, float *lvl, u32 K1) {
  u32 b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= K1) return;
  if (b == 0 || rep[b] != b) { ptr[b] = b == 0 ? 0u : rep[b]; return; }
  u64 k = best[b];
  if (k == ~0ull) { ptr[b] = 0u; return; }  // no neighbour at all (walled in): never raised
  float w = d_funord((u32)(k >> 32));
  ptr[b] = (u32)(k & 0xFFFFFFFFu);
  lvl[b] = fmaxf(lvl[b], w);
}

// break mutual pairs: the smaller id becomes the root of the pair's tree
__global__ void k_mutual(const u32 *rep, const u32 *ptr, u32 *ptr2, u32 K1) {
  u32 b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= K1) return;
  u32 p = ptr[b];
  if (b != 0 && rep[b] == b && p != 0 && ptr[p] == b && b < p) p = b;
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
__global__ void k_merge(u32 *rep, u32 *hp, const u32 *ptr, u32 K1, u32 *n_active) {
  u32 b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= K1 || b == 0) return;
  if (rep[b] != b) return;
  u32 r = ptr[b];
  if (r != b) { hp[b] = r; rep[b] = r; }
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
__global__ void k_levels(const u32 *hp, const float *lvl, float *level0, u32 K1) {
  u32 b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= K1) return;
  float m = -INFINITY;
  u32 x = b;
  while (x != 0) { m = fmaxf(m, lvl[x]); x = hp[x]; }
  level0[b] = m;
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

int fill_run(hc_ctx *ctx, const float *z, float *out, int64_t ny, int64_t nx, float nodata, int seed_mode,
             cudaStream_t st) {
  if (ny <= 0 || nx <= 0) return HC_OK;
  size_t n = (size_t)ny * nx;
  if (n >= 0x7FFFFFF0ull) { hc_set_error("raster too large for 31-bit cell indices"); return HC_ERR_TOO_LARGE; }
  const int tx = (int)((nx + T - 1) / T), ty = (int)((ny + T - 1) / T);
  const size_t tiles = (size_t)tx * ty;
  HC_TRY(arena_reset(ctx));

  uint8_t *ext = nullptr;
  if (seed_mode == HC_SEED_WBT) {
    ext = (uint8_t *)arena_take(ctx, n);
    if (!ext) return HC_ERR_NOMEM;
    int has_interior = 0;
    HC_TRY(ext_nodata_mask(ctx, z, ext, ny, nx, nodata, &has_interior, st));
    if (!has_interior) ext = nullptr;  // every nodata cell sits on the raster border: all exterior
  }
  u32 *P = (u32 *)arena_take(ctx, n * 4);
  uint8_t *actA = (uint8_t *)arena_take(ctx, tiles);
  uint8_t *actB = (uint8_t *)arena_take(ctx, tiles);
  u32 *cnt = (u32 *)arena_take(ctx, 256);  // [0]=pit counter [1]=changed [2]=n_active
  if (!P || !actA || !actB || !cnt) return HC_ERR_NOMEM;

  HC_CUDA(cudaMemsetAsync(cnt, 0, 256, st));
  dim3 tg(tx, ty);
  HC_PROF(ctx, st, "k_ptr");
  k_ptr<<<tg, NTHREADS, 0, st>>>(z, ext, P, (int)ny, (int)nx, nodata, cnt);
  HC_LAUNCH_CHECK(ctx);
  HC_PROF(ctx, st, "k_perim");
  k_perim<<<tg, 256, 0, st>>>(P, (int)ny, (int)nx);
  HC_LAUNCH_CHECK(ctx);
  int gblocks = ctx->sm_count * 8;
  HC_PROF(ctx, st, "k_flatten");
  k_flatten<<<gblocks, 256, 0, st>>>(P, n);
  HC_LAUNCH_CHECK(ctx);

  HC_CUDA(cudaMemcpyAsync(ctx->h_mail, cnt, 4, cudaMemcpyDeviceToHost, st));
  HC_CUDA(cudaStreamSynchronize(st));
  const u32 K = ctx->h_mail[0];
  const u32 K1 = K + 1;
  ctx->fill_basins = K;
  ctx->fill_rounds = 0;

  u64 *best = (u64 *)arena_take(ctx, (size_t)K1 * 8);
  u32 *rep = (u32 *)arena_take(ctx, (size_t)K1 * 4);
  u32 *hp = (u32 *)arena_take(ctx, (size_t)K1 * 4);
  u32 *ptr = (u32 *)arena_take(ctx, (size_t)K1 * 4);
  u32 *ptr2 = (u32 *)arena_take(ctx, (size_t)K1 * 4);
  float *lvl = (float *)arena_take(ctx, (size_t)K1 * 4);
  if (!best || !rep || !hp || !ptr || !ptr2 || !lvl) return HC_ERR_NOMEM;
  float *level0 = (float *)ptr2;  // reused after the rounds
  // compact edge list for the rounds after the second grid pass
  size_t ecap_sz = n / 4 > (size_t)(1u << 20) ? n / 4 : (size_t)(1u << 20);
  if (ecap_sz > (size_t)(1u << 28)) ecap_sz = (size_t)(1u << 28);
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
  k_tab_init<<<tgK, tb, 0, st>>>(rep, lvl, hp, K1);
  HC_LAUNCH_CHECK(ctx);

  u32 active = K;
  int round = 0;
  while (active > 0) {
    if (round > 64) { hc_set_error("fill: Boruvka did not converge"); return HC_ERR_INTERNAL; }
    HC_CUDA(cudaMemsetAsync(best, 0xFF, (size_t)K1 * 8, st));
    if (have_list) {
      HC_PROF(ctx, st, "k_minedge_list");
      k_minedge_list<<<(ne + 255) / 256, 256, 0, st>>>(eA, eB, eW, ne, rep, best);
      HC_LAUNCH_CHECK(ctx);
    } else if (round == 1) {
      HC_CUDA(cudaMemsetAsync(cnt + 4, 0, 4, st));
      HC_PROF(ctx, st, "k_minedge_emit");
      k_minedge<true><<<tg, NTHREADS, 0, st>>>(z, P, rep, best, (int)ny, (int)nx, eA, eB, eW, cnt + 4, ecap);
      HC_LAUNCH_CHECK(ctx);
      HC_CUDA(cudaMemcpyAsync(ctx->h_mail + 4, cnt + 4, 4, cudaMemcpyDeviceToHost, st));
    } else {
      HC_PROF(ctx, st, "k_minedge");
      k_minedge<false><<<tg, NTHREADS, 0, st>>>(z, P, rep, best, (int)ny, (int)nx, eA, eB, eW, cnt + 4, ecap);
      HC_LAUNCH_CHECK(ctx);
    }
    HC_PROF(ctx, st, "k_hook");
    k_hook<<<tgK, tb, 0, st>>>(rep, best, ptr2, lvl, K1);
    HC_LAUNCH_CHECK(ctx);
    HC_PROF(ctx, st, "k_mutual");
    k_mutual<<<tgK, tb, 0, st>>>(rep, ptr2, ptr, K1);
    HC_LAUNCH_CHECK(ctx);
    // pointer doubling to the roots: 4 doublings per read-back
    for (int it = 0; it < 16; it++) {
      for (int j = 0; j < 3; j++) {
        HC_PROF(ctx, st, "k_jump");
        k_jump<<<tgK, tb, 0, st>>>(ptr, K1, cnt + 3);
        HC_LAUNCH_CHECK(ctx);
      }
      HC_CUDA(cudaMemsetAsync(cnt + 1, 0, 4, st));
      HC_PROF(ctx, st, "k_jump");
      k_jump<<<tgK, tb, 0, st>>>(ptr, K1, cnt + 1);
      HC_LAUNCH_CHECK(ctx);
      HC_CUDA(cudaMemcpyAsync(ctx->h_mail, cnt + 1, 4, cudaMemcpyDeviceToHost, st));
      HC_CUDA(cudaStreamSynchronize(st));
      if (!ctx->h_mail[0]) break;
    }
    HC_CUDA(cudaMemsetAsync(cnt + 2, 0, 4, st));
    HC_PROF(ctx, st, "k_merge");
    k_merge<<<tgK, tb, 0, st>>>(rep, hp, ptr, K1, cnt + 2);
    HC_LAUNCH_CHECK(ctx);
    HC_PROF(ctx, st, "k_repflat");
    k_repflat<<<tgK, tb, 0, st>>>(rep, K1);
    HC_LAUNCH_CHECK(ctx);
    HC_CUDA(cudaMemcpyAsync(ctx->h_mail, cnt + 2, 4, cudaMemcpyDeviceToHost, st));
    HC_CUDA(cudaStreamSynchronize(st));
    active = ctx->h_mail[0];
    if (round == 1 && !have_list) {  // the emit round's count arrived with the syncs above
      ne = ctx->h_mail[4];
      have_list = ne <= ecap;
    }
    if (getenv("HC_TRACE")) fprintf(stderr, "[fill] round %d: %u closed supernodes left of %u basins (edge list %u%s)\n", round, active, K, ne, have_list ? "" : " unused");
    round++;
  }
  ctx->fill_rounds = round;

  HC_PROF(ctx, st, "k_levels");
  k_levels<<<tgK, tb, 0, st>>>(hp, lvl, level0, K1);
  HC_LAUNCH_CHECK(ctx);
  HC_PROF(ctx, st, "k_apply");
  k_apply<<<gblocks, 256, 0, st>>>(z, P, level0, out, n);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}
