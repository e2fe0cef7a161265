// label.cu — connected-component labelling with scipy.ndimage.label's numbering, and the
// exterior-nodata mask the WBT seed rule needs.
//
// Replaces scipy.ndimage.label as called by dep.create_clump_array (tools/depressions.py:1646-1669)
// and endo.find_sink (tools/endo.py:323): 8- or 4-connectivity, ids 1..n in raster-scan order of
// each component's first cell.
//
// Block-based union-find (each cell links to the smaller root, so a component's root is its first cell in
// raster order): k_lab_tile merges every 64 x 64 tile in SHARED memory (16-bit-local ids, shared atomicMin, a
// cell looks at W / N and only at the diagonals its row neighbours do not already cover) and writes each
// cell's tile-local root as a global index; k_lab_border then unions across tile borders (3 % of the cells) in
// global memory, k_lab_compress flattens.  The roots are ranked by a three-kernel exclusive scan.
#include "common.cuh"

__device__ __forceinline__ u32 uf_find(volatile u32 *L, u32 i) {
  u32 p = L[i];
  while (p != i) { i = p; p = L[i]; }
  return i;
}

__device__ __forceinline__ void uf_union(u32 *L, u32 a, u32 b) {
  volatile u32 *VL = L;
  for (;;) {
    a = uf_find(VL, a);
    b = uf_find(VL, b);
    if (a == b) return;
    if (a < b) { u32 t = a; a = b; b = t; }  // a > b: hang a under b
    u32 old = atomicMin(&L[a], b);
    if (old == a) return;
    a = old;
  }
}

#define LT 64   // labelling tile edge

__device__ __forceinline__ u32 suf_find(volatile u32 *L, u32 i) {
  u32 p = L[i];
  while (p != i) { i = p; p = L[i]; }
  return i;
}
__device__ __forceinline__ void suf_union(u32 *L, u32 a, u32 b) {
  volatile u32 *VL = L;
  for (;;) {
    a = suf_find(VL, a);
    b = suf_find(VL, b);
    if (a == b) return;
    if (a < b) { u32 t = a; a = b; b = t; }
    u32 old = atomicMin(&L[a], b);
    if (old == a) return;
    a = old;
  }
}

// Tile-local components in shared memory; L[g] = global index of the cell's tile-local root (HC_NONE off the mask).
// Horizontal connectivity needs no union at all: a row of the tile is one 64-bit word, and a cell's run start comes
// out of a count-leading-ones on that word.  Union-find then works on RUN STARTS only, with one union per pair of
// touching runs in consecutive rows (a cell asks for a union only where a new upper run begins), finds halve paths.
__device__ __forceinline__ u32 suf_find_h(u32 *L, u32 i) {
  volatile u32 *VL = L;
  u32 p = VL[i];
  while (p != i) {
    const u32 gp = VL[p];
    if (gp != p) VL[i] = gp;   // path halving (a plain store of an ancestor: links only ever move towards the root)
    i = p;
    p = gp;
  }
  return i;
}
__device__ __forceinline__ void suf_union_h(u32 *L, u32 a, u32 b) {
  for (;;) {
    a = suf_find_h(L, a);
    b = suf_find_h(L, b);
    if (a == b) return;
    if (a < b) { u32 t = a; a = b; b = t; }
    u32 old = atomicMin(&L[a], b);
    if (old == a) return;
    a = old;
  }
}
// start column of the run of ones that contains bit c of `bits`
__device__ __forceinline__ int run_start(unsigned long long bits, int c) {
  const unsigned long long upto = bits << (63 - c);          // bit c moved to bit 63
  return c - __clzll((long long)~upto) + 1;                    // ones ending at c: leading ones of upto
}

__global__ void __launch_bounds__(256)
k_lab_tile(const uint8_t *__restrict__ mask, u32 *__restrict__ L, int ny, int nx, int conn8) {
  __shared__ unsigned long long rowbits[LT];
  __shared__ u32 sl[LT * LT];
  const int r0 = blockIdx.y * LT, c0 = blockIdx.x * LT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // row words by warp ballots: warp w builds rows w, w + 8, ...
  for (int lr = warp; lr < LT; lr += 8) {
    const int r = r0 + lr;
    const int ca = c0 + lane, cb = c0 + 32 + lane;
    const bool ma = r < ny && ca < nx && mask[(size_t)r * nx + ca] != 0;
    const bool mb = r < ny && cb < nx && mask[(size_t)r * nx + cb] != 0;
    const unsigned lo = __ballot_sync(0xFFFFFFFFu, ma), hi = __ballot_sync(0xFFFFFFFFu, mb);
    if (lane == 0) rowbits[lr] = (unsigned long long)lo | ((unsigned long long)hi << 32);
  }
  for (int i = tid; i < LT * LT; i += 256) sl[i] = (u32)i;
  __syncthreads();
  for (int i = tid; i < LT * LT; i += 256) {
    const int lr = i >> 6, lc = i & (LT - 1);
    const unsigned long long cur = rowbits[lr];
    if (!((cur >> lc) & 1ull) || lr == 0) continue;
    const unsigned long long up = rowbits[lr - 1];
    const bool start = lc == 0 || !((cur >> (lc - 1)) & 1ull);
    const bool u_m = (up >> lc) & 1ull;
    const bool u_l = lc > 0 && ((up >> (lc - 1)) & 1ull);
    const bool u_r = lc < LT - 1 && ((up >> (lc + 1)) & 1ull);
    const u32 me = (u32)(lr * LT + run_start(cur, lc));
    const u32 upbase = (u32)((lr - 1) * LT);
    if (!conn8) {
      if (u_m && (start || !u_l)) suf_union_h(sl, me, upbase + (u32)run_start(up, lc));
    } else if (start) {
      if (u_m) suf_union_h(sl, me, upbase + (u32)run_start(up, lc));
      else {
        if (u_l) suf_union_h(sl, me, upbase + (u32)run_start(up, lc - 1));
        if (u_r) suf_union_h(sl, me, upbase + (u32)(lc + 1));
      }
    } else if (u_r && !u_m) {
      suf_union_h(sl, me, upbase + (u32)(lc + 1));       // a new upper run begins right of this cell
    }
  }
  __syncthreads();
  // every run start chases its chain ONCE and keeps the root (a store of an ancestor never misleads a concurrent chase);
  // the cells of the run then read it with one shared load (before, each cell walked its run's chain by itself)
  for (int i = tid; i < LT * LT; i += 256) {
    const int lr = i >> 6, lc = i & (LT - 1);
    const unsigned long long cur = rowbits[lr];
    if (((cur >> lc) & 1ull) && (lc == 0 || !((cur >> (lc - 1)) & 1ull))) {
      const u32 root = suf_find((volatile u32 *)sl, (u32)i);
      ((volatile u32 *)sl)[i] = root;
    }
  }
  __syncthreads();
  for (int i = tid; i < LT * LT; i += 256) {
    const int lr = i >> 6, lc = i & (LT - 1);
    const int r = r0 + lr, c = c0 + lc;
    if (r >= ny || c >= nx) continue;
    u32 out = HC_NONE;
    const unsigned long long cur = rowbits[lr];
    if ((cur >> lc) & 1ull) {
      const u32 root = sl[lr * LT + run_start(cur, lc)];
      out = (u32)((size_t)(r0 + (int)(root >> 6)) * nx + (c0 + (int)(root & (LT - 1))));
    }
    L[(size_t)r * nx + c] = out;
  }
}

// unions across tile borders: cells of a tile's top row / left / right column with their W, NW, N, NE neighbours that lie
// in another tile (every cross-tile pair is seen from its later cell)
__global__ void __launch_bounds__(256)
k_lab_border(const uint8_t *__restrict__ mask, u32 *L, int ny, int nx, int conn8) {
  const int r0 = blockIdx.y * LT, c0 = blockIdx.x * LT;
  const int side = threadIdx.x >> 6, j = threadIdx.x & (LT - 1);
  int lr, lc;
  if (side == 0) { lr = 0; lc = j; }                       // top row
  else if (side == 1) { lr = j; lc = 0; if (j == 0) return; }        // left column (corner belongs to the top row)
  else if (side == 2) { lr = j; lc = LT - 1; if (j == 0) return; }   // right column: only its NE neighbour leaves the tile
  else return;
  const int r = r0 + lr, c = c0 + lc;
  if (r >= ny || c >= nx) return;
  const size_t i = (size_t)r * nx + c;
  if (!mask[i]) return;
  if (side == 2) {
    if (conn8 && r > 0 && c < nx - 1 && mask[i - nx + 1]) uf_union(L, (u32)i, (u32)(i - nx + 1));
    return;
  }
  if (side == 1) {
    if (c > 0 && mask[i - 1]) uf_union(L, (u32)i, (u32)(i - 1));
    if (conn8 && r > 0 && c > 0 && mask[i - nx - 1]) uf_union(L, (u32)i, (u32)(i - nx - 1));
    return;
  }
  // top row
  if (lc == 0 && c > 0 && mask[i - 1]) uf_union(L, (u32)i, (u32)(i - 1));
  if (r > 0) {
    if (mask[i - nx]) uf_union(L, (u32)i, (u32)(i - nx));
    if (conn8) {
      if (c > 0 && mask[i - nx - 1]) uf_union(L, (u32)i, (u32)(i - nx - 1));
      if (c < nx - 1 && mask[i - nx + 1]) uf_union(L, (u32)i, (u32)(i - nx + 1));
    }
  }
}

__global__ void __launch_bounds__(256)
k_lab_compress(u32 *L, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u32 p = L[i];
  if (p == HC_NONE) return;
  L[i] = uf_find((volatile u32 *)L, p);
}

// ---- rank the roots: exclusive scan of (L[i] == i) ------------------------------------------------
#define SCAN_CHUNK 4096  // cells per block

__global__ void __launch_bounds__(256)
k_root_count(const u32 *__restrict__ L, u32 *__restrict__ bsum, size_t n) {
  __shared__ u32 s;
  if (threadIdx.x == 0) s = 0;
  __syncthreads();
  size_t base = (size_t)blockIdx.x * SCAN_CHUNK;
  u32 cnt = 0;
  for (int j = threadIdx.x; j < SCAN_CHUNK; j += 256) {
    size_t i = base + j;
    if (i < n && L[i] == (u32)i) cnt++;
  }
  atomicAdd(&s, cnt);
  __syncthreads();
  if (threadIdx.x == 0) bsum[blockIdx.x] = s;
}

__global__ void __launch_bounds__(1024)
k_scan_blocks(u32 *bsum, u32 nb, u32 *total) {
  // single block, sequential over chunks of 1024 with a running carry
  __shared__ u32 sh[1024];
  __shared__ u32 carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (u32 base = 0; base < nb; base += 1024) {
    u32 i = base + threadIdx.x;
    u32 v = i < nb ? bsum[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
      u32 t = threadIdx.x >= (u32)off ? sh[threadIdx.x - off] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    u32 incl = sh[threadIdx.x];
    if (i < nb) bsum[i] = carry + incl - v;  // exclusive
    __syncthreads();
    if (threadIdx.x == 1023) carry += incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(256)
k_root_rank(const u32 *__restrict__ L, const u32 *__restrict__ bsum, u32 *__restrict__ R, size_t n) {
  // in-block exclusive scan of root flags, 16 consecutive cells per thread
  __shared__ u32 sh[256];
  size_t base = (size_t)blockIdx.x * SCAN_CHUNK + (size_t)threadIdx.x * 16;
  u32 cnt = 0;
  u32 flags = 0;
  for (int j = 0; j < 16; j++) {
    size_t i = base + j;
    if (i < n && L[i] == (u32)i) { cnt++; flags |= 1u << j; }
  }
  sh[threadIdx.x] = cnt;
  __syncthreads();
  for (int off = 1; off < 256; off <<= 1) {
    u32 t = threadIdx.x >= (u32)off ? sh[threadIdx.x - off] : 0;
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  u32 run = bsum[blockIdx.x] + sh[threadIdx.x] - cnt;
  for (int j = 0; j < 16; j++)
    if (flags & (1u << j)) R[base + j] = ++run;  // 1-based id
}

__global__ void __launch_bounds__(256)
k_lab_out(const u32 *__restrict__ L, const u32 *__restrict__ R, int32_t *__restrict__ lab, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    u32 p = L[i];
    if (p != HC_NONE) {
      // (L is not flattened for this caller: a cell points at its tile-local root, which the border pass hung under
      //  the component's first cell -- one or two hops)
      u32 q = __ldg(&L[p]);
      while (q != p) { p = q; q = __ldg(&L[p]); }
    }
    lab[i] = p == HC_NONE ? 0 : (int32_t)__ldg(&R[p]);
  }
}

size_t label_workspace(int64_t ny, int64_t nx) {
  size_t n = (size_t)ny * nx;
  return 2 * hc_align(n * 4) + hc_align(((n + SCAN_CHUNK - 1) / SCAN_CHUNK) * 4) + hc_align(256);
}

static int label_roots(hc_ctx *ctx, const uint8_t *mask, u32 *L, int64_t ny, int64_t nx, int conn, cudaStream_t st,
                       bool flatten = true) {
  size_t n = (size_t)ny * nx;
  unsigned nb = (unsigned)((n + 255) / 256);
  dim3 tg((unsigned)((nx + LT - 1) / LT), (unsigned)((ny + LT - 1) / LT));
  HC_PROF(ctx, st, "k_lab_tile");
  k_lab_tile<<<tg, 256, 0, st>>>(mask, L, (int)ny, (int)nx, conn == 8);
  HC_LAUNCH_CHECK(ctx);
  HC_PROF(ctx, st, "k_lab_border");
  k_lab_border<<<tg, 256, 0, st>>>(mask, L, (int)ny, (int)nx, conn == 8);
  HC_LAUNCH_CHECK(ctx);
  if (flatten) {
    HC_PROF(ctx, st, "k_lab_compress");
    k_lab_compress<<<nb, 256, 0, st>>>(L, n);
    HC_LAUNCH_CHECK(ctx);
  }
  return HC_OK;
}

int label_run(hc_ctx *ctx, const uint8_t *mask, int32_t *lab, int64_t ny, int64_t nx, int conn, int32_t *nlab,
              cudaStream_t st) {
  if (nlab) *nlab = 0;
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (conn != 4 && conn != 8) { hc_set_error("label: conn must be 4 or 8"); return HC_ERR_ARG; }
  size_t n = (size_t)ny * nx;
  if (n >= 0x7FFFFFF0ull) { hc_set_error("raster too large"); return HC_ERR_TOO_LARGE; }
  HC_TRY(arena_reset(ctx));
  u32 nchunks = (u32)((n + SCAN_CHUNK - 1) / SCAN_CHUNK);
  u32 *L = (u32 *)arena_take(ctx, n * 4);
  u32 *R = (u32 *)arena_take(ctx, n * 4);
  u32 *bsum = (u32 *)arena_take(ctx, (size_t)nchunks * 4);
  u32 *total = (u32 *)arena_take(ctx, 256);
  if (!L || !R || !bsum || !total) return HC_ERR_NOMEM;
  HC_TRY(label_roots(ctx, mask, L, ny, nx, conn, st, false));   // (k_lab_out chases the one or two remaining hops itself)
  HC_PROF(ctx, st, "k_root_count");
  k_root_count<<<nchunks, 256, 0, st>>>(L, bsum, n);
  HC_LAUNCH_CHECK(ctx);
  HC_PROF(ctx, st, "k_scan_blocks");
  k_scan_blocks<<<1, 1024, 0, st>>>(bsum, nchunks, total);
  HC_LAUNCH_CHECK(ctx);
  HC_PROF(ctx, st, "k_root_rank");
  k_root_rank<<<nchunks, 256, 0, st>>>(L, bsum, R, n);
  HC_LAUNCH_CHECK(ctx);
  HC_PROF(ctx, st, "k_lab_out");
  k_lab_out<<<ctx->sm_count * 8, 256, 0, st>>>(L, R, lab, n);
  HC_LAUNCH_CHECK(ctx);
  HC_CUDA(cudaMemcpyAsync(ctx->h_mail, total, 4, cudaMemcpyDeviceToHost, st));
  HC_CUDA(cudaStreamSynchronize(st));
  if (nlab) *nlab = (int32_t)ctx->h_mail[0];
  return HC_OK;
}

// ---- exterior nodata (WBT seed rule) ------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_nodata_mask(const float *__restrict__ z, uint8_t *__restrict__ m, int ny, int nx, float nodata, u32 *interior, int last_col) {
  size_t n = (size_t)ny * nx;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  u32 any = 0;
  for (; i < n; i += stride) {
    float v = z[i];
    uint8_t nd = d_is_nodata(v, nodata);
    m[i] = nd;
    if (nd) {
      int r = (int)(i / (size_t)nx), c = (int)(i - (size_t)r * nx);
      if (r > 0 && c > 0 && r < ny - 1 && c < last_col) any = 1;   // (last_col: last real column; padding beyond it is exterior)
    }
  }
  if (any) *interior = 1u;
}

__global__ void __launch_bounds__(256)
k_ext_flag_border(const u32 *__restrict__ L, uint8_t *__restrict__ flag, int ny, int nx) {
  // threads over the 2*nx + 2*ny border cells
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int r, c;
  if (t < nx) { r = 0; c = t; }
  else if (t < 2 * nx) { r = ny - 1; c = t - nx; }
  else if (t < 2 * nx + ny) { r = t - 2 * nx; c = 0; }
  else if (t < 2 * nx + 2 * ny) { r = t - 2 * nx - ny; c = nx - 1; }
  else return;
  u32 p = L[(size_t)r * nx + c];
  if (p != HC_NONE) flag[p] = 1;
}

__global__ void __launch_bounds__(256)
k_ext_out(const u32 *__restrict__ L, const uint8_t *__restrict__ flag, uint8_t *__restrict__ ext, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    u32 p = L[i];
    ext[i] = p == HC_NONE ? 0 : flag[p];
  }
}

// ext[i] = 1 iff cell i is nodata and 8-connected through nodata to the raster border.
// ext must have been taken from the arena by the caller (no arena_reset here).
int ext_nodata_mask(hc_ctx *ctx, const float *z, uint8_t *ext, int64_t ny, int64_t nx, float nodata,
                    int *has_interior, cudaStream_t st) {
  size_t n = (size_t)ny * nx;
  u32 *flagw = (u32 *)arena_take(ctx, 256);
  if (!flagw) return HC_ERR_NOMEM;
  HC_CUDA(cudaMemsetAsync(flagw, 0, 4, st));
  HC_PROF(ctx, st, "k_nodata_mask");
  k_nodata_mask<<<ctx->sm_count * 8, 256, 0, st>>>(z, ext, (int)ny, (int)nx, nodata, flagw, (ctx->pad_from > 0 ? ctx->pad_from : (int)nx) - 1);
  HC_LAUNCH_CHECK(ctx);
  HC_CUDA(cudaMemcpyAsync(ctx->h_mail, flagw, 4, cudaMemcpyDeviceToHost, st));
  HC_CUDA(cudaStreamSynchronize(st));
  *has_interior = ctx->h_mail[0] != 0;
  if (!*has_interior) return HC_OK;
  u32 *L = (u32 *)arena_take(ctx, n * 4);
  uint8_t *flag = (uint8_t *)arena_take(ctx, n);
  if (!L || !flag) return HC_ERR_NOMEM;
  HC_TRY(label_roots(ctx, ext, L, ny, nx, 8, st));
  HC_CUDA(cudaMemsetAsync(flag, 0, n, st));
  int nbord = (int)(2 * nx + 2 * ny);
  HC_PROF(ctx, st, "k_ext_flag_border");
  k_ext_flag_border<<<(nbord + 255) / 256, 256, 0, st>>>(L, flag, (int)ny, (int)nx);
  HC_LAUNCH_CHECK(ctx);
  HC_PROF(ctx, st, "k_ext_out");
  k_ext_out<<<ctx->sm_count * 8, 256, 0, st>>>(L, flag, ext, n);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

// =================================================================================================
// Row bands (SURVEY.md §8e "Labelling"): local union-find per band; a component's GLOBAL root is the
// smallest global cell index over all its parts.  Parts are joined by min-label propagation across the
// seams: every round each band publishes, for its seam cells, the current global root of their component
// and lowers its own components to the smallest root seen across the seam (one all_gather + one small
// kernel per round, until nothing changes anywhere: as many rounds as a component's path has seam
// crossings).  Ids then follow scipy's rule: rank of the global root among all roots in raster order =
// owner band's offset (all_gather of the bands' owned-root counts) + local rank; components rooted in
// another band look their id up in the (root, id) pairs the owners publish for their seam cells.
// =================================================================================================
__global__ void k_band_lab_groot_init(const u32 *__restrict__ L, u32 *__restrict__ G, u32 base, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) G[i] = (L[i] == (u32)i) ? base + (u32)i : HC_NONE;   // defined at local roots only
}
// out[side][c] = global root of the seam cell's component, or NONE
__global__ void k_band_lab_seams(const u32 *__restrict__ L, const u32 *__restrict__ G, int ny, int nx, u32 *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * nx) return;
  int side = i / nx, c = i - side * nx;
  u32 p = L[(size_t)(side ? ny - 1 : 0) * nx + c];
  out[i] = p == HC_NONE ? HC_NONE : G[p];
}
// lower own components to the smallest root seen across the seams; *changed != 0 if anything moved
__global__ void k_band_lab_relax(const u32 *__restrict__ L, u32 *G, int ny, int nx, const u32 *__restrict__ all, int band,
                                 int nbands, int conn8, u32 *changed) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * nx) return;
  int side = i / nx, c = i - side * nx;
  int ob = side ? band + 1 : band - 1;   // the band across this seam
  if (ob < 0 || ob >= nbands) return;
  u32 p = L[(size_t)(side ? ny - 1 : 0) * nx + c];
  if (p == HC_NONE) return;
  const u32 *row = all + ((size_t)ob * 2 + (side ? 0 : 1)) * nx;   // its facing seam row
  u32 m = HC_NONE;
  for (int d = conn8 ? -1 : 0; d <= (conn8 ? 1 : 0); d++) {
    int cc = c + d;
    if (cc < 0 || cc >= nx) continue;
    u32 g = row[cc];
    if (g < m) m = g;
  }
  if (m != HC_NONE && m < G[p]) {
    if (atomicMin(&G[p], m) > m) *changed = 1u;
  }
}
// Lo[i] = i for local roots whose component's first cell lies in this band (owned), else NONE
__global__ void k_band_lab_owned(const u32 *__restrict__ L, const u32 *__restrict__ G, u32 base, u32 *__restrict__ Lo, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) Lo[i] = (L[i] == (u32)i && G[i] == base + (u32)i) ? (u32)i : HC_NONE;
}
// pairs[side][c] = (global root, id) for seam cells of OWNED components, (NONE, 0) otherwise
__global__ void k_band_lab_pairs(const u32 *__restrict__ L, const u32 *__restrict__ G, const u32 *__restrict__ R, u32 base,
                                 u32 offset, int ny, int nx, u32 *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * nx) return;
  int side = i / nx, c = i - side * nx;
  u32 p = L[(size_t)(side ? ny - 1 : 0) * nx + c];
  u32 g = HC_NONE, id = 0;
  if (p != HC_NONE && G[p] == base + p) { g = G[p]; id = offset + R[p]; }
  out[2 * (size_t)i] = g;
  out[2 * (size_t)i + 1] = id;
}
__global__ void __launch_bounds__(256)
k_band_lab_out(const u32 *__restrict__ L, const u32 *__restrict__ G, const u32 *__restrict__ R, u32 base, u32 offset,
               const u32 *__restrict__ keys, const u32 *__restrict__ vals, int nkeys, int32_t *__restrict__ lab, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    u32 p = L[i];
    int32_t id = 0;
    if (p != HC_NONE) {
      u32 g = G[p];
      if (g == base + p) id = (int32_t)(offset + __ldg(&R[p]));
      else {
        int lo = 0, hi = nkeys - 1;   // sorted (root, id) pairs published by the owners
        while (lo < hi) { int mid = (lo + hi) >> 1; if (keys[mid] < g) lo = mid + 1; else hi = mid; }
        id = (nkeys > 0 && keys[lo] == g) ? (int32_t)vals[lo] : -1;
      }
    }
    lab[i] = id;
  }
}


// local: union-find of the band's own rows; seam_out[2][nx] = global roots of the seam cells; *n_owned_guess unused
extern "C" int hc_band_label_local_dev(hc_ctx *ctx, const uint8_t *mask, int64_t nb, int64_t nx, int conn, int64_t row0,
                                       uint32_t *seam_out, void *stream) {
  if (!ctx) { hc_set_error("null context"); return HC_ERR_ARG; }
  hc_guard hc_guard__(ctx);
  HC_CUDA(cudaSetDevice(ctx->device));
  if (!mask || !seam_out || nb <= 0 || nx <= 0 || (conn != 4 && conn != 8)) { hc_set_error("hc_band_label_local: bad argument"); return HC_ERR_ARG; }
  if ((double)(row0 + nb) * (double)nx >= 2147483632.0) { hc_set_error("labelled raster exceeds the 31-bit cell index"); return HC_ERR_TOO_LARGE; }
  cudaStream_t st = (cudaStream_t)stream;
  size_t n = (size_t)nb * nx;
  HC_TRY(arena_reset(ctx));
  band_label_state *s = &ctx->band_label;
  s->L = (u32 *)arena_take(ctx, n * 4);
  s->G = (u32 *)arena_take(ctx, n * 4);
  s->R = (u32 *)arena_take(ctx, n * 4);
  s->flag = (u32 *)arena_take(ctx, 256);
  if (!s->L || !s->G || !s->R || !s->flag) return HC_ERR_NOMEM;
  s->nb = nb; s->nx = nx; s->conn = conn; s->base = (u32)(row0 * nx); s->ready = 1;
  HC_TRY(label_roots(ctx, mask, s->L, nb, nx, conn, st));
  unsigned gb = (unsigned)((n + 255) / 256);
  k_band_lab_groot_init<<<gb, 256, 0, st>>>(s->L, s->G, s->base, n);
  HC_LAUNCH_CHECK(ctx);
  k_band_lab_seams<<<(unsigned)((2 * nx + 255) / 256), 256, 0, st>>>(s->L, s->G, (int)nb, (int)nx, seam_out);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}
// one propagation round: takes every band's seam record, lowers own components, republishes; *changed (host)
extern "C" int hc_band_label_relax_dev(hc_ctx *ctx, const uint32_t *all_seams, int64_t nbands, int64_t band_index,
                                       uint32_t *seam_out, int *changed, void *stream) {
  if (!ctx) { hc_set_error("null context"); return HC_ERR_ARG; }
  hc_guard hc_guard__(ctx);
  HC_CUDA(cudaSetDevice(ctx->device));
  band_label_state *s = &ctx->band_label;
  if (!s->ready || !all_seams || !seam_out) { hc_set_error("hc_band_label_relax: local phase first"); return HC_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  HC_CUDA(cudaMemsetAsync(s->flag, 0, 4, st));
  unsigned gs = (unsigned)((2 * s->nx + 255) / 256);
  k_band_lab_relax<<<gs, 256, 0, st>>>(s->L, s->G, (int)s->nb, (int)s->nx, all_seams, (int)band_index, (int)nbands, s->conn == 8, s->flag);
  HC_LAUNCH_CHECK(ctx);
  k_band_lab_seams<<<gs, 256, 0, st>>>(s->L, s->G, (int)s->nb, (int)s->nx, seam_out);
  HC_LAUNCH_CHECK(ctx);
  HC_CUDA(cudaMemcpyAsync(ctx->h_mail, s->flag, 4, cudaMemcpyDeviceToHost, st));
  HC_CUDA(cudaStreamSynchronize(st));
  if (changed) *changed = ctx->h_mail[0] ? 1 : 0;
  return HC_OK;
}
// rank the owned roots locally; *n_owned (host) = components whose first cell lies in this band
extern "C" int hc_band_label_rank_dev(hc_ctx *ctx, int64_t *n_owned, void *stream) {
  if (!ctx) { hc_set_error("null context"); return HC_ERR_ARG; }
  hc_guard hc_guard__(ctx);
  HC_CUDA(cudaSetDevice(ctx->device));
  band_label_state *s = &ctx->band_label;
  if (!s->ready) { hc_set_error("hc_band_label_rank: local phase first"); return HC_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  size_t n = (size_t)s->nb * s->nx;
  u32 nchunks = (u32)((n + SCAN_CHUNK - 1) / SCAN_CHUNK);
  u32 *Lo = (u32 *)arena_take(ctx, n * 4);
  u32 *bsum = (u32 *)arena_take(ctx, (size_t)nchunks * 4);
  if (!Lo || !bsum) return HC_ERR_NOMEM;
  k_band_lab_owned<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s->L, s->G, s->base, Lo, n);
  HC_LAUNCH_CHECK(ctx);
  k_root_count<<<nchunks, 256, 0, st>>>(Lo, bsum, n);
  HC_LAUNCH_CHECK(ctx);
  k_scan_blocks<<<1, 1024, 0, st>>>(bsum, nchunks, s->flag + 1);
  HC_LAUNCH_CHECK(ctx);
  k_root_rank<<<nchunks, 256, 0, st>>>(Lo, bsum, s->R, n);
  HC_LAUNCH_CHECK(ctx);
  HC_CUDA(cudaMemcpyAsync(ctx->h_mail, s->flag + 1, 4, cudaMemcpyDeviceToHost, st));
  HC_CUDA(cudaStreamSynchronize(st));
  s->owned = ctx->h_mail[0];
  if (n_owned) *n_owned = s->owned;
  return HC_OK;
}
// pairs_out[2][nx][2]: (global root, id) of the seam cells of owned components (id = offset + local rank)
extern "C" int hc_band_label_pairs_dev(hc_ctx *ctx, int64_t offset, uint32_t *pairs_out, void *stream) {
  if (!ctx) { hc_set_error("null context"); return HC_ERR_ARG; }
  hc_guard hc_guard__(ctx);
  HC_CUDA(cudaSetDevice(ctx->device));
  band_label_state *s = &ctx->band_label;
  if (!s->ready || !pairs_out) { hc_set_error("hc_band_label_pairs: rank phase first"); return HC_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  k_band_lab_pairs<<<(unsigned)((2 * s->nx + 255) / 256), 256, 0, st>>>(s->L, s->G, s->R, s->base, (u32)offset, (int)s->nb, (int)s->nx, pairs_out);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}
// final ids: keys/vals = every band's published (root, id) pairs with id > 0, sorted by root (device pointers)
extern "C" int hc_band_label_finish_dev(hc_ctx *ctx, int64_t offset, const uint32_t *keys, const uint32_t *vals,
                                        int64_t nkeys, int32_t *lab, void *stream) {
  if (!ctx) { hc_set_error("null context"); return HC_ERR_ARG; }
  hc_guard hc_guard__(ctx);
  HC_CUDA(cudaSetDevice(ctx->device));
  band_label_state *s = &ctx->band_label;
  if (!s->ready || !lab) { hc_set_error("hc_band_label_finish: rank phase first"); return HC_ERR_ARG; }
  s->ready = 0;
  cudaStream_t st = (cudaStream_t)stream;
  size_t n = (size_t)s->nb * s->nx;
  k_band_lab_out<<<ctx->sm_count * 8, 256, 0, st>>>(s->L, s->G, s->R, s->base, (u32)offset, keys, vals, (int)nkeys, lab, n);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

// out[i] = flag[lab[i]] (flag[0] must be 0: background): the exterior-nodata raster from banded labels
__global__ void k_gather_flag(const int32_t *__restrict__ lab, const uint8_t *__restrict__ flag, uint8_t *__restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = flag[lab[i]];
}
extern "C" int hc_gather_flag_dev(hc_ctx *ctx, const int32_t *lab, const uint8_t *flag, uint8_t *out, int64_t n, void *stream) {
  if (!ctx) { hc_set_error("null context"); return HC_ERR_ARG; }
  hc_guard hc_guard__(ctx);
  HC_CUDA(cudaSetDevice(ctx->device));
  if (n <= 0) return HC_OK;
  if (!lab || !flag || !out) { hc_set_error("hc_gather_flag: null pointer"); return HC_ERR_ARG; }
  k_gather_flag<<<ctx->sm_count * 8, 256, 0, (cudaStream_t)stream>>>(lab, flag, out, (size_t)n);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}
