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
// so they are computed by chaotic relaxation: each 64x64 tile iterates to its local fixed point in
// shared memory, tiles whose halo changed are re-queued, until nothing changes.  No flat labels
// are needed: two 8-adjacent cells of equal elevation are always in the same flat, FlatHeight is
// a per-flat constant that cancels in every comparison the masked D8 makes, and a flat without a
// low edge is exactly one whose cells the "towards lower" field never reaches.
#include "common.cuh"

#define T HC_TILE
#define TH (T + 2)
#define NTHREADS 256

// ---- classify ---------------------------------------------------------------------------------
// nf[c]  = 1 if c is valid and has no direction (NOFLOW)
// dlo[c] = 1 for low edges (valid, has a direction, touches an equal-elevation NOFLOW cell)
// dhi[c] = 1 for high edges (NOFLOW, touches a higher valid cell); 0 = not reached yet
__global__ void __launch_bounds__(NTHREADS)
k_flat_init(const float *__restrict__ z, const uint8_t *__restrict__ dir, uint8_t *__restrict__ nf,
            u32 *__restrict__ dlo, u32 *__restrict__ dhi, uint8_t *__restrict__ tile_has, int ny, int nx,
            float nodata) {
  __shared__ float sz[TH * TH];
  __shared__ uint8_t sd[TH * TH];
  const int r0 = blockIdx.y * T, c0 = blockIdx.x * T;
  const int tid = threadIdx.x;
  const float qnan = __int_as_float(0x7fc00000);
  for (int i = tid; i < TH * TH; i += NTHREADS) {
    int lr = i / TH, lc = i - lr * TH;
    int r = r0 + lr - 1, c = c0 + lc - 1;
    float v = qnan;
    uint8_t d = HC_DIR_NODATA;
    if (r >= 0 && r < ny && c >= 0 && c < nx) {
      size_t g = (size_t)r * nx + c;
      v = __ldg(&z[g]);
      if (v == nodata) v = qnan;
      d = __ldg(&dir[g]);
    }
    sz[i] = v;
    sd[i] = d;
  }
  __syncthreads();
  int any = 0;
  for (int i = tid; i < T * T; i += NTHREADS) {
    int lr = i / T, lc = i - lr * T;
    int r = r0 + lr, c = c0 + lc;
    if (r >= ny || c >= nx) continue;
    int ci = (lr + 1) * TH + lc + 1;
    float v = sz[ci];
    uint8_t isnf = 0;
    u32 lo = 0, hi = 0;
    if (v == v) {
      isnf = sd[ci] == HC_DIR_NOFLOW;
#pragma unroll
      for (int dr = -1; dr <= 1; dr++)
#pragma unroll
        for (int dc = -1; dc <= 1; dc++) {
          if (!dr && !dc) continue;
          int ni = ci + dr * TH + dc;
          float zn = sz[ni];
          if (zn != zn) continue;
          if (isnf) { if (zn > v) hi = 1; }
          else { if (zn == v && sd[ni] == HC_DIR_NOFLOW) lo = 1; }
        }
    }
    size_t g = (size_t)r * nx + c;
    nf[g] = isnf;
    dlo[g] = lo;
    dhi[g] = hi;
    any |= isnf;
  }
  any = __syncthreads_or(any);
  if (tid == 0) tile_has[blockIdx.y * gridDim.x + blockIdx.x] = (uint8_t)(any ? 1 : 0);
}

// ---- tile relaxation ---------------------------------------------------------------------------
// dynamic smem: z[TH*TH] f32 | lo[TH*TH] u32 | hi[TH*TH] u32 | nf[TH*TH] u8
#define FLAT_SMEM (TH * TH * (4 + 4 + 4 + 1))

__global__ void __launch_bounds__(NTHREADS)
k_flat_relax(const float *__restrict__ z, const uint8_t *__restrict__ nf, u32 *dlo, u32 *dhi,
             const uint8_t *__restrict__ tile_has, const uint8_t *__restrict__ dirty_in,
             uint8_t *__restrict__ dirty_out, u32 *__restrict__ changed_flag, int ny, int nx, float nodata) {
  extern __shared__ __align__(16) unsigned char smem[];
  float *sz = (float *)smem;
  u32 *slo = (u32 *)(smem + TH * TH * 4);
  u32 *shi = (u32 *)(smem + TH * TH * 8);
  uint8_t *snf = (uint8_t *)(smem + TH * TH * 12);
  const int tile = blockIdx.y * gridDim.x + blockIdx.x;
  if (!tile_has[tile] || !dirty_in[tile]) return;
  const int r0 = blockIdx.y * T, c0 = blockIdx.x * T;
  const int tid = threadIdx.x;
  const float qnan = __int_as_float(0x7fc00000);
  for (int i = tid; i < TH * TH; i += NTHREADS) {
    int lr = i / TH, lc = i - lr * TH;
    int r = r0 + lr - 1, c = c0 + lc - 1;
    float v = qnan;
    u32 lo = 0, hi = 0;
    uint8_t f = 0;
    if (r >= 0 && r < ny && c >= 0 && c < nx) {
      size_t g = (size_t)r * nx + c;
      v = __ldg(&z[g]);
      if (v == nodata) v = qnan;
      lo = __ldcg(&dlo[g]);
      hi = __ldcg(&dhi[g]);
      f = __ldg(&nf[g]);
    }
    sz[i] = v; slo[i] = lo; shi[i] = hi; snf[i] = f;
  }
  __syncthreads();
  int ever = 0;
  for (;;) {
    int changed = 0;
    for (int i = tid; i < T * T; i += NTHREADS) {
      int lr = i / T, lc = i - lr * T;
      int ci = (lr + 1) * TH + lc + 1;
      if (!snf[ci]) continue;
      float v = sz[ci];
      u32 lo = slo[ci], hi = shi[ci];
      u32 blo = lo ? lo : 0xFFFFFFFFu, bhi = hi ? hi : 0xFFFFFFFFu;
#pragma unroll
      for (int dr = -1; dr <= 1; dr++)
#pragma unroll
        for (int dc = -1; dc <= 1; dc++) {
          if (!dr && !dc) continue;
          int ni = ci + dr * TH + dc;
          if (sz[ni] != v) continue;
          u32 nlo = slo[ni];
          if (nlo && nlo + 1 < blo) blo = nlo + 1;
          u32 nhi = shi[ni];
          if (nhi && snf[ni] && nhi + 1 < bhi) bhi = nhi + 1;
        }
      if (blo != 0xFFFFFFFFu && blo != lo) { slo[ci] = blo; changed = 1; }
      if (bhi != 0xFFFFFFFFu && bhi != hi) { shi[ci] = bhi; changed = 1; }
    }
    if (!__syncthreads_or(changed)) break;
    ever = 1;
  }
  if (!ever) return;
  for (int i = tid; i < T * T; i += NTHREADS) {
    int lr = i / T, lc = i - lr * T;
    int r = r0 + lr, c = c0 + lc;
    if (r >= ny || c >= nx) continue;
    int ci = (lr + 1) * TH + lc + 1;
    if (!snf[ci]) continue;
    size_t g = (size_t)r * nx + c;
    dlo[g] = slo[ci];
    dhi[g] = shi[ci];
  }
  if (tid < 9) {
    int ty = (int)blockIdx.y + tid / 3 - 1, tx = (int)blockIdx.x + tid % 3 - 1;
    if (ty >= 0 && ty < (int)gridDim.y && tx >= 0 && tx < (int)gridDim.x) dirty_out[ty * gridDim.x + tx] = 1;
  }
  if (tid == 0) *changed_flag = 1u;
}

// ---- masked D8 ---------------------------------------------------------------------------------
template <int TABLE>
__global__ void __launch_bounds__(256)
k_flat_dirs(const float *__restrict__ z, const uint8_t *__restrict__ nf, const u32 *__restrict__ dlo,
            const u32 *__restrict__ dhi, uint8_t *__restrict__ dir, const uint8_t *__restrict__ tile_has,
            int ny, int nx, float nodata) {
  // one block per 64x64 tile, 16 cells per thread; direct (cached) neighbour reads: only the
  // sparse NOFLOW cells do any work
  const int tile = blockIdx.y * gridDim.x + blockIdx.x;
  if (!tile_has[tile]) return;
  const int r0 = blockIdx.y * T, c0 = blockIdx.x * T;
  for (int i = threadIdx.x; i < T * T; i += 256) {
    int lr = i / T, lc = i - lr * T;
    int r = r0 + lr, c = c0 + lc;
    if (r >= ny || c >= nx) continue;
    size_t g = (size_t)r * nx + c;
    if (!nf[g]) continue;
    u32 lo = dlo[g];
    if (!lo) continue;  // flat without an outlet: stays NOFLOW
    float v = z[g];
    long long own = 2ll * (long long)lo - (long long)dhi[g];
    long long mn = own;
    int best = -1;
    bool bdiag = false;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      int rr = r + c_dr[TABLE][k], cc = c + c_dc[TABLE][k];
      if (rr < 0 || rr >= ny || cc < 0 || cc >= nx) continue;
      size_t gn = (size_t)rr * nx + cc;
      float zn = z[gn];
      if (zn != v) continue;  // also drops nodata / NaN
      u32 nlo = dlo[gn];
      long long key;
      if (nf[gn]) {
        if (!nlo) continue;   // cannot happen inside a reached flat; keep it safe
        key = 2ll * (long long)nlo - (long long)dhi[gn];
      } else {
        key = -0x4000000000000000ll;  // low edge: below every mask value
      }
      bool kdiag = (c_dr[TABLE][k] != 0) && (c_dc[TABLE][k] != 0);
      if (key < mn || (key == mn && best >= 0 && bdiag && !kdiag)) { mn = key; best = k; bdiag = kdiag; }
    }
    if (best >= 0) dir[g] = (uint8_t)c_code[TABLE][best];
  }
}

size_t flats_workspace(int64_t ny, int64_t nx) {
  size_t n = (size_t)ny * nx;
  size_t tiles = (size_t)((ny + T - 1) / T) * ((nx + T - 1) / T);
  return hc_align(n) + 2 * hc_align(n * 4) + 3 * hc_align(tiles) + hc_align(256);
}

int flats_run(hc_ctx *ctx, const float *z, uint8_t *dir, int64_t ny, int64_t nx, float nodata, int table,
              cudaStream_t st) {
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (table != HC_TABLE_WBT && table != HC_TABLE_TAUDEM) { hc_set_error("flats: bad table %d", table); return HC_ERR_ARG; }
  size_t n = (size_t)ny * nx;
  const int tx = (int)((nx + T - 1) / T), ty = (int)((ny + T - 1) / T);
  const size_t tiles = (size_t)tx * ty;
  HC_TRY(arena_reset(ctx));
  uint8_t *nf = (uint8_t *)arena_take(ctx, n);
  u32 *dlo = (u32 *)arena_take(ctx, n * 4);
  u32 *dhi = (u32 *)arena_take(ctx, n * 4);
  uint8_t *tile_has = (uint8_t *)arena_take(ctx, tiles);
  uint8_t *dA = (uint8_t *)arena_take(ctx, tiles);
  uint8_t *dB = (uint8_t *)arena_take(ctx, tiles);
  u32 *flag = (u32 *)arena_take(ctx, 256);
  if (!nf || !dlo || !dhi || !tile_has || !dA || !dB || !flag) return HC_ERR_NOMEM;

  static bool attr_set = false;
  if (!attr_set) {
    HC_CUDA(cudaFuncSetAttribute(k_flat_relax, cudaFuncAttributeMaxDynamicSharedMemorySize, FLAT_SMEM));
    attr_set = true;
  }
  dim3 tg(tx, ty);
  HC_PROF(ctx, st, "k_flat_init");
  k_flat_init<<<tg, NTHREADS, 0, st>>>(z, dir, nf, dlo, dhi, tile_has, (int)ny, (int)nx, nodata);
  HC_LAUNCH_CHECK(ctx);
  HC_CUDA(cudaMemsetAsync(dA, 1, tiles, st));
  uint8_t *din = dA, *dout = dB;
  for (int it = 0;; it++) {
    if (it > 100000) { hc_set_error("flats: relaxation did not converge"); return HC_ERR_INTERNAL; }
    HC_CUDA(cudaMemsetAsync(dout, 0, tiles, st));
    HC_CUDA(cudaMemsetAsync(flag, 0, 4, st));
    HC_PROF(ctx, st, "k_flat_relax");
    k_flat_relax<<<tg, NTHREADS, FLAT_SMEM, st>>>(z, nf, dlo, dhi, tile_has, din, dout, flag, (int)ny, (int)nx, nodata);
    HC_LAUNCH_CHECK(ctx);
    HC_CUDA(cudaMemcpyAsync(ctx->h_mail, flag, 4, cudaMemcpyDeviceToHost, st));
    HC_CUDA(cudaStreamSynchronize(st));
    if (!ctx->h_mail[0]) break;
    uint8_t *t = din; din = dout; dout = t;
  }
  HC_PROF(ctx, st, "k_flat_dirs");
  if (table == HC_TABLE_WBT)
    k_flat_dirs<HC_TABLE_WBT><<<tg, 256, 0, st>>>(z, nf, dlo, dhi, dir, tile_has, (int)ny, (int)nx, nodata);
  else
    k_flat_dirs<HC_TABLE_TAUDEM><<<tg, 256, 0, st>>>(z, nf, dlo, dhi, dir, tile_has, (int)ny, (int)nx, nodata);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}
