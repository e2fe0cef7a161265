// pits.cu — the small stencils around the fill: WhiteboxTools' single-cell-pit tools and the
// TaudemCheck problem mask.
//
//   hc_fill_single_cell_pits   wbt.fill_single_cell_pits   models/depression_handling.py:261,
//                                                           models/find_sinks.py:118
//   hc_breach_single_cell_pits wbt.breach_single_cell_pits models/depression_handling.py:259
//   hc_taudem_check            TaudemCheck.check_basin / fix_basin  models/taudem_check.py:143-147,183-194
//
// Semantics follow oracle/hydro_oracle.c (orc_fill_single_cell_pits / orc_breach_single_cell_pits):
// both tools read the INPUT raster only.  The serial breach writes bridge cells in raster order of
// the pits (and ring order within a pit), the last write winning; here every cell instead GATHERS
// the pits that may cut it and keeps the last one in that same order, so the result is identical
// and race free.
#include "common.cuh"

__device__ __forceinline__ float ldz(const float *z, int r, int c, int ny, int nx, float nodata) {
  if (r < 0 || r >= ny || c < 0 || c >= nx) return __int_as_float(0x7fc00000);
  float v = __ldg(&z[(size_t)r * nx + c]);
  return v == nodata ? __int_as_float(0x7fc00000) : v;
}

// is (r,c) an interior cell whose 8 neighbours are all valid and strictly higher?  *mn = lowest neighbour
__device__ __forceinline__ bool is_single_pit(const float *z, int r, int c, int ny, int nx, float nodata,
                                              int margin, float *vout, float *mn) {
  if (r < margin || r >= ny - margin || c < margin || c >= nx - margin) return false;
  float v = ldz(z, r, c, ny, nx, nodata);
  if (v != v) return false;
  float m = INFINITY;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    float zn = ldz(z, r + c_dr[0][k], c + c_dc[0][k], ny, nx, nodata);
    if (!(zn > v)) return false;  // also false for NaN (nodata)
    m = fminf(m, zn);
  }
  *vout = v;
  *mn = m;
  return true;
}

__global__ void __launch_bounds__(256)
k_fill_scp(const float *__restrict__ z, float *__restrict__ out, int ny, int nx, float nodata) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  int r = blockIdx.y;
  if (c >= nx) return;
  size_t g = (size_t)r * nx + c;
  float v, mn;
  out[g] = is_single_pit(z, r, c, ny, nx, nodata, 1, &v, &mn) ? mn : z[g];
}

// ring-2 offsets in the serial tool's order and the 1-ring slot (WBT table) each one is bridged through
__constant__ const int c_dc2[16] = {2, 2, 2, 2, 2, 1, 0, -1, -2, -2, -2, -2, -2, -1, 0, 1};
__constant__ const int c_dr2[16] = {-2, -1, 0, 1, 2, 2, 2, 2, 2, 1, 0, -1, -2, -2, -2, -2};
__constant__ const int c_bridge[16] = {0, 1, 1, 1, 2, 3, 3, 3, 4, 5, 5, 5, 6, 7, 7, 7};

__global__ void __launch_bounds__(256)
k_breach_scp(const float *__restrict__ z, float *__restrict__ out, int ny, int nx, float nodata) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  int r = blockIdx.y;
  if (c >= nx) return;
  size_t g = (size_t)r * nx + c;
  float res = z[g];
  // candidate pits = my 8 neighbours, visited in raster order so that the last writer wins
#pragma unroll
  for (int dr = -1; dr <= 1; dr++)
#pragma unroll
    for (int dc = -1; dc <= 1; dc++) {
      if (!dr && !dc) continue;
      int pr = r + dr, pc = c + dc;
      float v, mn;
      if (!is_single_pit(z, pr, pc, ny, nx, nodata, 2, &v, &mn)) continue;
      // which 1-ring slot of the pit am I?  (pit + slot offset = me)
      int slot = -1;
#pragma unroll
      for (int s = 0; s < 8; s++)
        if (c_dr[0][s] == -dr && c_dc[0][s] == -dc) slot = s;
      for (int k = 0; k < 16; k++) {
        if (c_bridge[k] != slot) continue;
        float z2 = ldz(z, pr + c_dr2[k], pc + c_dc2[k], ny, nx, nodata);
        if (z2 < v) res = (float)(((double)v + (double)z2) / 2.0);
      }
    }
  out[g] = res;
}

// problem cells of TaudemCheck: D8 is nodata on a valid, near-shore water cell
__global__ void __launch_bounds__(256)
k_taudem_check(const uint8_t *__restrict__ p, const float *__restrict__ fel, const uint8_t *__restrict__ flat,
               const uint8_t *__restrict__ edm, uint8_t *__restrict__ mask, u32 *__restrict__ count, size_t n,
               float nodata) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  u32 local = 0;
  for (; i < n; i += stride) {
    float v = fel[i];
    uint8_t f = flat[i];
    uint8_t m = p[i] == HC_DIR_NODATA && !d_is_nodata(v, nodata) && f > 1 && f < 6 && edm[i] < 4;
    mask[i] = m;
    local += m;
  }
  local = __reduce_add_sync(0xffffffffu, local);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
}

// fix_basin: raise every water cell (flat mask > 0) by `bump` metres, nodata untouched
__global__ void __launch_bounds__(256)
k_taudem_fix(const float *__restrict__ dem, const uint8_t *__restrict__ flat, float *__restrict__ out, size_t n,
             float nodata, float bump) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float v = dem[i];
    out[i] = (!d_is_nodata(v, nodata) && flat[i] > 0) ? __fadd_rn(v, bump) : v;
  }
}

int scp_run(hc_ctx *ctx, const float *z, float *out, int64_t ny, int64_t nx, float nodata, int breach, cudaStream_t st) {
  if (ny <= 0 || nx <= 0) return HC_OK;
  dim3 grid((unsigned)((nx + 255) / 256), (unsigned)ny);
  if (ny > 65535) { hc_set_error("single-cell-pit tools: more than 65535 rows"); return HC_ERR_TOO_LARGE; }
  if (breach) {
    HC_PROF(ctx, st, "k_breach_scp");
    k_breach_scp<<<grid, 256, 0, st>>>(z, out, (int)ny, (int)nx, nodata);
  } else {
    HC_PROF(ctx, st, "k_fill_scp");
    k_fill_scp<<<grid, 256, 0, st>>>(z, out, (int)ny, (int)nx, nodata);
  }
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

int taudem_check_run(hc_ctx *ctx, const uint8_t *p, const float *fel, const uint8_t *flat, const uint8_t *edm,
                     uint8_t *mask, int64_t *n_problem, int64_t ny, int64_t nx, float nodata, cudaStream_t st) {
  if (n_problem) *n_problem = 0;
  if (ny <= 0 || nx <= 0) return HC_OK;
  size_t n = (size_t)ny * nx;
  HC_TRY(arena_reset(ctx));
  u32 *cnt = (u32 *)arena_take(ctx, 256);
  if (!cnt) return HC_ERR_NOMEM;
  HC_CUDA(cudaMemsetAsync(cnt, 0, 4, st));
  HC_PROF(ctx, st, "k_taudem_check");
  k_taudem_check<<<ctx->sm_count * 8, 256, 0, st>>>(p, fel, flat, edm, mask, cnt, n, nodata);
  HC_LAUNCH_CHECK(ctx);
  HC_CUDA(cudaMemcpyAsync(ctx->h_mail, cnt, 4, cudaMemcpyDeviceToHost, st));
  HC_CUDA(cudaStreamSynchronize(st));
  if (n_problem) *n_problem = ctx->h_mail[0];
  return HC_OK;
}

int taudem_fix_run(hc_ctx *ctx, const float *dem, const uint8_t *flat, float *out, int64_t ny, int64_t nx,
                   float nodata, float bump, cudaStream_t st) {
  if (ny <= 0 || nx <= 0) return HC_OK;
  HC_PROF(ctx, st, "k_taudem_fix");
  k_taudem_fix<<<ctx->sm_count * 8, 256, 0, st>>>(dem, flat, out, (size_t)ny * nx, nodata, bump);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

// ---- C ABI -----------------------------------------------------------------------------------------
#define PITS_CTX(ctx) HC_ENTER(ctx)

extern "C" int hc_fill_single_cell_pits_dev(hc_ctx *ctx, const float *z, float *out, int64_t ny, int64_t nx,
                                            float nodata, void *stream) {
  PITS_CTX(ctx);
  if (ny < 0 || nx < 0 || (ny && nx && (!z || !out || z == out))) { hc_set_error("hc_fill_single_cell_pits: bad argument"); return HC_ERR_ARG; }
  return scp_run(ctx, z, out, ny, nx, nodata, 0, (cudaStream_t)stream);
}

extern "C" int hc_breach_single_cell_pits_dev(hc_ctx *ctx, const float *z, float *out, int64_t ny, int64_t nx,
                                              float nodata, void *stream) {
  PITS_CTX(ctx);
  if (ny < 0 || nx < 0 || (ny && nx && (!z || !out || z == out))) { hc_set_error("hc_breach_single_cell_pits: bad argument"); return HC_ERR_ARG; }
  return scp_run(ctx, z, out, ny, nx, nodata, 1, (cudaStream_t)stream);
}

extern "C" int hc_taudem_check_dev(hc_ctx *ctx, const uint8_t *p, const float *fel, const uint8_t *flat,
                                   const uint8_t *edm, uint8_t *mask, int64_t *n_problem, int64_t ny, int64_t nx,
                                   float nodata, void *stream) {
  PITS_CTX(ctx);
  if (ny < 0 || nx < 0 || (ny && nx && (!p || !fel || !flat || !edm || !mask))) { hc_set_error("hc_taudem_check: bad argument"); return HC_ERR_ARG; }
  return taudem_check_run(ctx, p, fel, flat, edm, mask, n_problem, ny, nx, nodata, (cudaStream_t)stream);
}

extern "C" int hc_taudem_fix_dev(hc_ctx *ctx, const float *dem, const uint8_t *flat, float *out, int64_t ny,
                                 int64_t nx, float nodata, float bump, void *stream) {
  PITS_CTX(ctx);
  if (ny < 0 || nx < 0 || (ny && nx && (!dem || !flat || !out))) { hc_set_error("hc_taudem_fix: bad argument"); return HC_ERR_ARG; }
  return taudem_fix_run(ctx, dem, flat, out, ny, nx, nodata, bump, (cudaStream_t)stream);
}

// np.round(a - b, 3) on float32 rasters (tools/depressions.py:1308, :1395: the fill / carve difference maps): numpy
// rounds float32 to `decimals` places as rint(x * 10^d) / 10^d in float32, every step rounded.
__global__ void __launch_bounds__(256)
k_diff_round(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ out, size_t n, float scale) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = __fdiv_rn(rintf(__fmul_rn(__fsub_rn(a[i], b[i]), scale)), scale);
}

extern "C" int hc_diff_round_dev(hc_ctx *ctx, const float *a, const float *b, float *out, int64_t n, int decimals,
                                 void *stream) {
  HC_ENTER(ctx);
  if (n <= 0) return HC_OK;
  if (!a || !b || !out || decimals < 0 || decimals > 6) { hc_set_error("hc_diff_round: bad argument"); return HC_ERR_ARG; }
  float scale = 1.f;
  for (int k = 0; k < decimals; k++) scale *= 10.f;
  cudaStream_t st = (cudaStream_t)stream;
  HC_PROF(ctx, st, "k_diff_round");
  k_diff_round<<<ctx->sm_count * 8, 256, 0, st>>>(a, b, out, (size_t)n, scale);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}
