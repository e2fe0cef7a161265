// d8.cu — D8 steepest-descent direction + slope stencil.
//
// Replaces wbt.d8_pointer (run_aws.py:539) and the steepest-descent half of TauDEM D8FlowDir
// -p -sd8 (run_aws.py:594-603).  Arithmetic follows oracle/hydro_oracle.c orc_d8:
//   WBT table    slope_k = (double)(z - zk) / len_k      (IEEE double division)
//   TauDEM table slope_k = (float)(1/len_k) * (z - zk)   (two float32 roundings, no FMA)
// first strict maximum > 0 in the tool's own neighbour order; TauDEM marks border cells and cells
// with a nodata neighbour as contaminated (code 255).
#include "common.cuh"

#define D8_TR 16
#define D8_TC 128
#define D8_THREADS 256

struct d8_params {
  double len[8];
  float fact[8];
};

template <int TABLE, bool SLOPE>
__global__ void __launch_bounds__(D8_THREADS)
k_d8(const float *__restrict__ z, uint8_t *__restrict__ dir, float *__restrict__ slope, int ny, int nx,
     float nodata, d8_params prm) {
  __shared__ float sz[(D8_TR + 2) * (D8_TC + 2)];
  const int r0 = blockIdx.y * D8_TR, c0 = blockIdx.x * D8_TC;
  const int W = D8_TC + 2;
  const float qnan = __int_as_float(0x7fc00000);
  for (int i = threadIdx.x; i < (D8_TR + 2) * W; i += D8_THREADS) {
    int lr = i / W, lc = i - lr * W;
    int r = r0 + lr - 1, c = c0 + lc - 1;
    float v = qnan;
    if (r >= 0 && r < ny && c >= 0 && c < nx) {
      v = __ldg(&z[(size_t)r * nx + c]);
      if (v == nodata) v = qnan;
    }
    sz[i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D8_TR * D8_TC; i += D8_THREADS) {
    int lr = i / D8_TC, lc = i - lr * D8_TC;
    int r = r0 + lr, c = c0 + lc;
    if (r >= ny || c >= nx) continue;
    size_t g = (size_t)r * nx + c;
    float v = sz[(lr + 1) * W + lc + 1];
    if (v != v) {
      dir[g] = HC_DIR_NODATA;
      if (SLOPE) slope[g] = -1.0f;
      continue;
    }
    int best = -1;
    bool contaminated = false;
    double smax = 0.0;
    float smaxf = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      float zn = sz[(lr + 1 + c_dr[TABLE][k]) * W + lc + 1 + c_dc[TABLE][k]];
      if (zn != zn) { contaminated = true; continue; }
      if (TABLE == HC_TABLE_WBT) {
        if (zn < v) {
          double s = ((double)v - (double)zn) / prm.len[k];
          if (s > smax) { smax = s; best = k; }
        }
      } else {
        float s = __fmul_rn(prm.fact[k], __fsub_rn(v, zn));
        if (s > smaxf) { smaxf = s; best = k; }
      }
    }
    if (TABLE == HC_TABLE_TAUDEM && contaminated) {
      dir[g] = HC_DIR_NODATA;
      if (SLOPE) slope[g] = -1.0f;
      continue;
    }
    dir[g] = best < 0 ? (uint8_t)HC_DIR_NOFLOW : (uint8_t)c_code[TABLE][best];
    if (SLOPE) slope[g] = TABLE == HC_TABLE_WBT ? (float)smax : smaxf;
  }
}

int d8_run(hc_ctx *ctx, const float *z, uint8_t *dir, float *slope, int64_t ny, int64_t nx, float nodata,
           double dx, double dy, int table, cudaStream_t st) {
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (table != HC_TABLE_WBT && table != HC_TABLE_TAUDEM) { hc_set_error("d8: bad table %d", table); return HC_ERR_ARG; }
  if (!(dx > 0) || !(dy > 0)) { hc_set_error("d8: dx, dy must be > 0"); return HC_ERR_ARG; }
  static const int DR[2][8] = {{-1, 0, 1, 1, 1, 0, -1, -1}, {0, -1, -1, -1, 0, 1, 1, 1}};
  static const int DC[2][8] = {{1, 1, 1, 0, -1, -1, -1, 0}, {1, 1, 0, -1, -1, -1, 0, 1}};
  d8_params prm;
  for (int k = 0; k < 8; k++) {
    double l = sqrt((double)(DC[table][k] * DC[table][k]) * dx * dx + (double)(DR[table][k] * DR[table][k]) * dy * dy);
    prm.len[k] = l;
    prm.fact[k] = (float)(1.0 / l);
  }
  dim3 grid((unsigned)((nx + D8_TC - 1) / D8_TC), (unsigned)((ny + D8_TR - 1) / D8_TR));
  if (table == HC_TABLE_WBT) {
    if (slope) { HC_PROF(ctx, st, "k_d8"); k_d8<HC_TABLE_WBT, true><<<grid, D8_THREADS, 0, st>>>(z, dir, slope, (int)ny, (int)nx, nodata, prm); }
    else { HC_PROF(ctx, st, "k_d8"); k_d8<HC_TABLE_WBT, false><<<grid, D8_THREADS, 0, st>>>(z, dir, slope, (int)ny, (int)nx, nodata, prm); }
  } else {
    if (slope) { HC_PROF(ctx, st, "k_d8"); k_d8<HC_TABLE_TAUDEM, true><<<grid, D8_THREADS, 0, st>>>(z, dir, slope, (int)ny, (int)nx, nodata, prm); }
    else { HC_PROF(ctx, st, "k_d8"); k_d8<HC_TABLE_TAUDEM, false><<<grid, D8_THREADS, 0, st>>>(z, dir, slope, (int)ny, (int)nx, nodata, prm); }
  }
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}
