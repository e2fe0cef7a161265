// d8.cu — D8 steepest-descent direction + slope stencil.
//
// Replaces wbt.d8_pointer (run_aws.py:539) and the steepest-descent half of TauDEM D8FlowDir
// -p -sd8 (run_aws.py:594-603).  Arithmetic follows oracle/hydro_oracle.c orc_d8:
//   WBT table    slope_k = (double)(z - zk) / len_k      (IEEE double division)
//   TauDEM table slope_k = (float)(1/len_k) * (z - zk)   (two float32 roundings, no FMA)
// first strict maximum > 0 in the tool's own neighbour order; TauDEM marks border cells and cells
// with a nodata neighbour as contaminated (code 255).
#include "common.cuh"

#define D8_RS 32        // rows per thread (column strip)
#define D8_THREADS 256  // columns per block

struct d8_params {
  double len[8];
  float fact[8];
};

// Column-strip stencil: lane = column, each thread walks D8_RS rows keeping the 3x3 window in
// registers, so a cell costs 3 coalesced loads (two of them L1 hits), no shared memory, no barrier.
template <int TABLE, bool SLOPE>
__global__ void __launch_bounds__(D8_THREADS)
k_d8(const float *__restrict__ z, uint8_t *__restrict__ dir, uint8_t *__restrict__ dir2, float *__restrict__ slope,
     int ny, int nx, float nodata, d8_params prm, hc_rows rw) {
  const int c = blockIdx.x * D8_THREADS + threadIdx.x;
  if (c >= nx) return;
  const int rbeg = blockIdx.y * D8_RS, rend = min(rbeg + D8_RS, ny);
  const float qnan = __int_as_float(0x7fc00000);
  const bool hasl = c > 0, hasr = c < nx - 1;
  auto ld3 = [&](int r, float &l, float &m, float &rr) {
    l = m = rr = qnan;
    if (r < rw.rlo || r >= rw.rhi) return;
    const float *row = z + ((long long)r * nx + c);  // r = -1 / ny: a band's halo rows
    float vm = __ldg(row);
    float vl = hasl ? __ldg(row - 1) : qnan;
    float vr = hasr ? __ldg(row + 1) : qnan;
    m = vm == nodata ? qnan : vm;
    l = vl == nodata ? qnan : vl;
    rr = vr == nodata ? qnan : vr;
  };
  float w[3][3];
  ld3(rbeg - 1, w[0][0], w[0][1], w[0][2]);
  ld3(rbeg, w[1][0], w[1][1], w[1][2]);
#pragma unroll 4
  for (int r = rbeg; r < rend; r++) {
    ld3(r + 1, w[2][0], w[2][1], w[2][2]);
    const size_t g = (size_t)r * nx + c;
    const float v = w[1][1];
    uint8_t code;
    float sl;
    if (v != v) {
      code = HC_DIR_NODATA;
      sl = -1.0f;
    } else {
      int best = -1;
      bool contaminated = false;
      double smax = 0.0;
      float smaxf = 0.0f;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const float zn = w[1 + c_dr[TABLE][k]][1 + c_dc[TABLE][k]];
        contaminated |= zn != zn;   // (a NaN neighbour never wins below: every comparison with NaN is false)
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
        code = HC_DIR_NODATA;
        sl = -1.0f;
      } else {
        code = best < 0 ? (uint8_t)HC_DIR_NOFLOW : (uint8_t)c_code[TABLE][best];
        sl = TABLE == HC_TABLE_WBT ? (float)smax : smaxf;
      }
    }
    dir[g] = code;
    if (dir2) dir2[g] = code;  // second copy: the flat pass then only rewrites NOFLOW cells
    if (SLOPE) slope[g] = sl;
#pragma unroll
    for (int j = 0; j < 3; j++) { w[0][j] = w[1][j]; w[1][j] = w[2][j]; }
  }
}

int d8_run(hc_ctx *ctx, const float *z, uint8_t *dir, float *slope, int64_t ny, int64_t nx, float nodata,
           double dx, double dy, int table, cudaStream_t st) {
  return d8_run_rows(ctx, z, dir, slope, ny, nx, nodata, dx, dy, table, hc_make_rows(ny, 0), st, nullptr);
}

int d8_run_rows(hc_ctx *ctx, const float *z, uint8_t *dir, float *slope, int64_t ny, int64_t nx, float nodata,
                double dx, double dy, int table, hc_rows rw, cudaStream_t st, uint8_t *dir2) {
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
  dim3 grid((unsigned)((nx + D8_THREADS - 1) / D8_THREADS), (unsigned)((ny + D8_RS - 1) / D8_RS));
  if (table == HC_TABLE_WBT) {
    if (slope) { HC_PROF(ctx, st, "k_d8"); k_d8<HC_TABLE_WBT, true><<<grid, D8_THREADS, 0, st>>>(z, dir, dir2, slope, (int)ny, (int)nx, nodata, prm, rw); }
    else { HC_PROF(ctx, st, "k_d8"); k_d8<HC_TABLE_WBT, false><<<grid, D8_THREADS, 0, st>>>(z, dir, dir2, slope, (int)ny, (int)nx, nodata, prm, rw); }
  } else {
    if (slope) { HC_PROF(ctx, st, "k_d8"); k_d8<HC_TABLE_TAUDEM, true><<<grid, D8_THREADS, 0, st>>>(z, dir, dir2, slope, (int)ny, (int)nx, nodata, prm, rw); }
    else { HC_PROF(ctx, st, "k_d8"); k_d8<HC_TABLE_TAUDEM, false><<<grid, D8_THREADS, 0, st>>>(z, dir, dir2, slope, (int)ny, (int)nx, nodata, prm, rw); }
  }
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}
