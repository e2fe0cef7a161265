// stencil.cu — the dem_noise_removal / veg-bias smoothing stencils.
//
//   hc_conv2d_symm      pdt.apply_convolve_filter = scipy.signal.convolve2d(boundary='symm', mode='same')
//                       tools/derive.py:66-99 (DW kernels from bop.build_kernel, tools/build_operators.py:327-351;
//                       also the box dilation of dep.create_clump_array, tools/depressions.py:1640-1644)
//   hc_slope            pdt.slope_D2 / slope_D4 -> int8 degrees            tools/derive.py:102-130, 176-246
//   hc_curvature_d2     pdt.curvature_D2 (geometric or Laplacian)          tools/derive.py:133-173
//   hc_mean_std         np.std(curvature), the global sigma of CreateDTM    models/dem_noise_removal.py:413
//   hc_dtm_select       the filter-selection block of CreateDTM.run_tile    models/dem_noise_removal.py:414-471
//
// Arithmetic: np.gradient on float32 is (a - b) / float32(2h) inside, (a - b) / float32(h) one-sided
// at the array edges, every step rounded to float32 - reproduced with _rn intrinsics so gradients and
// curvature are bit-identical; atan/rad2deg go through CUDA's atanf, so int8 degrees can differ by one
// at a truncation boundary; the convolution accumulates in float32 (FMA), scipy's summation order is
// not reproducible, so it is a tolerance-class result (tests state the tolerance).
#include <stdlib.h>

#include "common.cuh"

#define CONV_MAXK 31
__constant__ float c_kern[CONV_MAXK * CONV_MAXK];

__device__ __forceinline__ int symm(int i, int n) {
  // half-sample symmetric reflection (d c b a | a b c d | d c b a)
  while (i < 0 || i >= n) i = i < 0 ? -i - 1 : 2 * n - i - 1;
  return i;
}

#define CV_TR 32
#define CV_TC 64
// generic fallback (any odd kh, kw <= 31): one output per thread-iteration, taps straight from shared memory
__global__ void __launch_bounds__(256)
k_conv2d_symm(const float *__restrict__ in, float *__restrict__ out, int ny, int nx, int kh, int kw) {
  extern __shared__ float sm[];
  const int ph = kh / 2, pw = kw / 2;
  const int W = CV_TC + kw - 1, H = CV_TR + kh - 1;
  const int r0 = blockIdx.y * CV_TR, c0 = blockIdx.x * CV_TC;
  for (int i = threadIdx.x; i < W * H; i += 256) {
    int lr = i / W, lc = i - lr * W;
    int r = symm(r0 + lr - ph, ny), c = symm(c0 + lc - pw, nx);
    sm[i] = __ldg(&in[(size_t)r * nx + c]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < CV_TR * CV_TC; i += 256) {
    int lr = i / CV_TC, lc = i - lr * CV_TC;
    int r = r0 + lr, c = c0 + lc;
    if (r >= ny || c >= nx) continue;
    float s = 0.f;
    // true convolution: out[i,j] = sum_ab ker[a,b] * in[i + ph - a, j + pw - b]
    for (int a = 0; a < kh; a++) {
      const float *row = &sm[(lr + kh - 1 - a) * W + lc + kw - 1];
      const float *kr = &c_kern[a * kw];
      for (int b = 0; b < kw; b++) s = fmaf(kr[b], row[-b], s);
    }
    out[(size_t)r * nx + c] = s;
  }
}

// Register-tiled form for the kernel heights the reference uses (3/5/7/9: the DW smoothers and the box
// dilation).  A thread owns CT_RPT consecutive rows of ONE column: per kernel column it loads the
// CT_RPT + KH - 1 values of that column's vertical window (lane-contiguous, conflict-free) and reuses
// each of them for up to KH taps from registers: KH * CT_RPT FMAs per CT_RPT + KH - 1 shared loads,
// which moves the bound from shared-memory bandwidth to the FMA pipe.
#define CT_TC 64
#define CT_RPT 16
#define CT_TR (4 * CT_RPT)
template <int KH>
__global__ void __launch_bounds__(256)
k_conv2d_symm_rt(const float *__restrict__ in, float *__restrict__ out, int ny, int nx, int kw) {
  extern __shared__ float sm[];
  const int ph = KH / 2, pw = kw / 2;
  const int W = CT_TC + kw - 1, H = CT_TR + KH - 1;
  const int r0 = blockIdx.y * CT_TR, c0 = blockIdx.x * CT_TC;
  for (int i = threadIdx.x; i < W * H; i += 256) {
    int lr = i / W, lc = i - lr * W;
    int r = symm(r0 + lr - ph, ny), c = symm(c0 + lc - pw, nx);
    sm[i] = __ldg(&in[(size_t)r * nx + c]);
  }
  __syncthreads();
  const int x = threadIdx.x & (CT_TC - 1), y0 = (threadIdx.x / CT_TC) * CT_RPT;
  float acc[CT_RPT];
#pragma unroll
  for (int j = 0; j < CT_RPT; j++) acc[j] = 0.f;
  for (int b = 0; b < kw; b++) {
    float kc[KH], w[CT_RPT + KH - 1];
#pragma unroll
    for (int a = 0; a < KH; a++) kc[a] = c_kern[a * kw + b];
    const float *col = &sm[y0 * W + x + kw - 1 - b];
#pragma unroll
    for (int m = 0; m < CT_RPT + KH - 1; m++) w[m] = col[m * W];
    // out[y0 + j] += ker[a][b] * in-window[j + KH - 1 - a]
#pragma unroll
    for (int a = 0; a < KH; a++)
#pragma unroll
      for (int j = 0; j < CT_RPT; j++) acc[j] = fmaf(kc[a], w[j + KH - 1 - a], acc[j]);
  }
  const int c = c0 + x;
  if (c >= nx) return;
#pragma unroll
  for (int j = 0; j < CT_RPT; j++) {
    int r = r0 + y0 + j;
    if (r < ny) out[(size_t)r * nx + c] = acc[j];
  }
}

// The four DW smoothers of CreateDTM (9x9, 7x7, 5x5, 3x3 on the SAME raster, models/dem_noise_removal.py:399-402)
// in one pass: the shared tile (halo 4) is loaded once and each kernel is applied from it with the register-tiled
// column walk above.  c_kern holds the four kernels back to back (81 + 49 + 25 + 9 floats).
template <int KH>
__device__ __forceinline__ void conv_apply_rt(const float *sm, int W, int off, const float *kern, int x, int y0,
                                              float *__restrict__ out, int r0, int c0, int ny, int nx) {
  // sm row 0 / col 0 = raster (r0 - 4, c0 - 4); this kernel's window starts `off` = 4 - KH/2 cells in
  float acc[CT_RPT];
#pragma unroll
  for (int j = 0; j < CT_RPT; j++) acc[j] = 0.f;
  for (int b = 0; b < KH; b++) {
    float kc[KH], w[CT_RPT + KH - 1];
#pragma unroll
    for (int a = 0; a < KH; a++) kc[a] = kern[a * KH + b];
    const float *col = &sm[(y0 + off) * W + off + x + KH - 1 - b];
#pragma unroll
    for (int m = 0; m < CT_RPT + KH - 1; m++) w[m] = col[m * W];
#pragma unroll
    for (int a = 0; a < KH; a++)
#pragma unroll
      for (int j = 0; j < CT_RPT; j++) acc[j] = fmaf(kc[a], w[j + KH - 1 - a], acc[j]);
  }
  const int c = c0 + x;
  if (c >= nx) return;
#pragma unroll
  for (int j = 0; j < CT_RPT; j++) {
    int r = r0 + y0 + j;
    if (r < ny) out[(size_t)r * nx + c] = acc[j];
  }
}

__global__ void __launch_bounds__(256)
k_conv2d_symm_dw4(const float *__restrict__ in, float *__restrict__ o9, float *__restrict__ o7, float *__restrict__ o5,
                  float *__restrict__ o3, int ny, int nx) {
  extern __shared__ float sm[];
  const int W = CT_TC + 8, H = CT_TR + 8;
  const int r0 = blockIdx.y * CT_TR, c0 = blockIdx.x * CT_TC;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int lr = warp; lr < H; lr += 8) {
    const float *row = in + (size_t)symm(r0 + lr - 4, ny) * nx;
    for (int lc = lane; lc < W; lc += 32) sm[lr * W + lc] = __ldg(row + symm(c0 + lc - 4, nx));
  }
  __syncthreads();
  const int x = threadIdx.x & (CT_TC - 1), y0 = (threadIdx.x / CT_TC) * CT_RPT;
  conv_apply_rt<9>(sm, W, 0, c_kern, x, y0, o9, r0, c0, ny, nx);
  conv_apply_rt<7>(sm, W, 1, c_kern + 81, x, y0, o7, r0, c0, ny, nx);
  conv_apply_rt<5>(sm, W, 2, c_kern + 81 + 49, x, y0, o5, r0, c0, ny, nx);
  conv_apply_rt<3>(sm, W, 3, c_kern + 81 + 49 + 25, x, y0, o3, r0, c0, ny, nx);
}

// ---- np.gradient pieces ------------------------------------------------------------------------
// Division by a per-call constant.  np.gradient divides every difference by float32(h) or float32(2h); the IEEE
// division sequence costs ~20 instructions and these stencils issue up to nine of them per cell.  For a fixed divisor
// h the quotient is q' = fma(r, y, q) with y = RN(1 / h), q = RN(x * y), r = fma(-q, h, x) (Markstein's correction).
// Whether that equals RN(x / h) for EVERY float32 x is not argued but checked: the first call with a new h compares the
// two forms over all 2^32 bit patterns on the device (k_cdiv_check, a few ms) and the short form is used only if not
// one differs; zeros, tiny / huge quotients, infinities and NaNs take the IEEE sequence in any case.
struct cdiv_t { float h, rcp; int ok; };   // (rcp is NaN while the short form is not validated: every quotient then fails the range test)
__device__ __forceinline__ float cdiv(float x, const cdiv_t &d) {
  const float q = __fmul_rn(x, d.rcp);
  const float r = __fmaf_rn(-q, d.h, x);
  const float q2 = __fmaf_rn(r, d.rcp, q);
  const float aq = fabsf(q2);
  if (!(aq >= 1e-30f && aq <= 1e30f)) return __fdiv_rn(x, d.h);   // zeros (sign!), tiny / huge quotients, inf, NaN
  return q2;
}
__global__ void k_cdiv_check(cdiv_t d, unsigned long long *bad) {
  const u32 i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 64u;   // 2^26 threads x 64 patterns
  u32 n = 0;
  for (u32 k = 0; k < 64u; k++) {
    const float x = __uint_as_float(i0 + k);
    const float a = cdiv(x, d), b = __fdiv_rn(x, d.h);
    if (__float_as_uint(a) != __float_as_uint(b) && !(a != a && b != b)) n++;
  }
  if (n) atomicAdd(bad, (unsigned long long)n);
}

struct grad_prm { cdiv_t h2x, h2y, hx, hy; double dxy; };  // float32(2dx), float32(2dy), float32(dx), float32(dy); float64 diagonal factor

__device__ __forceinline__ float gy(const float *z, int r, int c, int ny, int nx, const grad_prm &p) {
  if (ny == 1) return 0.f;
  if (r == 0) return cdiv(__fsub_rn(z[(size_t)nx + c], z[c]), p.hy);
  if (r == ny - 1) return cdiv(__fsub_rn(z[(size_t)r * nx + c], z[(size_t)(r - 1) * nx + c]), p.hy);
  return cdiv(__fsub_rn(z[(size_t)(r + 1) * nx + c], z[(size_t)(r - 1) * nx + c]), p.h2y);
}
__device__ __forceinline__ float gx(const float *z, int r, int c, int ny, int nx, const grad_prm &p) {
  if (nx == 1) return 0.f;
  const float *row = z + (size_t)r * nx;
  if (c == 0) return cdiv(__fsub_rn(row[1], row[0]), p.hx);
  if (c == nx - 1) return cdiv(__fsub_rn(row[c], row[c - 1]), p.hx);
  return cdiv(__fsub_rn(row[c + 1], row[c - 1]), p.h2x);
}

__device__ __forceinline__ int8_t to_deg_i8(float slope) {
  float deg = atanf(slope) * 57.29577951308232f;  // np.rad2deg(np.arctan(.)) in float32
  return (int8_t)(int)deg;                       // astype(int8): truncation toward zero
}
// The int8 truncation keeps 91 distinct answers, so the transcendental is not needed at all: thr[k - 1] = smallest
// float32 slope whose truncated degree is >= k (k = 1..90), found on the host by bisection over the float32 bit patterns
// WITH THE CALLER'S OWN arctan (numpy's, see stencils.degree_thresholds) -- the result is then bit-identical to
// np.rad2deg(np.arctan(x)).astype(int8) of that numpy, where the atanf form is off by one at a few truncation
// boundaries -- and a 7-step binary search replaces ~40 instructions of atanf.
struct deg_thr { float t[90]; int n; };
#define DEG_PAD 92   // shared copy of the thresholds, padded with NaN (never reached)
__device__ __forceinline__ void deg_stage(float *sthr, const deg_thr &d) {
  if (threadIdx.x < DEG_PAD) sthr[threadIdx.x] = threadIdx.x < 90 ? d.t[threadIdx.x] : __int_as_float(0x7fc00000);
}
// number of thresholds <= slope.  A cheap arctangent (x (pi/4 + 0.273 (1 - x)) on [0, 1], reflected above 1: within
// 0.25 degree) brackets the answer to e - 1 .. e + 1, two table comparisons settle it: ~20 instructions where the
// seven-step binary search over the table took ~80 (it was 40 % of k_slope's instruction stream).
__device__ __forceinline__ int8_t deg_lookup(float slope, const float *sthr) {
  const float a = fabsf(slope);
  const bool inv = a > 1.f;
  const float x = inv ? __fdividef(1.f, a) : a;
  float deg = x * (0.78539816f + 0.273f * (1.f - x)) * 57.29578f;
  if (inv) deg = 90.f - deg;
  int lo = (int)deg - 1;          // (NaN converts to 0)
  lo = lo < 0 ? 0 : (lo > 90 ? 90 : lo);
  return (int8_t)(lo + (slope >= sthr[lo] ? 1 : 0) + (slope >= sthr[lo + 1] ? 1 : 0));
}

// mode 0: slope_D2, mode 1: slope_D4 (max |central difference| over x, y and the two diagonals).
// Column strips: a thread walks SL_RS rows of one column with the 3 x 3 window in registers (three coalesced loads
// per cell instead of eight with reflected index arithmetic); raster-border cells take the general path.  The degree
// thresholds sit in shared memory (a per-lane index into kernel parameters would serialise in the constant cache).
#define SL_RS 32
__global__ void __launch_bounds__(256)
k_slope(const float *__restrict__ z, int8_t *__restrict__ out, float *__restrict__ mag, int ny, int nx, grad_prm p,
        int mode, const __grid_constant__ deg_thr thr) {
  __shared__ float sthr[DEG_PAD];
  deg_stage(sthr, thr);
  __syncthreads();
  const bool use_thr = thr.n == 90;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nx) return;
  const int rbeg = blockIdx.y * SL_RS, rend = min(rbeg + SL_RS, ny);
  const bool cin = c > 0 && c < nx - 1;
  auto to_deg = [&](float s_) -> int8_t { return use_thr ? deg_lookup(s_, sthr) : to_deg_i8(s_); };
  float w0[3], w1[3], w2[3];
  auto ld3 = [&](int r, float *w) {
    w[0] = w[1] = w[2] = 0.f;
    if (r < 0 || r >= ny || !cin) return;
    const float *row = z + (size_t)r * nx + c;
    w[0] = __ldg(row - 1); w[1] = __ldg(row); w[2] = __ldg(row + 1);
  };
  ld3(rbeg - 1, w0);
  ld3(rbeg, w1);
  for (int r = rbeg; r < rend; r++) {
    ld3(r + 1, w2);
    float s;
    if (cin && r > 0 && r < ny - 1) {
      const float dy = cdiv(__fsub_rn(w2[1], w0[1]), p.h2y), dx = cdiv(__fsub_rn(w1[2], w1[0]), p.h2x);
      if (mode == 0) s = __fsqrt_rn(__fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dx, dx)));
      else {
        const double dxy = p.dxy;
        const float dnw = (float)__dadd_rn(__dmul_rn(dxy, (double)w2[2]), __dmul_rn(-dxy, (double)w0[0]));
        const float dne = (float)__dadd_rn(__dmul_rn(dxy, (double)w2[0]), __dmul_rn(-dxy, (double)w0[2]));
        s = fmaxf(fmaxf(fabsf(dx), fabsf(dne)), fmaxf(fabsf(dy), fabsf(dnw)));
      }
    } else {
      const float dy = gy(z, r, c, ny, nx, p), dx = gx(z, r, c, ny, nx, p);
      if (mode == 0) s = __fsqrt_rn(__fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dx, dx)));
      else {
        // diagonal kernels are convolved in float64 and narrowed to float32 (tools/derive.py:226-241)
        int ru = symm(r - 1, ny), rd = symm(r + 1, ny), cl = symm(c - 1, nx), cr = symm(c + 1, nx);
        const double dxy = p.dxy;
        float dnw = (float)__dadd_rn(__dmul_rn(dxy, (double)z[(size_t)rd * nx + cr]), __dmul_rn(-dxy, (double)z[(size_t)ru * nx + cl]));
        float dne = (float)__dadd_rn(__dmul_rn(dxy, (double)z[(size_t)rd * nx + cl]), __dmul_rn(-dxy, (double)z[(size_t)ru * nx + cr]));
        s = fmaxf(fmaxf(fabsf(dx), fabsf(dne)), fmaxf(fabsf(dy), fabsf(dnw)));
      }
    }
    if (out) out[(size_t)r * nx + c] = to_deg(s);
    if (mag) mag[(size_t)r * nx + c] = s;
#pragma unroll
    for (int j = 0; j < 3; j++) { w0[j] = w1[j]; w1[j] = w2[j]; }
  }
}

__device__ __forceinline__ void norm_grad(const float *z, int r, int c, int ny, int nx, const grad_prm &p, float *ny_,
                                          float *nx_) {
  float dy = gy(z, r, c, ny, nx, p), dx = gx(z, r, c, ny, nx, p);
  float m = __fadd_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dx, dx))), 0.1f);
  *ny_ = __fdiv_rn(dy, m);
  *nx_ = __fdiv_rn(dx, m);
}

// Tiled: the first-stage field (normalised gradient, or the plain gradient for the Laplacian form) costs five
// IEEE divisions / square roots per point; it is evaluated ONCE per point of the tile + 1-cell ring into shared
// memory and the second differences read it from there (the untiled form re-evaluated it at four neighbours of
// every cell: 22 such operations per output instead of ~7.5).  Same operations in the same order per value, so
// the result is bit-identical.
#define CU_TR 16
#define CU_TC 64
__global__ void __launch_bounds__(256)
k_curv_d2(const float *__restrict__ z, float *__restrict__ out, int ny, int nx, grad_prm p, int geometric) {
  __shared__ float fy[(CU_TR + 2) * (CU_TC + 2)], fx[(CU_TR + 2) * (CU_TC + 2)];
  const int r0 = blockIdx.y * CU_TR, c0 = blockIdx.x * CU_TC;
  const int W = CU_TC + 2;
  for (int i = threadIdx.x; i < (CU_TR + 2) * W; i += 256) {
    int lr = i / W, lc = i - lr * W;
    int rr = r0 + lr - 1, cc = c0 + lc - 1;
    float a = 0.f, b = 0.f;
    if (rr >= 0 && rr < ny && cc >= 0 && cc < nx) {
      if (geometric) norm_grad(z, rr, cc, ny, nx, p, &a, &b);
      else { a = gy(z, rr, cc, ny, nx, p); b = gx(z, rr, cc, ny, nx, p); }
    }
    fy[i] = a;
    fx[i] = b;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < CU_TR * CU_TC; i += 256) {
    int lr = i / CU_TC, lc = i - lr * CU_TC;
    int r = r0 + lr, c = c0 + lc;
    if (r >= ny || c >= nx) continue;
    const int ci = (lr + 1) * W + lc + 1;
    float d2y, d2x;
    if (ny == 1) d2y = 0.f;
    else if (r == 0) d2y = cdiv(__fsub_rn(fy[ci + W], fy[ci]), p.hy);
    else if (r == ny - 1) d2y = cdiv(__fsub_rn(fy[ci], fy[ci - W]), p.hy);
    else d2y = cdiv(__fsub_rn(fy[ci + W], fy[ci - W]), p.h2y);
    if (nx == 1) d2x = 0.f;
    else if (c == 0) d2x = cdiv(__fsub_rn(fx[ci + 1], fx[ci]), p.hx);
    else if (c == nx - 1) d2x = cdiv(__fsub_rn(fx[ci], fx[ci - 1]), p.hx);
    else d2x = cdiv(__fsub_rn(fx[ci + 1], fx[ci - 1]), p.h2x);
    out[(size_t)r * nx + c] = __fadd_rn(d2y, d2x);
  }
}

// ---- global mean / population std (float64 accumulation) ----------------------------------------
__global__ void __launch_bounds__(256)
k_sum2(const float *__restrict__ x, size_t n, double *__restrict__ acc /* [sum, sumsq] */, double shift) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  double s = 0.0, q = 0.0;
  for (; i < n; i += stride) { double v = (double)x[i] - shift; s += v; q += v * v; }
  for (int o = 16; o; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); q += __shfl_down_sync(0xffffffffu, q, o); }
  __shared__ double ss[8], sq[8];
  int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { ss[w] = s; sq[w] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; k++) { s += ss[k]; q += sq[k]; }
    atomicAdd(&acc[0], s);
    atomicAdd(&acc[1], q);
  }
}

// ---- CreateDTM filter selection (models/dem_noise_removal.py:414-471) -------------------------------
struct dtm_prm { float cb0, cb1; int sb1, sb2, sb3, sb4; };  // sigma-scaled curvature breaks, slope breaks

__global__ void __launch_bounds__(256)
k_dtm_select(const float *__restrict__ dem, const float *__restrict__ z9, const float *__restrict__ z7,
             const float *__restrict__ z5, const float *__restrict__ z3, const int8_t *__restrict__ slope0,
             const float *__restrict__ curv, const uint8_t *__restrict__ swo, const uint8_t *__restrict__ edm,
             const uint8_t *__restrict__ tx, const uint8_t *__restrict__ artifact, float *__restrict__ zf,
             int8_t *__restrict__ filt, size_t n, dtm_prm p) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const float d = dem[i], cv = curv[i];
    const int s = slope0[i];
    const bool peak = cv < -p.cb1, peakfill = cv >= -p.cb1 && cv < -p.cb0;
    const bool valley = cv > p.cb1, valleyfill = cv <= p.cb1 && cv > p.cb0;
    const bool steep = s >= p.sb4;
    float z = z9[i];
    int f = 9;
    if (s >= p.sb1 && s <= p.sb2) { f = 7; z = z7[i]; }
    if (s >= p.sb2 && s <= p.sb3) { f = 5; z = z5[i]; }
    if ((s >= p.sb3 && s <= p.sb4) || valleyfill || peakfill) { f = 3; z = z3[i]; }
    if (steep) f = 1;
    if (peak) f = 2;
    if (valley) f = 0;
    if (steep || peak || valley) z = d;
    if (swo && swo[i]) { f = 3; z = z3[i]; }
    if (edm && edm[i] == 5 && f < 3) { f = 3; z = z3[i]; }
    if (tx && tx[i]) { f = 1; z = d; }
    if (artifact && artifact[i] == 1) z = d;
    zf[i] = z;
    filt[i] = (int8_t)f;
  }
}

// ---- C ABI --------------------------------------------------------------------------------------------
#define ST_CTX(ctx) HC_ENTER(ctx)

// validated constant divisors of this process (see cdiv): h -> short form allowed
static pthread_mutex_t g_cdiv_mu = PTHREAD_MUTEX_INITIALIZER;
static struct { float h; int ok; } g_cdiv[32];
static int g_cdiv_n = 0;

static cdiv_t make_cdiv(float h) {
  cdiv_t d;
  d.h = h;
  d.rcp = nanf("");
  d.ok = 0;
  if (!(h > 0.f) || getenv("HC_NO_CDIV")) return d;
  pthread_mutex_lock(&g_cdiv_mu);
  int found = -1;
  for (int i = 0; i < g_cdiv_n; i++) if (g_cdiv[i].h == h) found = i;
  if (found < 0 && g_cdiv_n < 32) {
    unsigned long long *bad = nullptr, hb = 1;
    cdiv_t t = d;
    t.ok = 1;
    t.rcp = 1.0f / h;
    if (cudaMalloc(&bad, 8) == cudaSuccess) {
      cudaMemset(bad, 0, 8);
      k_cdiv_check<<<(1u << 26) / 256, 256>>>(t, bad);
      if (cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost) != cudaSuccess) hb = 1;
      cudaFree(bad);
    }
    cudaGetLastError();
    if (getenv("HC_TRACE")) fprintf(stderr, "[cdiv] h = %.9g: %llu of 2^32 patterns differ from the IEEE division\n", (double)h, hb);
    found = g_cdiv_n++;
    g_cdiv[found].h = h;
    g_cdiv[found].ok = hb == 0;
  }
  if (found >= 0) d.ok = g_cdiv[found].ok;
  if (d.ok) d.rcp = 1.0f / h;
  pthread_mutex_unlock(&g_cdiv_mu);
  return d;
}

static grad_prm make_grad(double dx, double dy) {
  grad_prm p;
  p.h2x = make_cdiv((float)(2.0 * dx)); p.h2y = make_cdiv((float)(2.0 * dy)); p.hx = make_cdiv((float)dx); p.hy = make_cdiv((float)dy);
  p.dxy = 1.0 / (2.0 * sqrt(dx * dx + dy * dy));  // float64 Python scalar in the reference
  return p;
}

extern "C" int hc_conv2d_symm_dev(hc_ctx *ctx, const float *in, const float *kernel_host, int kh, int kw, float *out,
                                  int64_t ny, int64_t nx, void *stream) {
  ST_CTX(ctx);
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (!in || !out || !kernel_host || in == out) { hc_set_error("hc_conv2d_symm: bad pointer"); return HC_ERR_ARG; }
  if (kh < 1 || kw < 1 || kh > CONV_MAXK || kw > CONV_MAXK || !(kh & 1) || !(kw & 1)) {
    hc_set_error("hc_conv2d_symm: kernel must be odd x odd, at most %d", CONV_MAXK);
    return HC_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  HC_CUDA(cudaMemcpyToSymbolAsync(c_kern, kernel_host, sizeof(float) * kh * kw, 0, cudaMemcpyHostToDevice, st));
  if (kh == 3 || kh == 5 || kh == 7 || kh == 9) {
    dim3 grid((unsigned)((nx + CT_TC - 1) / CT_TC), (unsigned)((ny + CT_TR - 1) / CT_TR));
    size_t smem = sizeof(float) * (CT_TC + kw - 1) * (CT_TR + kh - 1);
    HC_PROF(ctx, st, "k_conv2d_symm");
    if (kh == 3) k_conv2d_symm_rt<3><<<grid, 256, smem, st>>>(in, out, (int)ny, (int)nx, kw);
    else if (kh == 5) k_conv2d_symm_rt<5><<<grid, 256, smem, st>>>(in, out, (int)ny, (int)nx, kw);
    else if (kh == 7) k_conv2d_symm_rt<7><<<grid, 256, smem, st>>>(in, out, (int)ny, (int)nx, kw);
    else k_conv2d_symm_rt<9><<<grid, 256, smem, st>>>(in, out, (int)ny, (int)nx, kw);
    HC_LAUNCH_CHECK(ctx);
    return HC_OK;
  }
  dim3 grid((unsigned)((nx + CV_TC - 1) / CV_TC), (unsigned)((ny + CV_TR - 1) / CV_TR));
  size_t smem = sizeof(float) * (CV_TC + kw - 1) * (CV_TR + kh - 1);
  HC_PROF(ctx, st, "k_conv2d_symm");
  k_conv2d_symm<<<grid, 256, smem, st>>>(in, out, (int)ny, (int)nx, kh, kw);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

// kernels_host: 81 + 49 + 25 + 9 float32 (the 9x9, 7x7, 5x5 and 3x3 kernels back to back, HOST pointer)
extern "C" int hc_conv2d_symm_dw4_dev(hc_ctx *ctx, const float *in, const float *kernels_host, float *out9, float *out7,
                                      float *out5, float *out3, int64_t ny, int64_t nx, void *stream) {
  ST_CTX(ctx);
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (!in || !kernels_host || !out9 || !out7 || !out5 || !out3) { hc_set_error("hc_conv2d_symm_dw4: null pointer"); return HC_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  HC_CUDA(cudaMemcpyToSymbolAsync(c_kern, kernels_host, sizeof(float) * (81 + 49 + 25 + 9), 0, cudaMemcpyHostToDevice, st));
  dim3 grid((unsigned)((nx + CT_TC - 1) / CT_TC), (unsigned)((ny + CT_TR - 1) / CT_TR));
  size_t smem = sizeof(float) * (CT_TC + 8) * (CT_TR + 8);
  HC_PROF(ctx, st, "k_conv2d_symm_dw4");
  k_conv2d_symm_dw4<<<grid, 256, smem, st>>>(in, out9, out7, out5, out3, (int)ny, (int)nx);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

extern "C" int hc_slope_dev(hc_ctx *ctx, const float *z, int8_t *deg, float *mag, int64_t ny, int64_t nx, double dx,
                            double dy, int d4, void *stream) {
  ST_CTX(ctx);
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (!z || (!deg && !mag)) { hc_set_error("hc_slope: bad argument"); return HC_ERR_ARG; }
  grad_prm p = make_grad(dx, dy);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((nx + 255) / 256), (unsigned)((ny + SL_RS - 1) / SL_RS));
  HC_PROF(ctx, st, "k_slope");
  deg_thr thr;
  thr.n = ctx->deg_thr_n;
  memcpy(thr.t, ctx->deg_thr, sizeof(thr.t));
  k_slope<<<grid, 256, 0, st>>>(z, deg, mag, (int)ny, (int)nx, p, d4 ? 1 : 0, thr);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

extern "C" int hc_slope_thresholds_set(hc_ctx *ctx, const float *thr, int32_t n) {
  ST_CTX(ctx);
  if (n == 0) { ctx->deg_thr_n = 0; return HC_OK; }   // back to the atanf form
  if (!thr || n != 90) { hc_set_error("hc_slope_thresholds_set: need the 90 degree thresholds (or n = 0)"); return HC_ERR_ARG; }
  for (int k = 0; k < 90; k++) {
    if (!(thr[k] == thr[k]) || (k && thr[k] < thr[k - 1])) { hc_set_error("hc_slope_thresholds_set: thresholds must ascend"); return HC_ERR_ARG; }
    ctx->deg_thr[k] = thr[k];
  }
  ctx->deg_thr_n = 90;
  return HC_OK;
}

extern "C" int hc_curvature_d2_dev(hc_ctx *ctx, const float *z, float *out, int64_t ny, int64_t nx, double dx, double dy,
                                   int geometric, void *stream) {
  ST_CTX(ctx);
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (!z || !out || ny > 65535 * CU_TR) { hc_set_error("hc_curvature_d2: bad argument"); return HC_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((nx + CU_TC - 1) / CU_TC), (unsigned)((ny + CU_TR - 1) / CU_TR));
  HC_PROF(ctx, st, "k_curv_d2");
  k_curv_d2<<<grid, 256, 0, st>>>(z, out, (int)ny, (int)nx, make_grad(dx, dy), geometric ? 1 : 0);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

extern "C" int hc_mean_std_dev(hc_ctx *ctx, const float *x, int64_t n, double *mean, double *std, void *stream) {
  ST_CTX(ctx);
  if (mean) *mean = 0;
  if (std) *std = 0;
  if (n <= 0) return HC_OK;
  if (!x) { hc_set_error("hc_mean_std: null pointer"); return HC_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  HC_TRY(arena_reset(ctx));
  double *acc = (double *)arena_take(ctx, 256);
  if (!acc) return HC_ERR_NOMEM;
  double host[2];
  // two passes: mean first, then centred second moment (stable)
  HC_CUDA(cudaMemsetAsync(acc, 0, 16, st));
  HC_PROF(ctx, st, "k_sum2");
  k_sum2<<<ctx->sm_count * 8, 256, 0, st>>>(x, (size_t)n, acc, 0.0);
  HC_LAUNCH_CHECK(ctx);
  HC_CUDA(cudaMemcpyAsync(host, acc, 16, cudaMemcpyDeviceToHost, st));
  HC_CUDA(cudaStreamSynchronize(st));
  double m = host[0] / (double)n;
  HC_CUDA(cudaMemsetAsync(acc, 0, 16, st));
  HC_PROF(ctx, st, "k_sum2");
  k_sum2<<<ctx->sm_count * 8, 256, 0, st>>>(x, (size_t)n, acc, m);
  HC_LAUNCH_CHECK(ctx);
  HC_CUDA(cudaMemcpyAsync(host, acc, 16, cudaMemcpyDeviceToHost, st));
  HC_CUDA(cudaStreamSynchronize(st));
  if (mean) *mean = m + host[0] / (double)n;
  if (std) {
    double d = host[0] / (double)n;
    *std = sqrt(fmax(host[1] / (double)n - d * d, 0.0));
  }
  return HC_OK;
}

// sums[0] = sum(x - shift), sums[1] = sum((x - shift)^2) in float64 (host pointer): the partial sums a row band
// contributes to the global sigma of CreateDTM (models/dem_noise_removal.py:413) before an all_reduce
extern "C" int hc_sum_shifted_dev(hc_ctx *ctx, const float *x, int64_t n, double shift, double *sums, void *stream) {
  ST_CTX(ctx);
  if (!sums) { hc_set_error("hc_sum_shifted: null pointer"); return HC_ERR_ARG; }
  sums[0] = sums[1] = 0.0;
  if (n <= 0) return HC_OK;
  if (!x) { hc_set_error("hc_sum_shifted: null pointer"); return HC_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  HC_TRY(arena_reset(ctx));
  double *acc = (double *)arena_take(ctx, 256);
  if (!acc) return HC_ERR_NOMEM;
  HC_CUDA(cudaMemsetAsync(acc, 0, 16, st));
  HC_PROF(ctx, st, "k_sum2");
  k_sum2<<<ctx->sm_count * 8, 256, 0, st>>>(x, (size_t)n, acc, shift);
  HC_LAUNCH_CHECK(ctx);
  HC_CUDA(cudaMemcpyAsync(sums, acc, 16, cudaMemcpyDeviceToHost, st));
  HC_CUDA(cudaStreamSynchronize(st));
  return HC_OK;
}

// ---- vegetation-bias lookup (models/dem_noise_removal.py:764-805) ------------------------------------------------
// 3x3 box mean with symmetric borders, the way apply_convolve_filter(x, np.ones((3,3)), normalize_kernel=True) makes
// it for a non-float32 raster: nine float64 products with w = 1/9 summed in float64, narrowed to float32, values
// above `zero_above` set to 0 (:771, tree cover only), then astype(int).  For integer-valued inputs (the Byte
// tree-cover / tree-height rasters) the float32 narrowing makes the result independent of the summation order, so
// the index is exact; `smooth` (optional) receives the float32 array the reference keeps as self.TC_array.
__global__ void __launch_bounds__(256)
k_box3_index(const float *__restrict__ in, const uint8_t *__restrict__ in_zero, float in_max, int ny, int nx,
             double w, float zero_above, int32_t *__restrict__ idx, float *__restrict__ smooth) {
  int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c >= nx) return;
  // (interior cells skip the reflection arithmetic; same taps, same order)
  const bool inner = r > 0 && r + 1 < ny && c > 0 && c + 1 < nx;
  int cl = inner ? c - 1 : symm(c - 1, nx), cr = inner ? c + 1 : symm(c + 1, nx);
  double acc = 0.0;
#pragma unroll
  for (int dr = -1; dr <= 1; dr++) {
    size_t o = (size_t)(inner ? r + dr : symm(r + dr, ny)) * nx;
    int cc[3] = {cl, c, cr};
#pragma unroll
    for (int k = 0; k < 3; k++) {
      float x = in[o + cc[k]];
      if (x > in_max) x = in_max;                  // Theight_array[Theight_array > max] = max (:777)
      if (in_zero && in_zero[o + cc[k]]) x = 0.f;  // TC_array[ocean_mask == 1] = 0 (:687)
      acc = __dadd_rn(acc, __dmul_rn((double)x, w));
    }
  }
  float f = (float)acc;
  if (f > zero_above) f = 0.f;
  size_t i = (size_t)r * nx + c;
  if (smooth) smooth[i] = f;
  idx[i] = (int32_t)f;
}

// bias = cube[i0, i1, min(i2, cap2)] per cell (:802-805; the 2-D table of :620-629 is the n1 == 1 case with i1 null),
// the capped third index written back (:792), bias zeroed under `zero_mask` (:361) and, when `dem` is given,
// dem_out = dem - bias in float64 (:364) -- one pass over the rasters; the table (a few hundred KB) stays in L1/L2.
__global__ void __launch_bounds__(256)
k_lut3d_gather(const int32_t *__restrict__ i0, const int32_t *__restrict__ i1, int8_t *__restrict__ i2, int cap2,
               const double *__restrict__ cube, int n0, int n1, int n2, const uint8_t *__restrict__ zero_mask,
               double *__restrict__ bias, const float *__restrict__ dem, double *__restrict__ dem_out, size_t n,
               int *__restrict__ oob) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int a = i0[i], b = i1 ? i1[i] : 0, s = i2[i];
    if (s > cap2) { s = cap2; i2[i] = (int8_t)s; }
    double v = 0.0;
    if ((unsigned)a < (unsigned)n0 && (unsigned)b < (unsigned)n1 && (unsigned)s < (unsigned)n2)
      v = __ldg(cube + ((size_t)a * n1 + b) * n2 + s);
    else
      *oob = 1;
    if (zero_mask && zero_mask[i]) v = 0.0;
    if (bias) bias[i] = v;
    if (dem_out) dem_out[i] = (double)dem[i] - v;
  }
}

extern "C" int hc_box3_index_dev(hc_ctx *ctx, const float *in, const uint8_t *in_zero, float in_max, int64_t ny,
                                 int64_t nx, float zero_above, int32_t *idx, float *smooth, void *stream) {
  ST_CTX(ctx);
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (!in || !idx || ny > 65535) { hc_set_error("hc_box3_index: bad argument"); return HC_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((nx + 255) / 256), (unsigned)ny);
  HC_PROF(ctx, st, "k_box3_index");
  k_box3_index<<<grid, 256, 0, st>>>(in, in_zero, in_max, (int)ny, (int)nx, 1.0 / 9.0, zero_above, idx, smooth);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

extern "C" int hc_lut3d_gather_dev(hc_ctx *ctx, const int32_t *i0, const int32_t *i1, int8_t *i2, int cap2,
                                   const double *cube, int n0, int n1, int n2, const uint8_t *zero_mask, double *bias,
                                   const float *dem, double *dem_out, int64_t n, int *oob_host, void *stream) {
  ST_CTX(ctx);
  if (oob_host) *oob_host = 0;
  if (n <= 0) return HC_OK;
  if (!i0 || !i2 || !cube || n0 <= 0 || n1 <= 0 || n2 <= 0 || (!bias && !dem_out) || (dem_out && !dem) ||
      (n1 > 1 && !i1)) {
    hc_set_error("hc_lut3d_gather: bad argument");
    return HC_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  HC_TRY(arena_reset(ctx));
  int *oob = (int *)arena_take(ctx, 256);
  if (!oob) return HC_ERR_NOMEM;
  HC_CUDA(cudaMemsetAsync(oob, 0, sizeof(int), st));
  HC_PROF(ctx, st, "k_lut3d_gather");
  k_lut3d_gather<<<ctx->sm_count * 8, 256, 0, st>>>(i0, i1, i2, cap2, cube, n0, n1, n2, zero_mask, bias, dem, dem_out,
                                                    (size_t)n, oob);
  HC_LAUNCH_CHECK(ctx);
  int h = 0;
  HC_CUDA(cudaMemcpyAsync(&h, oob, sizeof(int), cudaMemcpyDeviceToHost, st));
  HC_CUDA(cudaStreamSynchronize(st));
  if (oob_host) *oob_host = h;
  if (h) { hc_set_error("hc_lut3d_gather: index outside the lookup table"); return HC_ERR_ARG; }
  return HC_OK;
}

extern "C" int hc_dtm_select_dev(hc_ctx *ctx, const float *dem, const float *z9, const float *z7, const float *z5,
                                 const float *z3, const int8_t *slope0, const float *curv, const uint8_t *swo,
                                 const uint8_t *edm, const uint8_t *tx, const uint8_t *artifact, float *zf, int8_t *filt,
                                 int64_t n, float cbreak0, float cbreak1, const int *slopebreaks5, void *stream) {
  ST_CTX(ctx);
  if (n <= 0) return HC_OK;
  if (!dem || !z9 || !z7 || !z5 || !z3 || !slope0 || !curv || !zf || !filt || !slopebreaks5) {
    hc_set_error("hc_dtm_select: null pointer");
    return HC_ERR_ARG;
  }
  dtm_prm p;
  // cbreak = [c * np.std(curv) for c in curvebreaks] are np.float32 scalars in the reference (NEP 50):
  // the caller passes them already narrowed
  p.cb0 = cbreak0;
  p.cb1 = cbreak1;
  p.sb1 = slopebreaks5[1]; p.sb2 = slopebreaks5[2]; p.sb3 = slopebreaks5[3]; p.sb4 = slopebreaks5[4];
  cudaStream_t st = (cudaStream_t)stream;
  HC_PROF(ctx, st, "k_dtm_select");
  k_dtm_select<<<ctx->sm_count * 8, 256, 0, st>>>(dem, z9, z7, z5, z3, slope0, curv, swo, edm, tx, artifact, zf, filt,
                                                  (size_t)n, p);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

// =====================================================================================================================
// Fused smoothing core of CreateDTM.run_tile (models/dem_noise_removal.py:397-471).  The unfused chain writes the four
// smoothed rasters, the slope and curvature of the 5x5 one, reads the curvature twice more for sigma and then reads all
// seven rasters back for the selection: ~33 GB on a 20000^2 tile for 12 B/cell of results.  Sigma is a global
// quantity, so two passes over the DEM are unavoidable -- but the smoothed rasters never have to exist:
//   k_dtm_pre     tile + 4-cell halo of the DEM in shared memory -> 5x5 smoothing on tile + 2 -> normalised gradient
//                 field on tile + 1 -> curvature (float32) and slope (int8 degrees) of the tile: 5 B/cell out, plus
//                 the float64 sum / sum of squares of the curvature for sigma
//   k_dtm_apply   reads the DEM tile, the two small rasters and the masks, decides the filter per cell, runs the
//                 9x9 / 7x7 / 5x5 / 3x3 smoothings from the staged tile and writes only Zf (4 B) and the filter code
//                 (1 B) -- 18 B/cell of traffic in all instead of ~80.
// Every value is produced by the same operations in the same order as in k_conv2d_symm_dw4 / k_slope / k_curv_d2 /
// k_dtm_select, so the fused result is bit-identical to the unfused chain's.
// =====================================================================================================================
#define DF_T 64                 // output tile edge
#define DF_DW (DF_T + 8)        // DEM window (halo 4)
#define DF_ZW (DF_T + 4)        // 5x5-smoothed window (halo 2)
#define DF_GW (DF_T + 2)        // gradient-field window (halo 1)
#define DF_RPT 16
#define DF_PRE_SMEM ((DF_DW * DF_DW + DF_ZW * DF_ZW + 2 * DF_GW * DF_GW) * 4)
#define DF_APP_SMEM (DF_DW * DF_DW * 4)

struct dtm_pre_prm { grad_prm g; int d4; int row_lo, row_hi; };

// np.gradient pieces on the 5x5-smoothed window (lr, lc = window position of raster cell (r, c))
__device__ __forceinline__ float gy_w(const float *Z, int lr, int lc, int r, int ny, const grad_prm &p) {
  if (ny == 1) return 0.f;
  const float *q = Z + lr * DF_ZW + lc;
  if (r == 0) return cdiv(__fsub_rn(q[DF_ZW], q[0]), p.hy);
  if (r == ny - 1) return cdiv(__fsub_rn(q[0], q[-DF_ZW]), p.hy);
  return cdiv(__fsub_rn(q[DF_ZW], q[-DF_ZW]), p.h2y);
}
__device__ __forceinline__ float gx_w(const float *Z, int lr, int lc, int c, int nx, const grad_prm &p) {
  if (nx == 1) return 0.f;
  const float *q = Z + lr * DF_ZW + lc;
  if (c == 0) return cdiv(__fsub_rn(q[1], q[0]), p.hx);
  if (c == nx - 1) return cdiv(__fsub_rn(q[0], q[-1]), p.hx);
  return cdiv(__fsub_rn(q[1], q[-1]), p.h2x);
}

template <int KH>
__device__ __forceinline__ void conv_cols_rt(const float *D, int off, const float *kern, int x, int y0, float *acc) {
  // same accumulation order as conv_apply_rt: kernel columns outermost, kernel rows next
#pragma unroll
  for (int j = 0; j < DF_RPT; j++) acc[j] = 0.f;
  for (int b = 0; b < KH; b++) {
    float kc[KH], w[DF_RPT + KH - 1];
#pragma unroll
    for (int a = 0; a < KH; a++) kc[a] = kern[a * KH + b];
    const float *col = &D[(y0 + off) * DF_DW + off + x + KH - 1 - b];
#pragma unroll
    for (int m = 0; m < DF_RPT + KH - 1; m++) w[m] = col[m * DF_DW];
#pragma unroll
    for (int a = 0; a < KH; a++)
#pragma unroll
      for (int j = 0; j < DF_RPT; j++) acc[j] = fmaf(kc[a], w[j + KH - 1 - a], acc[j]);
  }
}

// DEM window with a 4-cell halo, half-sample symmetric at the raster edges (interior tiles skip the reflection)
__device__ __forceinline__ void dtm_load_window(const float *__restrict__ dem, float *D, int r0, int c0, int ny, int nx) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (r0 >= 4 && c0 >= 4 && r0 + DF_T + 4 <= ny && c0 + DF_T + 4 <= nx) {
    for (int lr = warp; lr < DF_DW; lr += 8) {
      const float *row = dem + (size_t)(r0 + lr - 4) * nx + (c0 - 4);
      for (int lc = lane; lc < DF_DW; lc += 32) D[lr * DF_DW + lc] = __ldg(row + lc);
    }
  } else {
    for (int lr = warp; lr < DF_DW; lr += 8) {
      const float *row = dem + (size_t)symm(r0 + lr - 4, ny) * nx;
      for (int lc = lane; lc < DF_DW; lc += 32) D[lr * DF_DW + lc] = __ldg(row + symm(c0 + lc - 4, nx));
    }
  }
}

__global__ void __launch_bounds__(256)
k_dtm_pre(const float *__restrict__ dem, float *__restrict__ curv, int8_t *__restrict__ slope0, double *__restrict__ sums,
          int ny, int nx, dtm_pre_prm P, const __grid_constant__ deg_thr thr) {
  extern __shared__ float sm[];
  float *D = sm, *Z = sm + DF_DW * DF_DW, *FY = Z + DF_ZW * DF_ZW, *FX = FY + DF_GW * DF_GW;
  __shared__ float sthr[DEG_PAD];
  __shared__ double s_red[16];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = blockIdx.y * DF_T, c0 = blockIdx.x * DF_T;
  deg_stage(sthr, thr);
  dtm_load_window(dem, D, r0, c0, ny, nx);
  __syncthreads();
  // 5x5 smoothing on tile + 2 (68 x 68); Z(i, j) = raster (r0 - 2 + i, c0 - 2 + j): column walks of 17 rows for the first
  // 64 columns, the last four columns cell by cell (same accumulation order: kernel columns outermost)
  {
    const float *kern = c_kern + 81 + 49;
    const int x = tid & 63, i0 = (tid >> 6) * 17;
    float acc[17];
#pragma unroll
    for (int j = 0; j < 17; j++) acc[j] = 0.f;
    for (int b = 0; b < 5; b++) {
      float kc[5], w[21];
#pragma unroll
      for (int a = 0; a < 5; a++) kc[a] = kern[a * 5 + b];
      const float *col = &D[i0 * DF_DW + x + 4 - b];
#pragma unroll
      for (int m = 0; m < 21; m++) w[m] = col[m * DF_DW];
#pragma unroll
      for (int a = 0; a < 5; a++)
#pragma unroll
        for (int j = 0; j < 17; j++) acc[j] = fmaf(kc[a], w[j + 4 - a], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 17; j++) Z[(i0 + j) * DF_ZW + x] = acc[j];
    for (int t = tid; t < DF_ZW * 4; t += 256) {
      const int i = t >> 2, xx = 64 + (t & 3);
      float s_ = 0.f;
      for (int b = 0; b < 5; b++)
#pragma unroll
        for (int a = 0; a < 5; a++) s_ = fmaf(kern[a * 5 + b], D[(i + 4 - a) * DF_DW + xx + 4 - b], s_);
      Z[i * DF_ZW + xx] = s_;
    }
  }
  __syncthreads();
  // normalised gradient field on tile + 1 (zero outside the raster, as k_curv_d2 has it)
  for (int i = tid; i < DF_GW * DF_GW; i += 256) {
    const int u = i / DF_GW, v = i - u * DF_GW;
    const int rr = r0 + u - 1, cc = c0 + v - 1;
    float a = 0.f, b = 0.f;
    if (rr >= 0 && rr < ny && cc >= 0 && cc < nx) {
      const float dy = gy_w(Z, u + 1, v + 1, rr, ny, P.g), dx = gx_w(Z, u + 1, v + 1, cc, nx, P.g);
      const float m = __fadd_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dx, dx))), 0.1f);
      a = __fdiv_rn(dy, m);
      b = __fdiv_rn(dx, m);
    }
    FY[i] = a;
    FX[i] = b;
  }
  __syncthreads();
  // per cell: curvature and the slope of the 5x5 result; rows of the tile are walked warp-wide (coalesced stores)
  double s1 = 0.0, s2 = 0.0;
  for (int i = tid; i < DF_T * DF_T; i += 256) {
    const int y = i >> 6, x = i & 63;
    const int r = r0 + y, c = c0 + x;
    if (r >= ny || c >= nx) continue;
    const int gi = (y + 1) * DF_GW + x + 1;
    float d2y, d2x;
    if (ny == 1) d2y = 0.f;
    else if (r == 0) d2y = cdiv(__fsub_rn(FY[gi + DF_GW], FY[gi]), P.g.hy);
    else if (r == ny - 1) d2y = cdiv(__fsub_rn(FY[gi], FY[gi - DF_GW]), P.g.hy);
    else d2y = cdiv(__fsub_rn(FY[gi + DF_GW], FY[gi - DF_GW]), P.g.h2y);
    if (nx == 1) d2x = 0.f;
    else if (c == 0) d2x = cdiv(__fsub_rn(FX[gi + 1], FX[gi]), P.g.hx);
    else if (c == nx - 1) d2x = cdiv(__fsub_rn(FX[gi], FX[gi - 1]), P.g.hx);
    else d2x = cdiv(__fsub_rn(FX[gi + 1], FX[gi - 1]), P.g.h2x);
    const float cv = __fadd_rn(d2y, d2x);
    if (r >= P.row_lo && r < P.row_hi) { s1 += (double)cv; s2 += (double)cv * (double)cv; }
    const size_t g = (size_t)r * nx + c;
    if (curv) curv[g] = cv;
    if (!slope0) continue;
    // slope of the 5x5 result (k_slope: D2 magnitude or D4 maximum), int8 degrees
    const int lr = y + 2, lc = x + 2;
    const float dy = gy_w(Z, lr, lc, r, ny, P.g), dx = gx_w(Z, lr, lc, c, nx, P.g);
    float sl;
    if (!P.d4) sl = __fsqrt_rn(__fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dx, dx)));
    else {
      // diagonal kernels are convolved in float64 and narrowed to float32 (tools/derive.py:226-241), symmetric at the edges
      const int du = r > 0 ? -1 : 0, dd = r < ny - 1 ? 1 : 0, dl = c > 0 ? -1 : 0, dr_ = c < nx - 1 ? 1 : 0;   // symm(r +- 1) - r
      const float *q = Z + lr * DF_ZW + lc;
      const double dxy = P.g.dxy;
      const float dnw = (float)__dadd_rn(__dmul_rn(dxy, (double)q[dd * DF_ZW + dr_]), __dmul_rn(-dxy, (double)q[du * DF_ZW + dl]));
      const float dne = (float)__dadd_rn(__dmul_rn(dxy, (double)q[dd * DF_ZW + dl]), __dmul_rn(-dxy, (double)q[du * DF_ZW + dr_]));
      sl = fmaxf(fmaxf(fabsf(dx), fabsf(dne)), fmaxf(fabsf(dy), fabsf(dnw)));
    }
    const int8_t s = thr.n == 90 ? deg_lookup(sl, sthr) : to_deg_i8(sl);
    slope0[g] = s;
  }
  for (int o = 16; o; o >>= 1) { s1 += __shfl_down_sync(0xffffffffu, s1, o); s2 += __shfl_down_sync(0xffffffffu, s2, o); }
  if (lane == 0) { s_red[warp] = s1; s_red[8 + warp] = s2; }
  __syncthreads();
  if (tid == 0) {
    for (int k = 1; k < 8; k++) { s1 += s_red[k]; s2 += s_red[8 + k]; }
    atomicAdd(&sums[0], s1);
    atomicAdd(&sums[1], s2);
  }
}

__global__ void __launch_bounds__(256)
k_dtm_apply(const float *__restrict__ dem, const float *__restrict__ curv, const int8_t *__restrict__ slope0,
            const uint8_t *__restrict__ swo, const uint8_t *__restrict__ edm, const uint8_t *__restrict__ tx,
            const uint8_t *__restrict__ artifact, float *__restrict__ zf, int8_t *__restrict__ filt, int ny, int nx,
            dtm_prm p) {
  extern __shared__ float sm[];
  float *D = sm;
  const int tid = threadIdx.x;
  const int r0 = blockIdx.y * DF_T, c0 = blockIdx.x * DF_T;
  // the selection inputs are requested first so that their latency hides behind the window load
  const int x = tid & 63, y0 = (tid >> 6) * DF_RPT;
  const int c = c0 + x;
  unsigned long long srcp = 0, fp = 0;   // 4 bits per cell: source (9 / 7 / 5 / 3 smoothing, 0 = the DEM itself), filter code
  if (c < nx) {
#pragma unroll 4
    for (int j = 0; j < DF_RPT; j++) {
      const int r = r0 + y0 + j;
      if (r >= ny) break;
      const size_t g = (size_t)r * nx + c;
      const float cv = __ldg(curv + g);
      const int s = (int)__ldg(slope0 + g);
      const bool peak = cv < -p.cb1, peakfill = cv >= -p.cb1 && cv < -p.cb0;
      const bool valley = cv > p.cb1, valleyfill = cv <= p.cb1 && cv > p.cb0;
      const bool steep = s >= p.sb4;
      int src = 9, f = 9;
      if (s >= p.sb1 && s <= p.sb2) { f = 7; src = 7; }
      if (s >= p.sb2 && s <= p.sb3) { f = 5; src = 5; }
      if ((s >= p.sb3 && s <= p.sb4) || valleyfill || peakfill) { f = 3; src = 3; }
      if (steep) f = 1;
      if (peak) f = 2;
      if (valley) f = 0;
      if (steep || peak || valley) src = 0;
      if (swo && swo[g]) { f = 3; src = 3; }
      if (edm && edm[g] == 5 && f < 3) { f = 3; src = 3; }
      if (tx && tx[g]) { f = 1; src = 0; }
      if (artifact && artifact[g] == 1) src = 0;
      srcp |= (unsigned long long)src << (4 * j);
      fp |= (unsigned long long)f << (4 * j);
    }
  }
  dtm_load_window(dem, D, r0, c0, ny, nx);
  __syncthreads();
  // which smoothings does this warp need at all?  (a warp whose cells all keep the DEM or share one filter skips the rest)
  u32 need = 0;
#pragma unroll
  for (int j = 0; j < DF_RPT; j++) need |= 1u << ((srcp >> (4 * j)) & 15u);
  need = __reduce_or_sync(0xFFFFFFFFu, need);
  float out[DF_RPT], acc[DF_RPT];
#pragma unroll
  for (int j = 0; j < DF_RPT; j++) out[j] = D[(y0 + j + 4) * DF_DW + x + 4];
  if (need & (1u << 9)) {
    conv_cols_rt<9>(D, 0, c_kern, x, y0, acc);
#pragma unroll
    for (int j = 0; j < DF_RPT; j++) if (((srcp >> (4 * j)) & 15u) == 9u) out[j] = acc[j];
  }
  if (need & (1u << 7)) {
    conv_cols_rt<7>(D, 1, c_kern + 81, x, y0, acc);
#pragma unroll
    for (int j = 0; j < DF_RPT; j++) if (((srcp >> (4 * j)) & 15u) == 7u) out[j] = acc[j];
  }
  if (need & (1u << 5)) {
    conv_cols_rt<5>(D, 2, c_kern + 81 + 49, x, y0, acc);
#pragma unroll
    for (int j = 0; j < DF_RPT; j++) if (((srcp >> (4 * j)) & 15u) == 5u) out[j] = acc[j];
  }
  if (need & (1u << 3)) {
    conv_cols_rt<3>(D, 3, c_kern + 81 + 49 + 25, x, y0, acc);
#pragma unroll
    for (int j = 0; j < DF_RPT; j++) if (((srcp >> (4 * j)) & 15u) == 3u) out[j] = acc[j];
  }
  if (c >= nx) return;
#pragma unroll
  for (int j = 0; j < DF_RPT; j++) {
    const int r = r0 + y0 + j;
    if (r >= ny) break;
    zf[(size_t)r * nx + c] = out[j];
    filt[(size_t)r * nx + c] = (int8_t)((fp >> (4 * j)) & 15u);
  }
}

static int dtm_core_setup(hc_ctx *ctx, const float *kernels_host, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    HC_CUDA(cudaFuncSetAttribute(k_dtm_pre, cudaFuncAttributeMaxDynamicSharedMemorySize, DF_PRE_SMEM));
    attr = true;
  }
  HC_CUDA(cudaMemcpyToSymbolAsync(c_kern, kernels_host, sizeof(float) * (81 + 49 + 25 + 9), 0, cudaMemcpyHostToDevice, st));
  (void)ctx;
  return HC_OK;
}

// First pass of the fused core: curvature_D2 (geometric, float32) and slope_D4 / slope_D2 (int8 degrees) of the 5x5-smoothed
// DEM (models/dem_noise_removal.py:405-410) without the smoothed raster ever reaching HBM, plus sums[0] = sum and
// sums[1] = sum of squares (float64, HOST pointer) of the curvature over rows [row_lo, row_hi) -- np.std(curvature) of
// :413.  curv / slope0 may be NULL (sums only).  kernels_host: the 9x9, 7x7, 5x5, 3x3 kernels back to back as for
// hc_conv2d_symm_dw4_dev.
extern "C" int hc_dtm_pre_dev(hc_ctx *ctx, const float *dem, const float *kernels_host, float *curv, int8_t *slope0,
                              int64_t ny, int64_t nx, double dx, double dy, int d4, int64_t row_lo, int64_t row_hi,
                              double *sums, void *stream) {
  ST_CTX(ctx);
  if (!sums) { hc_set_error("hc_dtm_pre: null pointer"); return HC_ERR_ARG; }
  sums[0] = sums[1] = 0.0;
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (!dem || !kernels_host || ny > 65535LL * DF_T) { hc_set_error("hc_dtm_pre: bad argument"); return HC_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  HC_TRY(arena_reset(ctx));
  double *acc = (double *)arena_take(ctx, 256);
  if (!acc) return HC_ERR_NOMEM;
  HC_TRY(dtm_core_setup(ctx, kernels_host, st));
  HC_CUDA(cudaMemsetAsync(acc, 0, 16, st));
  dtm_pre_prm P;
  memset(&P, 0, sizeof P);
  P.g = make_grad(dx, dy);
  P.d4 = d4 ? 1 : 0;
  P.row_lo = (int)(row_lo < 0 ? 0 : row_lo);
  P.row_hi = (int)(row_hi > ny ? ny : row_hi);
  deg_thr thr;
  thr.n = ctx->deg_thr_n;
  memcpy(thr.t, ctx->deg_thr, sizeof(thr.t));
  dim3 grid((unsigned)((nx + DF_T - 1) / DF_T), (unsigned)((ny + DF_T - 1) / DF_T));
  HC_PROF(ctx, st, "k_dtm_pre");
  k_dtm_pre<<<grid, 256, DF_PRE_SMEM, st>>>(dem, curv, slope0, acc, (int)ny, (int)nx, P, thr);
  HC_LAUNCH_CHECK(ctx);
  HC_CUDA(cudaMemcpyAsync(sums, acc, 16, cudaMemcpyDeviceToHost, st));
  HC_CUDA(cudaStreamSynchronize(st));
  return HC_OK;
}

// Second pass: the four DW smoothings (:399-402) and the selection block (:414-471) in one tile pass -> Zf and the filter
// code; curv / slope0 from hc_dtm_pre_dev, cbreak0/1 = c * sigma as np.float32 scalars, slopebreaks5 the five slope breaks.
extern "C" int hc_dtm_apply_dev(hc_ctx *ctx, const float *dem, const float *kernels_host, const float *curv,
                                const int8_t *slope0, const uint8_t *swo, const uint8_t *edm, const uint8_t *tx,
                                const uint8_t *artifact, float *zf, int8_t *filt, int64_t ny, int64_t nx, float cbreak0,
                                float cbreak1, const int *slopebreaks5, void *stream) {
  ST_CTX(ctx);
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (!dem || !kernels_host || !curv || !slope0 || !zf || !filt || !slopebreaks5 || dem == zf || ny > 65535LL * DF_T) {
    hc_set_error("hc_dtm_apply: bad argument");
    return HC_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  HC_TRY(dtm_core_setup(ctx, kernels_host, st));
  dtm_prm p;
  p.cb0 = cbreak0; p.cb1 = cbreak1;
  p.sb1 = slopebreaks5[1]; p.sb2 = slopebreaks5[2]; p.sb3 = slopebreaks5[3]; p.sb4 = slopebreaks5[4];
  dim3 grid((unsigned)((nx + DF_T - 1) / DF_T), (unsigned)((ny + DF_T - 1) / DF_T));
  HC_PROF(ctx, st, "k_dtm_apply");
  k_dtm_apply<<<grid, 256, DF_APP_SMEM, st>>>(dem, curv, slope0, swo, edm, tx, artifact, zf, filt, (int)ny, (int)nx, p);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}
