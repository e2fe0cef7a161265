// Water stencils of tools/water.py called from DepHandle.depression_handling (models/depression_handling.py:647,
// 889, 898) and burn_rivers (models/burn_rivers.py:841, 856): river_minimize :786, smooth_rivers :928,
// raise_small_streams :960, raise_small_rivers :1010.  Each is a chain of mask predicates, square moving-window
// minimum / maximum filters (scipy.ndimage.minimum_filter / maximum_filter, size=k, default mode 'reflect' =
// half-sample symmetric), an optional 5x5 mean, and a per-cell select.  min / max are exact, so everything but the
// 5x5 mean is bit-exact; the mean is accumulated in float64 like scipy does for a float64 kernel and narrowed to
// float32, which hides the summation order except at float32 rounding ties (about 0.3 % of cells for an equal-weight
// mean of float32 values: one ulp, decided by the summation order of the scipy build -- tolerance class).
//
// The reference guards each body with "if the masks are non-empty"; an empty mask makes the body a no-op (the
// test raster stays 1e6 / the apply mask stays 0), so the guard is not needed for the result and no count is
// read back.  NaN elevations are outside the contract (the reference replaces them only after river_minimize).
#include "common.cuh"

#define W_BIG 1e6f   // the "never wins a minimum" value of the reference (1e6, exact in float32)

__device__ __forceinline__ int wsymm(int i, int n) {
  while (i < 0 || i >= n) i = i < 0 ? -i - 1 : 2 * n - i - 1;
  return i;
}

template <typename T, bool IS_MAX> __device__ __forceinline__ T mm_op(T a, T b) {
  return IS_MAX ? (b > a ? b : a) : (b < a ? b : a);
}

// horizontal pass: out[r, c] = op over in[r, c - lo .. c + hi]
template <typename T, bool IS_MAX>
__global__ void __launch_bounds__(256)
k_mm_h(const T *__restrict__ in, T *__restrict__ out, int ny, int nx, int lo, int hi) {
  int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c >= nx) return;
  const T *row = in + (size_t)r * nx;
  T v = row[c];
  if (c - lo >= 0 && c + hi < nx) {
    const T *p = row + c - lo;
    int k = lo + hi + 1;
#pragma unroll 8
    for (int j = 0; j < k; j++) v = mm_op<T, IS_MAX>(v, p[j]);
  } else {
    for (int j = -lo; j <= hi; j++) v = mm_op<T, IS_MAX>(v, row[wsymm(c + j, nx)]);
  }
  out[(size_t)r * nx + c] = v;
}

// vertical pass; a (64 x 16) block keeps its (16 + k - 1) source rows in L1
template <typename T, bool IS_MAX>
__global__ void __launch_bounds__(1024)
k_mm_v(const T *__restrict__ in, T *__restrict__ out, int ny, int nx, int lo, int hi) {
  int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y * blockDim.y + threadIdx.y;
  if (c >= nx || r >= ny) return;
  T v = in[(size_t)r * nx + c];
  if (r - lo >= 0 && r + hi < ny) {
    const T *p = in + (size_t)(r - lo) * nx + c;
    int k = lo + hi + 1;
#pragma unroll 8
    for (int j = 0; j < k; j++) v = mm_op<T, IS_MAX>(v, p[(size_t)j * nx]);
  } else {
    for (int j = -lo; j <= hi; j++) v = mm_op<T, IS_MAX>(v, in[(size_t)wsymm(r + j, ny) * nx + c]);
  }
  out[(size_t)r * nx + c] = v;
}

// Shared-memory form of the two passes (windows up to a few hundred cells): a block stages its span plus the k - 1 halo
// cells once, then builds "minimum of the next 2^j cells" tables by doubling (log2 k steps of one op per cell, ping-pong
// between two buffers); a window of k cells is the op of two overlapping 2^p spans (p = floor(log2 k)).  ~3 log2 k
// shared-memory accesses per cell instead of k loads, and every input cell is read from global memory once per pass
// (plus the halo).  Tail cells whose span would leave the staged range are never a dependency of a stored result.
#define MMH_W 2048      // outputs per block, horizontal pass
#define MMV_R 128       // output rows per block, vertical pass

template <typename T, bool IS_MAX>
__global__ void __launch_bounds__(256)
k_mm_h_sm(const T *__restrict__ in, T *__restrict__ out, int ny, int nx, int lo, int k) {
  extern __shared__ __align__(16) unsigned char mm_smem[];
  const int len = MMH_W + k - 1;
  T *cur = (T *)mm_smem, *nxt = cur + len;
  const int c0 = blockIdx.x * MMH_W, r = blockIdx.y;
  const T *row = in + (size_t)r * nx;
  for (int i = threadIdx.x; i < len; i += 256) cur[i] = row[wsymm(c0 - lo + i, nx)];
  __syncthreads();
  int P = 1;
  for (int s = 1; 2 * s <= k; s *= 2) {
    for (int i = threadIdx.x; i < len; i += 256) nxt[i] = mm_op<T, IS_MAX>(cur[i], cur[min(i + s, len - 1)]);
    __syncthreads();
    T *t = cur; cur = nxt; nxt = t;
    P = 2 * s;
  }
  for (int j = threadIdx.x; j < MMH_W && c0 + j < nx; j += 256)
    out[(size_t)r * nx + c0 + j] = mm_op<T, IS_MAX>(cur[j], cur[j + k - P]);
}

// TC columns x MMV_R rows per block; threads (TC, 256 / TC): shared-memory rows are TC cells wide, so the doubling
// along rows is conflict free
template <typename T, bool IS_MAX, int TC>
__global__ void __launch_bounds__(256)
k_mm_v_sm(const T *__restrict__ in, T *__restrict__ out, int ny, int nx, int lo, int k) {
  extern __shared__ __align__(16) unsigned char mm_smem[];
  const int L = MMV_R + k - 1, TY = 256 / TC;
  T *cur = (T *)mm_smem, *nxt = cur + (size_t)L * TC;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = blockIdx.x * TC + tx, r0 = blockIdx.y * MMV_R;
  const bool live = c < nx;
  for (int rr = ty; rr < L; rr += TY)
    cur[rr * TC + tx] = live ? in[(size_t)wsymm(r0 - lo + rr, ny) * nx + c] : (T)0;
  __syncthreads();
  int P = 1;
  for (int s = 1; 2 * s <= k; s *= 2) {
    for (int rr = ty; rr < L; rr += TY)
      nxt[rr * TC + tx] = mm_op<T, IS_MAX>(cur[rr * TC + tx], cur[min(rr + s, L - 1) * TC + tx]);
    __syncthreads();
    T *t = cur; cur = nxt; nxt = t;
    P = 2 * s;
  }
  if (!live) return;
  for (int j = ty; j < MMV_R && r0 + j < ny; j += TY)
    out[(size_t)(r0 + j) * nx + c] = mm_op<T, IS_MAX>(cur[j * TC + tx], cur[(j + k - P) * TC + tx]);
}

#define MM_SMEM_MAX (96 * 1024)

// square window of `size` cells, scipy's origin convention: [i - size/2, i + size - 1 - size/2]
template <typename T>
static int mm_filter(hc_ctx *ctx, const T *in, T *out, T *tmp, int ny, int nx, int size, int is_max, cudaStream_t st) {
  int lo = size / 2, hi = size - 1 - size / 2;
  constexpr int TC = sizeof(T) == 1 ? 128 : 32;
  size_t smem_h = 2 * (size_t)(MMH_W + size - 1) * sizeof(T);
  size_t smem_v = 2 * (size_t)(MMV_R + size - 1) * TC * sizeof(T);
  if (smem_h <= MM_SMEM_MAX && smem_v <= MM_SMEM_MAX && ny <= 65535 * MMV_R) {
    // (per device and cheap: set on every call rather than cached, a process may drive several devices)
    if (is_max) {
      HC_CUDA(cudaFuncSetAttribute(k_mm_h_sm<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_SMEM_MAX));
      HC_CUDA(cudaFuncSetAttribute(k_mm_v_sm<T, true, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_SMEM_MAX));
    } else {
      HC_CUDA(cudaFuncSetAttribute(k_mm_h_sm<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_SMEM_MAX));
      HC_CUDA(cudaFuncSetAttribute(k_mm_v_sm<T, false, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_SMEM_MAX));
    }
    dim3 gh((unsigned)((nx + MMH_W - 1) / MMH_W), (unsigned)ny);
    dim3 bv(TC, 256 / TC), gv((unsigned)((nx + TC - 1) / TC), (unsigned)((ny + MMV_R - 1) / MMV_R));
    HC_PROF(ctx, st, "k_mm_h_sm");
    if (is_max) k_mm_h_sm<T, true><<<gh, 256, smem_h, st>>>(in, tmp, ny, nx, lo, size);
    else k_mm_h_sm<T, false><<<gh, 256, smem_h, st>>>(in, tmp, ny, nx, lo, size);
    HC_LAUNCH_CHECK(ctx);
    HC_PROF(ctx, st, "k_mm_v_sm");
    if (is_max) k_mm_v_sm<T, true, TC><<<gv, bv, smem_v, st>>>(tmp, out, ny, nx, lo, size);
    else k_mm_v_sm<T, false, TC><<<gv, bv, smem_v, st>>>(tmp, out, ny, nx, lo, size);
    HC_LAUNCH_CHECK(ctx);
    return HC_OK;
  }
  // very large windows: direct form
  dim3 gh((unsigned)((nx + 255) / 256), (unsigned)ny);
  dim3 bv(64, 16), gv((unsigned)((nx + 63) / 64), (unsigned)((ny + 15) / 16));
  HC_PROF(ctx, st, "k_mm_h");
  if (is_max) k_mm_h<T, true><<<gh, 256, 0, st>>>(in, tmp, ny, nx, lo, hi);
  else k_mm_h<T, false><<<gh, 256, 0, st>>>(in, tmp, ny, nx, lo, hi);
  HC_LAUNCH_CHECK(ctx);
  HC_PROF(ctx, st, "k_mm_v");
  if (is_max) k_mm_v<T, true><<<gv, bv, 0, st>>>(tmp, out, ny, nx, lo, hi);
  else k_mm_v<T, false><<<gv, bv, 0, st>>>(tmp, out, ny, nx, lo, hi);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

#define W_CTX(ctx) HC_ENTER(ctx)

static bool w_dims_ok(int64_t ny, int64_t nx, int size) {
  return ny > 0 && nx > 0 && ny <= 65535 && nx < (1 << 30) && size >= 1 && size <= 1025;
}

extern "C" int hc_minmax_filter_dev(hc_ctx *ctx, const void *in, void *out, int64_t ny, int64_t nx, int size,
                                    int is_max, int elem_bytes, void *stream) {
  W_CTX(ctx);
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (!in || !out || !w_dims_ok(ny, nx, size) || (elem_bytes != 1 && elem_bytes != 4)) {
    hc_set_error("hc_minmax_filter: bad argument");
    return HC_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  HC_TRY(arena_reset(ctx));
  void *tmp = arena_take(ctx, (size_t)ny * nx * elem_bytes);
  if (!tmp) return HC_ERR_NOMEM;
  if (elem_bytes == 1)
    return mm_filter<uint8_t>(ctx, (const uint8_t *)in, (uint8_t *)out, (uint8_t *)tmp, (int)ny, (int)nx, size, is_max, st);
  return mm_filter<float>(ctx, (const float *)in, (float *)out, (float *)tmp, (int)ny, (int)nx, size, is_max, st);
}

// ---- k x k convolution accumulated in float64 (apply_convolve_filter with a float64 kernel, tools/derive.py:94-97) ----
#define WC_MAXK 15
__constant__ double c_wkern[WC_MAXK * WC_MAXK];

__global__ void __launch_bounds__(256)
k_conv_f64(const float *__restrict__ in, float *__restrict__ out, int ny, int nx, int kh, int kw) {
  int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c >= nx) return;
  int ry = kh / 2, rx = kw / 2;
  double acc = 0.0;
  // out[r, c] = sum_{a, b} kern[a, b] * in[r + ry - a, c + rx - b]  (true convolution: the kernel is flipped)
  if (r >= ry && r + ry < ny && c >= rx && c + rx < nx) {   // interior: same taps in the same order, no reflection
    for (int a = 0; a < kh; a++) {
      const float *row = in + (size_t)(r + ry - a) * nx + c + rx;
      for (int b = 0; b < kw; b++) acc = __dadd_rn(acc, __dmul_rn(c_wkern[a * kw + b], (double)row[-b]));
    }
  } else {
    for (int a = 0; a < kh; a++) {
      const float *row = in + (size_t)wsymm(r + ry - a, ny) * nx;
      for (int b = 0; b < kw; b++)
        acc = __dadd_rn(acc, __dmul_rn(c_wkern[a * kw + b], (double)row[wsymm(c + rx - b, nx)]));
    }
  }
  out[(size_t)r * nx + c] = (float)acc;
}

static int conv_f64(hc_ctx *ctx, const float *in, const double *kernel_host, int kh, int kw, float *out, int ny, int nx,
                    cudaStream_t st) {
  HC_CUDA(cudaMemcpyToSymbolAsync(c_wkern, kernel_host, sizeof(double) * kh * kw, 0, cudaMemcpyHostToDevice, st));
  dim3 grid((unsigned)((nx + 255) / 256), (unsigned)ny);
  HC_PROF(ctx, st, "k_conv_f64");
  k_conv_f64<<<grid, 256, 0, st>>>(in, out, ny, nx, kh, kw);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

static int mean5_f64(hc_ctx *ctx, const float *in, float *out, int ny, int nx, cudaStream_t st) {
  // kernel = (1. / np.sum(np.ones((5, 5)))) * np.ones((5, 5))  (tools/water.py:939-940, 983-984)
  static double k25[25];
  for (int i = 0; i < 25; i++) k25[i] = (1. / 25.0) * 1.0;
  return conv_f64(ctx, in, k25, 5, 5, out, ny, nx, st);
}

extern "C" int hc_conv2d_symm_f64_dev(hc_ctx *ctx, const float *in, const double *kernel_host, int kh, int kw,
                                      float *out, int64_t ny, int64_t nx, void *stream) {
  W_CTX(ctx);
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (!in || !out || !kernel_host || kh < 1 || kw < 1 || kh > WC_MAXK || kw > WC_MAXK || !(kh & 1) || !(kw & 1) ||
      ny > 65535) {
    hc_set_error("hc_conv2d_symm_f64: bad argument (odd kernel sizes <= 15)");
    return HC_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  HC_TRY(conv_f64(ctx, in, kernel_host, kh, kw, out, (int)ny, (int)nx, st));
  HC_CUDA(cudaStreamSynchronize(st));  // kernel_host is the caller's memory
  return HC_OK;
}

// ---- river_minimize (tools/water.py:786-847) -------------------------------------------------------------------
__global__ void k_rm_class(const uint8_t *__restrict__ wm, uint8_t *__restrict__ t0, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint8_t w = wm[i];
    t0[i] = (w > 2 && w < 8) ? 1 : ((w > 0 && w < 3) ? 2 : 0);   // :818-819
  }
}
__global__ void k_rm_test(const uint8_t *__restrict__ test, const float *__restrict__ dem, float nodata,
                          float *__restrict__ tdem, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float z = dem[i];
    tdem[i] = (test[i] == 1 && z != nodata) ? z : W_BIG;           // :825-828
  }
}
__global__ void k_rm_final(const uint8_t *__restrict__ test, const float *__restrict__ dem,
                           const int16_t *__restrict__ swo, const float *__restrict__ minarr, float nodata,
                           float max_burn, float *__restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float z = dem[i];
    bool inner = test[i] == 1;
    float td = (inner && z != nodata) ? z : W_BIG;
    if (inner && swo[i] > 5) td = minarr[i];                       // :822-823, :833
    if (__fsub_rn(z, td) > max_burn) td = __fsub_rn(z, max_burn);  // :836-838
    float o = inner ? td : z;                                      // :840
    if (o == W_BIG) o = z;                                         // :841
    out[i] = o;
  }
}

extern "C" int hc_river_minimize_dev(hc_ctx *ctx, const float *dem, const uint8_t *wm, const int16_t *swo, float nodata,
                                     float max_burn, int filter_size, float *out, int64_t ny, int64_t nx, void *stream) {
  W_CTX(ctx);
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (!dem || !wm || !swo || !out || !w_dims_ok(ny, nx, filter_size)) {
    hc_set_error("hc_river_minimize: bad argument");
    return HC_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  size_t n = (size_t)ny * nx;
  HC_TRY(arena_reset(ctx));
  uint8_t *t0 = (uint8_t *)arena_take(ctx, n), *t1 = (uint8_t *)arena_take(ctx, n), *test = (uint8_t *)arena_take(ctx, n);
  float *tdem = (float *)arena_take(ctx, n * 4), *ftmp = (float *)arena_take(ctx, n * 4),
        *minarr = (float *)arena_take(ctx, n * 4);
  if (!t0 || !t1 || !test || !tdem || !ftmp || !minarr) return HC_ERR_NOMEM;
  int grid = ctx->sm_count * 8;
  HC_PROF(ctx, st, "k_rm_class");
  k_rm_class<<<grid, 256, 0, st>>>(wm, t0, n);
  HC_LAUNCH_CHECK(ctx);
  HC_TRY(mm_filter<uint8_t>(ctx, t0, test, t1, (int)ny, (int)nx, 3, 1, st));            // :820
  HC_PROF(ctx, st, "k_rm_test");
  k_rm_test<<<grid, 256, 0, st>>>(test, dem, nodata, tdem, n);
  HC_LAUNCH_CHECK(ctx);
  HC_TRY(mm_filter<float>(ctx, tdem, minarr, ftmp, (int)ny, (int)nx, filter_size, 0, st));  // :831-832
  HC_PROF(ctx, st, "k_rm_final");
  k_rm_final<<<grid, 256, 0, st>>>(test, dem, swo, minarr, nodata, max_burn, out, n);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

// ---- shared tail of the three "raise / lower toward the moving minimum" functions ---------------------------------
// mode 0 raise_small_rivers :1038-1043, mode 1 raise_small_streams :998-1005: where `target` and apply == 1,
//   d = burn - minf; d <= extra && d >= -max_raise  ->  burn = minf + extra
// mode 2 smooth_rivers :947-955: where target, d = burn - minf; d >= 0.01 && d <= max_lower -> burn = minf
// (dem and burn may be the same raster: no __restrict__ on them)
__global__ void k_w_apply(int mode, const float *dem, float *burn,
                          const uint8_t *__restrict__ wm, const int16_t *__restrict__ swo,
                          const uint8_t *__restrict__ apply, const float *__restrict__ minf, float nodata, int swo_cut,
                          float p0, float p1, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint8_t w = wm[i];
    bool valid = dem[i] != nodata;
    bool target;
    if (mode == 0) target = w > 0 && w < 8 && valid && swo[i] < swo_cut && apply[i] == 1;          // :1015-1016, :1032
    else if (mode == 1) target = w > 7 && w != 11 && w != 15 && valid && apply[i] == 1;            // :990-992
    else target = w > 2 && w < 8 && valid && swo[i] > swo_cut;                                     // :932-933
    if (!target) continue;
    float m = minf[i], d = __fsub_rn(burn[i], m);
    if (mode == 2) {
      if (d >= 0.01f && d <= p1) burn[i] = m;
    } else {
      if (d <= p0 && d >= -p1) burn[i] = __fadd_rn(m, p0);
    }
  }
}

// big-river mask + test raster: mask = wm in (lo_excl, 8) & valid & swo > swo_cut (swo may be null: no SWO test);
// test = mask ? src : 1e6
__global__ void k_w_big(const float *dem, const float *src, const uint8_t *__restrict__ wm,
                        const int16_t *__restrict__ swo, const uint8_t *__restrict__ excl, float nodata, int wm_lo,
                        int swo_cut, uint8_t *__restrict__ mask, float *__restrict__ tdem, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint8_t w = wm[i];
    bool m = w > wm_lo && w < 8 && dem[i] != nodata;
    if (swo) m = m && swo[i] > swo_cut;
    if (excl) m = m && excl[i] == 0;
    if (mask) mask[i] = m ? 1 : 0;
    tdem[i] = m ? src[i] : W_BIG;
  }
}

__global__ void k_w_stream_mask(const float *__restrict__ dem, const uint8_t *__restrict__ wm, float nodata,
                                uint8_t *__restrict__ mask, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint8_t w = wm[i];
    mask[i] = (w > 7 && w != 11 && w != 15 && dem[i] != nodata) ? 1 : 0;   // :964-966
  }
}

extern "C" int hc_raise_small_rivers_dev(hc_ctx *ctx, const float *dem, float *burn, const uint8_t *wm,
                                         const int16_t *swo, float nodata, int swo_high, int swo_low, float extra_raise,
                                         float max_raise, int filter_size, int64_t ny, int64_t nx, void *stream) {
  W_CTX(ctx);
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (!dem || !burn || !wm || !swo || !w_dims_ok(ny, nx, filter_size) || filter_size - 16 < 1) {
    hc_set_error("hc_raise_small_rivers: bad argument");
    return HC_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  size_t n = (size_t)ny * nx;
  HC_TRY(arena_reset(ctx));
  uint8_t *mask = (uint8_t *)arena_take(ctx, n), *t1 = (uint8_t *)arena_take(ctx, n), *apply = (uint8_t *)arena_take(ctx, n);
  float *tdem = (float *)arena_take(ctx, n * 4), *ftmp = (float *)arena_take(ctx, n * 4),
        *minf = (float *)arena_take(ctx, n * 4);
  if (!mask || !t1 || !apply || !tdem || !ftmp || !minf) return HC_ERR_NOMEM;
  int grid = ctx->sm_count * 8;
  HC_PROF(ctx, st, "k_w_big");
  k_w_big<<<grid, 256, 0, st>>>(dem, burn, wm, swo, nullptr, nodata, 2, swo_high, mask, tdem, n);   // :1017-1018, :1025-1026
  HC_LAUNCH_CHECK(ctx);
  HC_TRY(mm_filter<float>(ctx, tdem, minf, ftmp, (int)ny, (int)nx, filter_size, 0, st));            // :1027-1028
  HC_TRY(mm_filter<uint8_t>(ctx, mask, apply, t1, (int)ny, (int)nx, filter_size - 16, 1, st));      // :1032-1035
  HC_PROF(ctx, st, "k_w_apply");
  k_w_apply<<<grid, 256, 0, st>>>(0, dem, burn, wm, swo, apply, minf, nodata, swo_low, extra_raise, max_raise, n);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

extern "C" int hc_raise_small_streams_dev(hc_ctx *ctx, const float *dem, float *burn, const uint8_t *wm, float nodata,
                                          float extra_raise, float max_raise, int fsize, int64_t ny, int64_t nx,
                                          void *stream) {
  W_CTX(ctx);
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (!dem || !burn || !wm || !w_dims_ok(ny, nx, fsize) || fsize - 16 < 1) {
    hc_set_error("hc_raise_small_streams: bad argument");
    return HC_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  size_t n = (size_t)ny * nx;
  HC_TRY(arena_reset(ctx));
  uint8_t *mask = (uint8_t *)arena_take(ctx, n), *t1 = (uint8_t *)arena_take(ctx, n), *sbuf = (uint8_t *)arena_take(ctx, n),
          *apply = (uint8_t *)arena_take(ctx, n);
  float *smooth = (float *)arena_take(ctx, n * 4), *tdem = (float *)arena_take(ctx, n * 4),
        *ftmp = (float *)arena_take(ctx, n * 4), *minf = (float *)arena_take(ctx, n * 4);
  if (!mask || !t1 || !sbuf || !apply || !smooth || !tdem || !ftmp || !minf) return HC_ERR_NOMEM;
  int grid = ctx->sm_count * 8;
  HC_PROF(ctx, st, "k_w_stream_mask");
  k_w_stream_mask<<<grid, 256, 0, st>>>(dem, wm, nodata, mask, n);
  HC_LAUNCH_CHECK(ctx);
  HC_TRY(mm_filter<uint8_t>(ctx, mask, sbuf, t1, (int)ny, (int)nx, 5, 1, st));                      // :975-978
  HC_TRY(mean5_f64(ctx, dem, smooth, (int)ny, (int)nx, st));                                        // :983-985
  HC_PROF(ctx, st, "k_w_big");
  k_w_big<<<grid, 256, 0, st>>>(dem, smooth, wm, nullptr, sbuf, nodata, 3, 0, mask, tdem, n);       // :979-980, :987-988
  HC_LAUNCH_CHECK(ctx);
  HC_TRY(mm_filter<float>(ctx, tdem, minf, ftmp, (int)ny, (int)nx, fsize, 0, st));                  // :989
  HC_TRY(mm_filter<uint8_t>(ctx, mask, apply, t1, (int)ny, (int)nx, fsize - 16, 1, st));            // :991-994
  HC_PROF(ctx, st, "k_w_apply");
  k_w_apply<<<grid, 256, 0, st>>>(1, dem, burn, wm, nullptr, apply, minf, nodata, 0, extra_raise, max_raise, n);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

extern "C" int hc_smooth_rivers_dev(hc_ctx *ctx, const float *dem, float *burn, const uint8_t *wm, const int16_t *swo,
                                    float nodata, int swo_cut, float max_lower, int filter_size, int64_t ny, int64_t nx,
                                    void *stream) {
  W_CTX(ctx);
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (!dem || !burn || !wm || !swo || !w_dims_ok(ny, nx, filter_size)) {
    hc_set_error("hc_smooth_rivers: bad argument");
    return HC_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  size_t n = (size_t)ny * nx;
  HC_TRY(arena_reset(ctx));
  float *smooth = (float *)arena_take(ctx, n * 4), *tdem = (float *)arena_take(ctx, n * 4),
        *ftmp = (float *)arena_take(ctx, n * 4), *minf = (float *)arena_take(ctx, n * 4);
  if (!smooth || !tdem || !ftmp || !minf) return HC_ERR_NOMEM;
  int grid = ctx->sm_count * 8;
  HC_TRY(mean5_f64(ctx, burn, smooth, (int)ny, (int)nx, st));                                       // :939-942
  HC_PROF(ctx, st, "k_w_big");
  k_w_big<<<grid, 256, 0, st>>>(dem, smooth, wm, swo, nullptr, nodata, 2, swo_cut, nullptr, tdem, n);  // :944-945
  HC_LAUNCH_CHECK(ctx);
  HC_TRY(mm_filter<float>(ctx, tdem, minf, ftmp, (int)ny, (int)nx, filter_size, 0, st));            // :946-948
  HC_PROF(ctx, st, "k_w_apply");
  k_w_apply<<<grid, 256, 0, st>>>(2, dem, burn, wm, swo, nullptr, minf, nodata, swo_cut, 0.f, max_lower, n);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// wbt.fill_missing_data(i, output, filter, weight) as pydrodem calls it after blanking cells it wants re-interpolated
// (tools/water.py:1396 filter=11, :1774 filter=17, weight=2; models/burn_rivers.py:1283,1582; models/dem_noise_removal.py:739).
// WhiteboxTools' FillMissingData is not available here; restated from its documentation (parity unpinned): the sample
// points are the valid cells on the EDGE of nodata areas (a valid cell with a nodata 8-neighbour); a nodata cell takes
// the inverse-distance-weighted mean (weight 1 / d^w, d in cells) of the sample points within filter / 2 cells of it and
// stays nodata when there is none.  Sums in float64, result narrowed to float32.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_fmd_edges(const float *__restrict__ z, uint8_t *__restrict__ edge, int ny, int nx, float nodata) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)ny * nx) return;
  const int r = (int)(i / nx), c = (int)(i - (long long)r * nx);
  const float v = z[i];
  uint8_t e = 0;
  if (!(v == nodata || v != v)) {
    for (int dr = -1; dr <= 1 && !e; dr++)
      for (int dc = -1; dc <= 1; dc++) {
        const int rr = r + dr, cc = c + dc;
        if (rr < 0 || rr >= ny || cc < 0 || cc >= nx) continue;
        const float w = z[(size_t)rr * nx + cc];
        if (w == nodata || w != w) { e = 1; break; }
      }
  }
  edge[i] = e;
}

// one warp per nodata cell would be overkill for the few thousand cells pydrodem blanks: one thread per cell, the
// window walked row by row (edge flags are bytes, the window of a 17-cell filter is 289 of them)
__global__ void __launch_bounds__(256)
k_fmd_fill(const float *__restrict__ z, const uint8_t *__restrict__ edge, float *__restrict__ out, int ny, int nx,
           float nodata, int half, double radius, double weight) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)ny * nx) return;
  const float v = z[i];
  if (!(v == nodata || v != v)) { out[i] = v; return; }
  const int r = (int)(i / nx), c = (int)(i - (long long)r * nx);
  double sw = 0.0, sz = 0.0;
  for (int dr = -half; dr <= half; dr++) {
    const int rr = r + dr;
    if (rr < 0 || rr >= ny) continue;
    for (int dc = -half; dc <= half; dc++) {
      const int cc = c + dc;
      if (cc < 0 || cc >= nx) continue;
      const size_t g = (size_t)rr * nx + cc;
      if (!edge[g]) continue;
      const double d = sqrt((double)(dr * dr + dc * dc));
      if (d > radius) continue;
      const double w = 1.0 / (weight == 2.0 ? d * d : pow(d, weight));
      sw += w;
      sz += w * (double)z[g];
    }
  }
  out[i] = sw > 0.0 ? (float)(sz / sw) : v;
}

extern "C" int hc_fill_missing_data_dev(hc_ctx *ctx, const float *z, float *out, int64_t ny, int64_t nx, float nodata,
                                        int filter, double weight, void *stream) {
  W_CTX(ctx);
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (!z || !out || z == out || filter < 1 || filter > 1001 || ny * nx >= 2147483647LL) {
    hc_set_error("hc_fill_missing_data: bad argument");
    return HC_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)ny * nx;
  HC_TRY(arena_reset(ctx));
  uint8_t *edge = (uint8_t *)arena_take(ctx, n);
  if (!edge) return HC_ERR_NOMEM;
  const unsigned grid = (unsigned)((n + 255) / 256);
  HC_PROF(ctx, st, "k_fmd_edges");
  k_fmd_edges<<<grid, 256, 0, st>>>(z, edge, (int)ny, (int)nx, nodata);
  HC_LAUNCH_CHECK(ctx);
  HC_PROF(ctx, st, "k_fmd_fill");
  k_fmd_fill<<<grid, 256, 0, st>>>(z, edge, out, (int)ny, (int)nx, nodata, filter / 2, (double)filter / 2.0, weight);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Raster halves of wtr.enforce_monotonic_rivers (tools/water.py:1407-1779), one water level h at a time:
//   k_mono_mask    the candidate cells of level h: flattened-river cells (classes 4..7) still "left", standing above h
//                  or 8-adjacent to a river cell that does, but not exactly on h (:1488-1496)
//   k_mono_stats   per 4-connected clump of that mask: cell count (inside the inner boundary when one is given), whether
//                  it touches the boundary "donut", bounding box (:1505-1524, :1535-1538)
// The per-clump geometry tests that follow are host logic on small windows (pydrodem_b200/tools/water.py).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_mono_mask(const float *__restrict__ dem, const uint8_t *__restrict__ wm, const uint8_t *__restrict__ left, float h,
            uint8_t *__restrict__ out, int ny, int nx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)ny * nx) return;
  const int r = (int)(i / nx), c = (int)(i - (long long)r * nx);
  uint8_t o = 0;
  const uint8_t w = wm[i];
  if (w > 3 && w < 8 && left[i]) {
    const float d = __fsub_rn(dem[i], h);
    if (d != 0.f) {   // (NaN counts as "not zero", like numpy's ==)
      for (int dr = -1; dr <= 1 && !o; dr++)
        for (int dc = -1; dc <= 1; dc++) {
          const int rr = r + dr, cc = c + dc;
          if (rr < 0 || rr >= ny || cc < 0 || cc >= nx) continue;   // (symmetric padding only repeats cells of the window)
          const size_t g = (size_t)rr * nx + cc;
          const uint8_t wn = wm[g];
          if (wn > 3 && wn < 8 && __fsub_rn(dem[g], h) > 0.f) { o = 1; break; }
        }
    }
  }
  out[i] = o;
}

// stats[id] = {count, touches donut, rmin, rmax, cmin, cmax} (int32 x 6), ids 1..nlab; rmin / cmin start at INT_MAX
__global__ void __launch_bounds__(256)
k_mono_stats(const int32_t *__restrict__ lab, const uint8_t *__restrict__ donut, const uint8_t *__restrict__ inner,
             int32_t *__restrict__ stats, int ny, int nx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)ny * nx) return;
  const int32_t id = lab[i];
  if (id <= 0) return;
  const int r = (int)(i / nx), c = (int)(i - (long long)r * nx);
  int32_t *s = stats + (size_t)id * 6;
  if (!inner || inner[i] == 1) atomicAdd(&s[0], 1);
  if (donut && donut[i] == 1) atomicOr(&s[1], 1);
  atomicMin(&s[2], r); atomicMax(&s[3], r);
  atomicMin(&s[4], c); atomicMax(&s[5], c);
}

extern "C" int hc_mono_mask_dev(hc_ctx *ctx, const float *dem, const uint8_t *wm, const uint8_t *left, float h,
                                uint8_t *out, int64_t ny, int64_t nx, void *stream) {
  W_CTX(ctx);
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (!dem || !wm || !left || !out || ny * nx >= 2147483647LL) { hc_set_error("hc_mono_mask: bad argument"); return HC_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  HC_PROF(ctx, st, "k_mono_mask");
  k_mono_mask<<<(unsigned)(((size_t)ny * nx + 255) / 256), 256, 0, st>>>(dem, wm, left, h, out, (int)ny, (int)nx);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

extern "C" int hc_mono_stats_dev(hc_ctx *ctx, const int32_t *lab, const uint8_t *donut, const uint8_t *inner,
                                 int32_t nlab, int32_t *stats, int64_t ny, int64_t nx, void *stream) {
  W_CTX(ctx);
  if (ny <= 0 || nx <= 0 || nlab <= 0) return HC_OK;
  if (!lab || !stats || ny * nx >= 2147483647LL) { hc_set_error("hc_mono_stats: bad argument"); return HC_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  HC_PROF(ctx, st, "k_mono_stats");
  k_mono_stats<<<(unsigned)(((size_t)ny * nx + 255) / 256), 256, 0, st>>>(lab, donut, inner, stats, (int)ny, (int)nx);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}
