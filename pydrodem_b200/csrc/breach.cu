// breach.cu -- host-side depression breaching (Lindsay 2016), the arithmetic behind
//   wbt.breach_depressions(dem, out, max_length=radius, max_depth=None, fill_pits=True, flat_increment=0.0)
// as dep.breach_pits calls it (tools/depressions.py:1376-1380; also :1012-1014 inside combined_solution).
//
// breaching follows ONE global priority order, it does not shard and it is not one of the
// bandwidth-bound passes; "host C++ first is acceptable".  This file is that host implementation -- product code of
// the library (single thread, O(N log N)), not a fallback for anything: there is no GPU breach yet.  The final
// "fill what could not be breached" step is the library's ordinary fill (GPU, hc_fill_dev) when the caller asks for
// do_fill = 0 and runs it afterwards, or a host priority flood (do_fill = 1) so that the whole tool can be checked
// without a GPU.
//
// WhiteboxTools is not in /root/reference (un-vendored, version unpinned), so this restates the published algorithm
// (Lindsay 2016, "Efficient hybrid breaching-filling sink removal methods for flow path enforcement in digital
// elevation models", Hydrol. Process. 30) -- PARITY UNPINNED:
//   0. fill_pits: single-cell pits are raised to their lowest neighbour first (same rule as hc_fill_single_cell_pits).
//   1. Priority flood from the seed cells (valid cells on the raster frame or 8-adjacent to exterior-connected nodata --
//      the seed rule of the WBT fill, oracle/hydro_oracle.c orc_seed_mask), min-heap on (elevation, push order);
//      neighbours are scanned NE, E, SE, S, SW, W, NW, N.  A cell discovered from cell c gets the back-link c.
//   2. When a popped cell p is a pit (no valid neighbour is lower on the current surface), the breach path is the chain
//      of back-links from p toward the seeds: it is followed until a cell lower than z(p) -- or a seed, which drains off
//      the raster -- is reached.  If the walk is no longer than max_length cells and never has to cut deeper than
//      max_depth, every cell of it that is higher than z(p) is lowered to z(p) (flat_increment = 0: a level channel);
//      otherwise the pit is left for step 3.  Path cells were all popped earlier, so cutting never disturbs the heap.
//   3. Depressions that remain are filled (epsilon 0).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <queue>
#include <vector>

#include "common.cuh"

namespace {

const int B_DR[8] = {-1, 0, 1, 1, 1, 0, -1, -1};
const int B_DC[8] = {1, 1, 1, 0, -1, -1, -1, 0};

struct item {
  float z;
  uint64_t seq;
  int64_t i;
};
struct item_after {
  bool operator()(const item &a, const item &b) const { return a.z > b.z || (a.z == b.z && a.seq > b.seq); }
};
typedef std::priority_queue<item, std::vector<item>, item_after> heap_t;

inline bool b_nodata(float v, float nodata) { return v == nodata || v != v; }

// seeds of the WBT rule: frame cells and cells 8-adjacent to nodata connected to the frame
void seed_mask(const float *z, std::vector<uint8_t> &seed, int64_t ny, int64_t nx, float nodata) {
  const int64_t n = ny * nx;
  seed.assign((size_t)n, 0);
  std::vector<uint8_t> ext((size_t)n, 0);
  std::vector<int64_t> stack;
  for (int64_t r = 0; r < ny; r++)
    for (int64_t c = 0; c < nx; c++) {
      if (!(r == 0 || c == 0 || r == ny - 1 || c == nx - 1)) continue;
      int64_t i = r * nx + c;
      if (b_nodata(z[i], nodata)) {
        if (!ext[i]) { ext[i] = 1; stack.push_back(i); }
      } else {
        seed[i] = 1;
      }
    }
  while (!stack.empty()) {
    int64_t i = stack.back();
    stack.pop_back();
    int64_t r = i / nx, c = i % nx;
    for (int k = 0; k < 8; k++) {
      int64_t rr = r + B_DR[k], cc = c + B_DC[k];
      if (rr < 0 || cc < 0 || rr >= ny || cc >= nx) continue;
      int64_t j = rr * nx + cc;
      if (b_nodata(z[j], nodata)) {
        if (!ext[j]) { ext[j] = 1; stack.push_back(j); }
      } else {
        seed[j] = 1;
      }
    }
  }
}

void fill_single_cell_pits(const float *z, float *out, int64_t ny, int64_t nx, float nodata) {
  memcpy(out, z, sizeof(float) * (size_t)(ny * nx));
  for (int64_t r = 1; r < ny - 1; r++)
    for (int64_t c = 1; c < nx - 1; c++) {
      float v = z[r * nx + c];
      if (b_nodata(v, nodata)) continue;
      float mn = INFINITY;
      bool pit = true;
      for (int k = 0; k < 8 && pit; k++) {
        float zn = z[(r + B_DR[k]) * nx + c + B_DC[k]];
        if (b_nodata(zn, nodata) || !(zn > v)) pit = false;
        else if (zn < mn) mn = zn;
      }
      if (pit) out[r * nx + c] = mn;
    }
}

}  // namespace

// stats[0] pits met, [1] pits breached, [2] pits left to the fill, [3] cells lowered, [4] cells raised by the final fill
extern "C" int hc_breach_depressions_host(const float *z, float *out, int64_t ny, int64_t nx, float nodata,
                                          int64_t max_length, double max_depth, int fill_pits, int do_fill,
                                          int64_t *stats) {
  if (stats) memset(stats, 0, sizeof(int64_t) * 5);
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (!z || !out || z == out) { hc_set_error("hc_breach_depressions: bad argument"); return HC_ERR_ARG; }
  const int64_t n = ny * nx;
  const bool lim_len = max_length > 0;
  const bool lim_depth = max_depth > 0 && std::isfinite(max_depth);
  try {
    float *w = out;
    if (fill_pits) fill_single_cell_pits(z, w, ny, nx, nodata);
    else memcpy(w, z, sizeof(float) * (size_t)n);
    std::vector<uint8_t> seed;
    seed_mask(w, seed, ny, nx, nodata);
    std::vector<uint8_t> visited((size_t)n, 0);
    std::vector<int8_t> back((size_t)n, -1);   // direction slot toward the cell this one was discovered from
    heap_t heap;
    uint64_t seq = 0;
    for (int64_t i = 0; i < n; i++)
      if (seed[i]) { visited[i] = 1; heap.push(item{w[i], seq++, i}); }
    std::vector<int64_t> path;
    int64_t s_pits = 0, s_done = 0, s_left = 0, s_cut = 0;
    while (!heap.empty()) {
      item t = heap.top();
      heap.pop();
      const int64_t p = t.i, r = p / nx, c = p % nx;
      if (!seed[p]) {
        const float zt = w[p];
        bool pit = true;
        for (int k = 0; k < 8 && pit; k++) {
          int64_t rr = r + B_DR[k], cc = c + B_DC[k];
          if (rr < 0 || cc < 0 || rr >= ny || cc >= nx) continue;
          float zn = w[rr * nx + cc];
          if (!b_nodata(zn, nodata) && zn < zt) pit = false;
        }
        if (pit) {
          s_pits++;
          path.clear();
          int64_t q = p, len = 0;
          double depth = 0.0;
          bool ok = true;
          while (back[q] >= 0) {
            int k = back[q];
            q = (q / nx + B_DR[k]) * nx + (q % nx + B_DC[k]);
            len++;
            if (lim_len && len > max_length) { ok = false; break; }
            if (w[q] < zt) break;                       // lower ground reached: the pit drains
            double d = (double)w[q] - (double)zt;
            if (d > depth) depth = d;
            if (lim_depth && depth > max_depth) { ok = false; break; }
            if (w[q] > zt) path.push_back(q);
          }
          if (ok) {
            s_done++;
            for (int64_t qq : path) w[qq] = zt;
            s_cut += (int64_t)path.size();
          } else {
            s_left++;
          }
        }
      }
      for (int k = 0; k < 8; k++) {
        int64_t rr = r + B_DR[k], cc = c + B_DC[k];
        if (rr < 0 || cc < 0 || rr >= ny || cc >= nx) continue;
        int64_t j = rr * nx + cc;
        if (visited[j] || b_nodata(w[j], nodata)) continue;
        visited[j] = 1;
        back[j] = (int8_t)((k + 4) & 7);
        heap.push(item{w[j], seq++, j});
      }
    }
    int64_t s_raised = 0;
    if (do_fill) {
      // priority-flood fill (epsilon 0) of what is left, same seeds
      std::fill(visited.begin(), visited.end(), 0);
      seq = 0;
      for (int64_t i = 0; i < n; i++)
        if (seed[i]) { visited[i] = 1; heap.push(item{w[i], seq++, i}); }
      while (!heap.empty()) {
        item t = heap.top();
        heap.pop();
        const int64_t r = t.i / nx, c = t.i % nx;
        for (int k = 0; k < 8; k++) {
          int64_t rr = r + B_DR[k], cc = c + B_DC[k];
          if (rr < 0 || cc < 0 || rr >= ny || cc >= nx) continue;
          int64_t j = rr * nx + cc;
          if (visited[j] || b_nodata(w[j], nodata)) continue;
          visited[j] = 1;
          if (w[j] < t.z) { w[j] = t.z; s_raised++; }
          heap.push(item{w[j], seq++, j});
        }
      }
    }
    if (stats) { stats[0] = s_pits; stats[1] = s_done; stats[2] = s_left; stats[3] = s_cut; stats[4] = s_raised; }
  } catch (const std::bad_alloc &) {
    hc_set_error("hc_breach_depressions: out of host memory");
    return HC_ERR_NOMEM;
  }
  return HC_OK;
}


// ---- flow-path formulation (specification: oracle/breach_flowpath.py) ---------------------------
// Order-independent breach: every pit is breached along the flow path of the FILLED surface.  The caller supplies
//   w      the surface to breach (after the optional single-cell-pit fill),
//   filled hc_fill(w, seed rule WBT),
//   dir    hc_d8 + hc_resolve_flats (table WBT: codes 1, 2, 4 ... 128 = NE, E, SE, S, SW, W, NW, N) of `filled` WITH
//          EVERY SEED LOWERED BY ONE FLOAT32 ULP (hc_wbt_seed_mask_host gives the seeds): a flat whose only way out is
//          a seed on its own level then resolves toward that seed and every non-seed cell has a pointer,
// all three from the device passes; this function does the part that is pointer chasing: one walk per pit with the
// length / depth budgets measured on the original surface, cuts combined with min().  out = the cut surface (fill it
// afterwards).
// EXPERIMENTAL: host reference of the walk a device kernel (one thread per pit) will replace; not yet the default of
// wbt.breach_depressions.  stats (4 x int64, nullable): pits, breached, left, cells lowered.
extern "C" int hc_breach_flowpath_host(const float *w, const float *filled, const uint8_t *dir, float *out, int64_t ny,
                                       int64_t nx, float nodata, int64_t max_length, double max_depth, int64_t *stats) {
  if (stats) memset(stats, 0, sizeof(int64_t) * 4);
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (!w || !filled || !dir || !out || w == out) { hc_set_error("hc_breach_flowpath: bad argument"); return HC_ERR_ARG; }
  const int64_t n = ny * nx;
  const bool lim_len = max_length > 0;
  const bool lim_depth = max_depth > 0 && std::isfinite(max_depth);
  try {
    std::vector<uint8_t> seed;
    seed_mask(w, seed, ny, nx, nodata);
    const uint8_t *d = dir;
    // (2) the walks
    memcpy(out, w, sizeof(float) * (size_t)n);
    int64_t s_pits = 0, s_done = 0, s_left = 0;
    std::vector<int64_t> path;
    for (int64_t p = 0; p < n; p++) {
      const float zt = w[p];
      if (seed[p] || b_nodata(zt, nodata)) continue;
      const int64_t r = p / nx, c = p % nx;
      bool pit = true;
      for (int k = 0; k < 8 && pit; k++) {
        int64_t rr = r + B_DR[k], cc = c + B_DC[k];
        if (rr < 0 || cc < 0 || rr >= ny || cc >= nx) continue;
        float zn = w[rr * nx + cc];
        if (!b_nodata(zn, nodata) && zn < zt) pit = false;
      }
      if (!pit) continue;
      s_pits++;
      path.clear();
      int64_t q = p, steps = 0;
      double depth = 0.0;
      bool ok = true;
      for (;;) {
        if (seed[q]) break;                          // the path leaves the raster here
        uint8_t code = d[q];
        if (code == 0 || (code & (code - 1))) {      // no pointer (or the nodata code 255) on a non-seed cell
          hc_set_error("hc_breach_flowpath: cell %lld of the filled surface has no pointer", (long long)q);
          return HC_ERR_ARG;
        }
        int k = 0;
        while (!((code >> k) & 1)) k++;
        int64_t rr = q / nx + B_DR[k], cc = q % nx + B_DC[k];
        if (rr < 0 || cc < 0 || rr >= ny || cc >= nx || ++steps > n) { hc_set_error("hc_breach_flowpath: pointer leaves the raster / cycles"); return HC_ERR_ARG; }
        q = rr * nx + cc;
        if (lim_len && steps > max_length) { ok = false; break; }
        if (w[q] < zt) break;
        double dd = (double)w[q] - (double)zt;
        if (dd > depth) depth = dd;
        if (lim_depth && depth > max_depth) { ok = false; break; }
        if (w[q] > zt) path.push_back(q);
      }
      if (ok) {
        s_done++;
        for (int64_t qq : path)
          if (out[qq] > zt) out[qq] = zt;
      } else {
        s_left++;
      }
    }
    if (stats) {
      int64_t cut = 0;
      for (int64_t i = 0; i < n; i++) cut += out[i] < w[i];
      stats[0] = s_pits; stats[1] = s_done; stats[2] = s_left; stats[3] = cut;
    }
  } catch (const std::bad_alloc &) {
    hc_set_error("hc_breach_flowpath: out of host memory");
    return HC_ERR_NOMEM;
  }
  return HC_OK;
}


// seeds of the WhiteboxTools rule (valid cells on the raster frame or 8-adjacent to nodata connected to the frame)
extern "C" int hc_wbt_seed_mask_host(const float *w, uint8_t *seed_out, int64_t ny, int64_t nx, float nodata) {
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (!w || !seed_out) { hc_set_error("hc_wbt_seed_mask: null pointer"); return HC_ERR_ARG; }
  try {
    std::vector<uint8_t> seed;
    seed_mask(w, seed, ny, nx, nodata);
    memcpy(seed_out, seed.data(), (size_t)(ny * nx));
  } catch (const std::bad_alloc &) {
    hc_set_error("hc_wbt_seed_mask: out of host memory");
    return HC_ERR_NOMEM;
  }
  return HC_OK;
}

// ---- flow-path breach on the device (EXPERIMENTAL: written against oracle/breach_flowpath.py, NOT YET RUN ON A GPU;
//      nothing in the default paths calls it -- tests/test_zz_gpu_experimental.py is its opt-in check) -----------------
// seed = valid cell on the raster frame or 8-adjacent to exterior nodata (`ext`, see ext_nodata_mask in label.cu)
__global__ void __launch_bounds__(256)
k_bf_seed(const float *__restrict__ w, const uint8_t *__restrict__ ext, uint8_t *__restrict__ seed, int ny, int nx,
          float nodata) {
  int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c >= nx) return;
  size_t i = (size_t)r * nx + c;
  uint8_t s = 0;
  if (!d_is_nodata(w[i], nodata)) {
    if (r == 0 || c == 0 || r == ny - 1 || c == nx - 1) s = 1;
    else {
#pragma unroll
      for (int k = 0; k < 8; k++) s |= ext[(size_t)(r + c_dr[0][k]) * nx + c + c_dc[0][k]];
    }
  }
  seed[i] = s;
}

// every seed one float32 ulp lower: flats on a seed's level resolve toward the seed
__global__ void k_bf_lower_seeds(float *__restrict__ filled, const uint8_t *__restrict__ seed, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    if (seed[i]) filled[i] = nextafterf(filled[i], -INFINITY);
}

__device__ __forceinline__ void bf_atomic_min(float *addr, float val) {
  int *a = (int *)addr;
  int old = *a;
  while (__int_as_float(old) > val) {
    int assumed = old;
    old = atomicCAS(a, assumed, __float_as_int(val));
    if (old == assumed) break;
  }
}

// one thread per cell; pits walk the pointer forest twice (decide, then cut).  cnt: [0] pits [1] breached [2] left
// [3] error (a non-seed cell without pointer, or a cycle)
__global__ void __launch_bounds__(256)
k_bf_walk(const float *__restrict__ w, const uint8_t *__restrict__ dir, const uint8_t *__restrict__ seed,
          float *__restrict__ out, int ny, int nx, float nodata, long long max_length, double max_depth,
          unsigned long long *__restrict__ cnt) {
  const size_t n = (size_t)ny * nx;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
    const float zt = w[p];
    if (seed[p] || d_is_nodata(zt, nodata)) continue;
    const int r = (int)(p / (size_t)nx), c = (int)(p - (size_t)r * nx);
    bool pit = true;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      int rr = r + c_dr[0][k], cc = c + c_dc[0][k];
      if (rr < 0 || cc < 0 || rr >= ny || cc >= nx) continue;
      float zn = w[(size_t)rr * nx + cc];
      if (!d_is_nodata(zn, nodata) && zn < zt) pit = false;
    }
    if (!pit) continue;
    atomicAdd(&cnt[0], 1ull);
    bool ok = true;
    long long steps_ok = 0;
    for (int pass = 0; pass < 2 && ok; pass++) {
      size_t q = p;
      long long steps = 0;
      double depth = 0.0;
      for (;;) {
        if (seed[q]) break;
        int k = d_code_to_k(HC_TABLE_WBT, dir[q]);
        if (k < 0 || (size_t)steps > n) { atomicExch(&cnt[3], 1ull); ok = false; break; }
        q = (size_t)((long long)q + (long long)c_dr[0][k] * nx + c_dc[0][k]);
        steps++;
        if (pass == 1 && steps > steps_ok) break;
        if (max_length > 0 && steps > max_length) { ok = false; break; }
        float zq = w[q];
        if (zq < zt) break;
        double dd = (double)zq - (double)zt;
        if (dd > depth) depth = dd;
        if (max_depth > 0.0 && depth > max_depth) { ok = false; break; }
        if (pass == 1 && zq > zt) bf_atomic_min(&out[q], zt);
      }
      if (pass == 0) steps_ok = steps;
    }
    atomicAdd(&cnt[ok ? 1 : 2], 1ull);
  }
}

__global__ void k_bf_count_cut(const float *__restrict__ w, const float *__restrict__ out, size_t n,
                               unsigned long long *__restrict__ cnt) {
  unsigned long long local = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    local += out[i] < w[i];
  if (local) atomicAdd(&cnt[4], local);
}

// w -> out (cut surface, fill it afterwards with hc_fill_dev, seed mode WBT).  filled [n float], dir [n uint8] and
// seed [n uint8] are caller-provided device rasters that receive the intermediate results.  stats: 4 x int64 (host).
extern "C" int hc_breach_flowpath_dev(hc_ctx *ctx, const float *w, float *out, float *filled, uint8_t *dir,
                                      uint8_t *seed, int64_t ny, int64_t nx, float nodata, double dx, double dy,
                                      int64_t max_length, double max_depth, int64_t *stats, void *stream) {
  if (!ctx) { hc_set_error("null context"); return HC_ERR_ARG; }
  hc_guard hc_guard__(ctx);
  HC_CUDA(cudaSetDevice(ctx->device));
  if (stats) memset(stats, 0, sizeof(int64_t) * 4);
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (!w || !out || !filled || !dir || !seed || w == out || ny > 65535) {
    hc_set_error("hc_breach_flowpath_dev: bad argument");
    return HC_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)ny * nx;
  // seeds of the WBT rule
  HC_TRY(arena_reset(ctx));
  uint8_t *ext = (uint8_t *)arena_take(ctx, n);
  if (!ext) return HC_ERR_NOMEM;
  int has_interior = 0;
  HC_TRY(ext_nodata_mask(ctx, w, ext, ny, nx, nodata, &has_interior, st));
  dim3 g2((unsigned)((nx + 255) / 256), (unsigned)ny);
  HC_PROF(ctx, st, "k_bf_seed");
  k_bf_seed<<<g2, 256, 0, st>>>(w, ext, seed, (int)ny, (int)nx, nodata);
  HC_LAUNCH_CHECK(ctx);
  // pointers of the filled surface with the seeds one ulp lower
  HC_TRY(fill_run(ctx, w, filled, ny, nx, nodata, HC_SEED_WBT, st));
  HC_PROF(ctx, st, "k_bf_lower_seeds");
  k_bf_lower_seeds<<<ctx->sm_count * 8, 256, 0, st>>>(filled, seed, n);
  HC_LAUNCH_CHECK(ctx);
  HC_TRY(d8_flats_run(ctx, filled, dir, nullptr, ny, nx, nodata, dx, dy, HC_TABLE_WBT, st));
  // the walks
  HC_CUDA(cudaMemcpyAsync(out, w, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  HC_TRY(arena_reset(ctx));
  unsigned long long *cnt = (unsigned long long *)arena_take(ctx, 256);
  if (!cnt) return HC_ERR_NOMEM;
  HC_CUDA(cudaMemsetAsync(cnt, 0, 64, st));
  HC_PROF(ctx, st, "k_bf_walk");
  k_bf_walk<<<ctx->sm_count * 8, 256, 0, st>>>(w, dir, seed, out, (int)ny, (int)nx, nodata, (long long)max_length,
                                               std::isfinite(max_depth) ? max_depth : 0.0, cnt);
  HC_LAUNCH_CHECK(ctx);
  HC_PROF(ctx, st, "k_bf_count_cut");
  k_bf_count_cut<<<ctx->sm_count * 8, 256, 0, st>>>(w, out, n, cnt);
  HC_LAUNCH_CHECK(ctx);
  unsigned long long h[8];
  HC_CUDA(cudaMemcpyAsync(h, cnt, 64, cudaMemcpyDeviceToHost, st));
  HC_CUDA(cudaStreamSynchronize(st));
  if (h[3]) { hc_set_error("hc_breach_flowpath_dev: a non-seed cell of the filled surface has no pointer"); return HC_ERR_ARG; }
  if (stats) { stats[0] = (int64_t)h[0]; stats[1] = (int64_t)h[1]; stats[2] = (int64_t)h[2]; stats[3] = (int64_t)h[4]; }
  return HC_OK;
}


// =================================================================================================
// Priority-Flood + epsilon (fix_flats / flat_increment > 0: dep.fill_pits' second fill, tools/depressions.py:195-198,
// :1271-1286; config_files/config_local.ini:161 leaves it at 0.0).  WhiteboxTools raises every cell to at least
// (the level of the cell it was reached from) + eps, in float64.  Cells are reached in non-decreasing order of level, so
// the parent is the neighbour with the lowest level and the surface is the UNIQUE solution of
//       W(c) = max(z(c), min over 8-neighbours n of W(n) + eps),   W = z on the seeds
// (every dependency points at a strictly lower level, so there is no cycle to make the heap's tie order matter).  It is
// computed here as the limit of the monotone iteration W <- max(W, F(W)) from the eps = 0 fill, which lies below it: the
// same float64 additions, one per step of a cell's chain, as the serial flood makes, hence the same doubles.  Jacobi
// sweeps over the raster, as many as the longest raised chain is long (a lake's diameter): a tolerance-class option, not
// a hot path -- what matters is that the key gives a result instead of an exception.
// =================================================================================================
// TauDEM seed rule: frame cells and cells 8-adjacent to any nodata cell
__global__ void __launch_bounds__(256)
k_bf_seed_any(const float *__restrict__ w, uint8_t *__restrict__ seed, int ny, int nx, float nodata) {
  int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c >= nx) return;
  size_t i = (size_t)r * nx + c;
  uint8_t s = 0;
  if (!d_is_nodata(w[i], nodata)) {
    if (r == 0 || c == 0 || r == ny - 1 || c == nx - 1) s = 1;
    else {
#pragma unroll
      for (int k = 0; k < 8; k++) s |= d_is_nodata(w[(size_t)(r + c_dr[0][k]) * nx + c + c_dc[0][k]], nodata) ? 1 : 0;
    }
  }
  seed[i] = s;
}

__global__ void __launch_bounds__(256)
k_eps_init(const float *__restrict__ z, const float *__restrict__ filled, double *__restrict__ W, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    (void)z;
    W[i] = (double)filled[i];
  }
}

__global__ void __launch_bounds__(256)
k_eps_sweep(const float *__restrict__ z, const uint8_t *__restrict__ seed, const double *__restrict__ A, double *__restrict__ B,
            int ny, int nx, float nodata, double eps, u32 *__restrict__ changed) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  int ch = 0;
  if (c < nx) {
    const size_t i = (size_t)r * nx + c;
    double w = A[i];
    const float zc = z[i];
    if (!d_is_nodata(zc, nodata) && !seed[i]) {
      double m = INFINITY;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int rr = r + c_dr[0][k], cc = c + c_dc[0][k];
        if (rr < 0 || rr >= ny || cc < 0 || cc >= nx) continue;
        const size_t j = (size_t)rr * nx + cc;
        if (d_is_nodata(z[j], nodata)) continue;
        const double wn = A[j];
        m = wn < m ? wn : m;
      }
      if (m < INFINITY) {
        const double cand = fmax((double)zc, m + eps);
        if (cand > w) { w = cand; ch = 1; }
      }
    }
    B[i] = w;
  }
  ch = __syncthreads_or(ch);
  if (ch && threadIdx.x == 0) *changed = 1u;
}

extern "C" int hc_fill_eps_dev(hc_ctx *ctx, const float *z, double *out, int64_t ny, int64_t nx, float nodata,
                               int seed_mode, double eps, int64_t *sweeps, void *stream) {
  if (!ctx) { hc_set_error("null context"); return HC_ERR_ARG; }
  hc_guard hc_guard__(ctx);
  HC_CUDA(cudaSetDevice(ctx->device));
  if (sweeps) *sweeps = 0;
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (!z || !out || !(eps >= 0.0)) { hc_set_error("hc_fill_eps: bad argument"); return HC_ERR_ARG; }
  if (seed_mode != HC_SEED_WBT && seed_mode != HC_SEED_TAUDEM) { hc_set_error("hc_fill_eps: bad seed_mode"); return HC_ERR_ARG; }
  if (ny > 65535) { hc_set_error("hc_fill_eps: more than 65535 rows"); return HC_ERR_TOO_LARGE; }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)ny * nx;
  float *filled = nullptr;
  double *B = nullptr;
  uint8_t *seed = nullptr, *ext = nullptr;
  u32 *flag = nullptr;
  HC_CUDA(cudaMalloc(&filled, n * 4));
  cudaError_t e = cudaMalloc(&B, n * 8);
  if (e == cudaSuccess) e = cudaMalloc(&seed, n);
  if (e == cudaSuccess) e = cudaMalloc(&ext, n);
  if (e == cudaSuccess) e = cudaMalloc(&flag, 256);
  int rc = HC_OK;
  auto cleanup = [&]() { cudaFree(filled); cudaFree(B); cudaFree(seed); cudaFree(ext); cudaFree(flag); };
  if (e != cudaSuccess) { cudaGetLastError(); cleanup(); hc_set_error("hc_fill_eps: out of device memory"); return HC_ERR_NOMEM; }
  // the eps = 0 surface (a lower bound), and the seed mask of the chosen rule
  rc = fill_run(ctx, z, filled, ny, nx, nodata, seed_mode, st);
  if (rc == HC_OK && seed_mode == HC_SEED_WBT) {
    // WBT rule: only EXTERIOR nodata seeds its neighbours.  ext_nodata_mask leaves the plain nodata mask in `ext` when
    // no nodata cell lies off the frame (then all of it is exterior) and the exterior subset otherwise.
    rc = arena_reset(ctx);
    int has_interior = 0;
    if (rc == HC_OK) rc = ext_nodata_mask(ctx, z, ext, ny, nx, nodata, &has_interior, st);
  }
  if (rc != HC_OK) { cleanup(); return rc; }
  dim3 g2((unsigned)((nx + 255) / 256), (unsigned)ny);
  if (seed_mode == HC_SEED_TAUDEM) {
    // seeds = frame cells + cells 8-adjacent to ANY nodata: a one-pass predicate on z itself
    k_bf_seed_any<<<g2, 256, 0, st>>>(z, seed, (int)ny, (int)nx, nodata);
  } else {
    k_bf_seed<<<g2, 256, 0, st>>>(z, ext, seed, (int)ny, (int)nx, nodata);
  }
  k_eps_init<<<ctx->sm_count * 8, 256, 0, st>>>(z, filled, out, n);
  ctx->launches += 2;
  double *A = out;
  int64_t nsweeps = 0;
  if (eps > 0.0) {
    for (;;) {
      // 8 sweeps per read-back; the ping-pong ends in `out` (an even number of sweeps per batch)
      cudaMemsetAsync(flag, 0, 4, st);
      for (int k = 0; k < 8; k++) {
        k_eps_sweep<<<g2, 256, 0, st>>>(z, seed, A, B, (int)ny, (int)nx, nodata, eps, flag);
        double *t = A; A = B; B = t;
        ctx->launches++;
      }
      nsweeps += 8;
      u32 h = 0;
      if (cudaMemcpyAsync(&h, flag, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
        cudaError_t e2 = cudaGetLastError();
        cleanup();
        hc_set_error("hc_fill_eps: %s", cudaGetErrorString(e2));
        return HC_ERR_CUDA;
      }
      if (!h) break;
      if (nsweeps > 4 * (ny + nx) + 64) { cleanup(); hc_set_error("hc_fill_eps: no fixed point after %lld sweeps", (long long)nsweeps); return HC_ERR_INTERNAL; }
    }
  }
  // (8 sweeps per batch: A is `out` again)
  cudaError_t e3 = cudaStreamSynchronize(st);
  cleanup();
  if (e3 != cudaSuccess) { hc_set_error("hc_fill_eps: %s", cudaGetErrorString(e3)); return HC_ERR_CUDA; }
  if (sweeps) *sweeps = nsweeps;
  return HC_OK;
}
