// leastcost.cu -- least-cost breaching on the device, the arithmetic behind
//   wbt.breach_depressions_least_cost(dem, output, dist, max_cost=, min_dist=, flat_increment=, fill=False)
// as dep.breach_pits(method='LC') and dep.combined_solution(carve_method='LC') call it
// (tools/depressions.py:1370-1374, :996-1000; config key carve_method, config_files/config_local.ini:154).
//
// WhiteboxTools is not in /root/reference (un-vendored, version unpinned): this follows the published method (Lindsay &
// Dhun 2015; the tool's documentation) in an ORDER-INDEPENDENT form -- PARITY UNPINNED.  The tool breaches its pits one
// after the other; here every pit is searched on the input surface and the cuts are combined with min(), so one CTA
// per pit can run all searches at once.  The rules (valid / edge / pit / cut / target / label / path / channel) are
// written out in oracle/breach_leastcost.py, the CPU specification this file is compared with bit for bit.
//
// One search: the (2 dist + 1)^2 window lives in shared memory (label u64, cut u32, target flag u8 per cell) and is
// loaded ring by ring as the search grows; labels are relaxed in place (cut depths are integers in 1/1024 units, so a
// path's cost is exact and the least label is the unique fixed point whatever the order of the relaxation), candidates
// dearer than the best target found so far are dropped, and a sweep that changes nothing with nothing left on the outer
// ring ends the search.  A pit whose neighbour is already lower needs one ring and one sweep.
#include <math.h>
#include <string.h>

#include "common.cuh"

#define LC_THREADS 256
#define LC_INF 0xFFFFFFFFFFFFFFFFull
#define LC_NOCELL 0xFFFFFFFFu
#define LC_CUT_CAP (1u << 30)
#define LC_COST_CAP (1ull << 40)
#define LC_VALID 1
#define LC_EDGE 2

// flags: bit 0 valid, bit 1 edge (valid and on the raster frame or 8-adjacent to a cell that is not valid)
__global__ void __launch_bounds__(256)
k_lc_flags(const float *__restrict__ w, uint8_t *__restrict__ flags, int ny, int nx, float nodata) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c >= nx) return;
  const size_t g = (size_t)r * nx + c;
  uint8_t f = 0;
  if (!d_is_nodata(w[g], nodata)) {
    f = LC_VALID;
    bool edge = r == 0 || c == 0 || r == ny - 1 || c == nx - 1;
    if (!edge) {
#pragma unroll
      for (int dr = -1; dr <= 1; dr++)
#pragma unroll
        for (int dc = -1; dc <= 1; dc++)
          if (d_is_nodata(w[g + (long long)dr * nx + dc], nodata)) edge = true;
    }
    if (edge) f |= LC_EDGE;
  }
  flags[g] = f;
}

// pit: valid, not edge, no neighbour lower, no neighbour of the same elevation with a smaller raster index
__global__ void __launch_bounds__(256)
k_lc_pits(const float *__restrict__ w, const uint8_t *__restrict__ flags, int ny, int nx, u32 *__restrict__ pits,
          unsigned long long *__restrict__ cnt, u32 cap) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c >= nx) return;
  const size_t g = (size_t)r * nx + c;
  if (flags[g] != LC_VALID) return;
  const float v = w[g];
  int k = 0;
#pragma unroll
  for (int dr = -1; dr <= 1; dr++)
#pragma unroll
    for (int dc = -1; dc <= 1; dc++) {
      if (!dr && !dc) continue;
      const float zn = w[g + (long long)dr * nx + dc];
      if (zn < v || (zn == v && k < 4)) return;
      k++;
    }
  const unsigned long long pos = atomicAdd(&cnt[0], 1ull);
  if (pos < cap) pits[pos] = (u32)g;
}

__device__ __forceinline__ void lc_atomic_min_float(float *addr, float v) {
  // (ordered as signed ints when >= 0, reversed as unsigned ints when negative)
  if (v >= 0.0f) atomicMin((int *)addr, __float_as_int(v));
  else atomicMax((unsigned int *)addr, __float_as_uint(v));
}

template <bool MIN_DIST>
__device__ __forceinline__ u64 lc_pack(u64 cost, u32 steps) {
  return MIN_DIST ? (((u64)steps << 44) | cost) : ((cost << 20) | (u64)steps);
}
template <bool MIN_DIST>
__device__ __forceinline__ void lc_unpack(u64 key, u64 &cost, u32 &steps) {
  if (MIN_DIST) { steps = (u32)(key >> 44); cost = key & ((1ull << 44) - 1); }
  else { steps = (u32)(key & 0xFFFFFu); cost = key >> 20; }
}

// cnt: [0] pits [1] breached [2] left [3] next pit to fetch [4] cells lowered
template <bool MIN_DIST>
__global__ void __launch_bounds__(LC_THREADS)
k_lc_search(const float *__restrict__ w, const uint8_t *__restrict__ flags, float *out, int ny, int nx,
            const u32 *__restrict__ pits, unsigned long long *cnt, u32 cap, int dist, u64 maxcost_u, double eps) {
  extern __shared__ __align__(16) unsigned char smem_lc[];
  const int W = 2 * dist + 1, W2 = W * W;
  u64 *key = (u64 *)smem_lc;
  u32 *cutv = (u32 *)(smem_lc + (size_t)W2 * 8);
  uint8_t *tflag = smem_lc + (size_t)W2 * 12;
  __shared__ unsigned long long s_best, s_pit;
  __shared__ u32 s_tidx;
  const int tid = threadIdx.x;
  const unsigned long long npits = cnt[0] < cap ? cnt[0] : cap;
  for (;;) {
    __syncthreads();
    if (tid == 0) { s_pit = atomicAdd(&cnt[3], 1ull); s_best = LC_INF; s_tidx = 0xFFFFFFFFu; }
    __syncthreads();
    if (s_pit >= npits) break;
    const u32 gp = pits[s_pit];
    const int pr = (int)(gp / (u32)nx), pc = (int)(gp % (u32)nx);
    const float zp = w[gp];
    const int centre = dist * W + dist;
    if (tid == 0) { key[centre] = 0ull; cutv[centre] = 0u; tflag[centre] = 0; }
    int r = 0, sweeps = 0;
    for (;;) {
      if (r < dist) {
        // the next ring of the window
        r++;
        for (int i = tid; i < 8 * r; i += LC_THREADS) {
          const int side = i / (2 * r), j = i - side * 2 * r;
          int dr, dc;
          if (side == 0) { dr = -r; dc = -r + j; }
          else if (side == 1) { dr = -r + j; dc = r; }
          else if (side == 2) { dr = r; dc = r - j; }
          else { dr = r - j; dc = -r; }
          const int rr = pr + dr, cc = pc + dc;
          const int wi = (dr + dist) * W + (dc + dist);
          u32 cu = LC_NOCELL;
          uint8_t tf = 0;
          if (rr >= 0 && rr < ny && cc >= 0 && cc < nx) {
            const size_t g = (size_t)rr * nx + cc;
            const uint8_t f = flags[g];
            if (f & LC_VALID) {
              const float z = w[g];
              const float d = __fsub_rn(z, zp);
              cu = 0u;
              if (d > 0.0f) {
                const float u = floorf(__fmul_rn(d, 1024.0f));
                cu = u >= (float)LC_CUT_CAP ? LC_CUT_CAP : (u32)u;
              }
              tf = (z < zp || (f & LC_EDGE)) ? 1 : 0;
            }
          }
          key[wi] = LC_INF; cutv[wi] = cu; tflag[wi] = tf;
        }
      }
      __syncthreads();
      // one in-place sweep over the loaded box
      const int side = 2 * r + 1;
      int flagsw = 0;
      const u64 best_now = *(volatile unsigned long long *)&s_best;
      for (int i = tid; i < side * side; i += LC_THREADS) {
        const int dr = i / side - r, dc = i - (i / side) * side - r;
        if (!dr && !dc) continue;
        const int wi = (dr + dist) * W + (dc + dist);
        const u32 cu = cutv[wi];
        if (cu == LC_NOCELL) continue;
        u64 cand = LC_INF;
#pragma unroll
        for (int a = -1; a <= 1; a++)
#pragma unroll
          for (int b = -1; b <= 1; b++) {
            if (!a && !b) continue;
            const int ndr = dr + a, ndc = dc + b;
            if (ndr < -r || ndr > r || ndc < -r || ndc > r) continue;
            const int ni = wi + a * W + b;
            if (tflag[ni]) continue;                       // a target is an end point, never an inner cell
            const u64 kn = *(volatile u64 *)&key[ni];
            if (kn == LC_INF) continue;
            u64 cost; u32 steps;
            lc_unpack<MIN_DIST>(kn, cost, steps);
            cost += cu;
            if (cost > maxcost_u) continue;
            const u64 k2 = lc_pack<MIN_DIST>(cost, steps + 1u);
            if (k2 < cand) cand = k2;
          }
        const u64 cur = key[wi];
        if (cand < cur && cand <= best_now) {
          key[wi] = cand;
          flagsw |= 1;
          if (tflag[wi]) atomicMin(&s_best, (unsigned long long)cand);
        }
        const u64 now = cand < cur ? cand : cur;
        if (now != LC_INF && now <= best_now && !tflag[wi] && (dr == -r || dr == r || dc == -r || dc == r)) flagsw |= 2;
      }
      // (__syncthreads_or answers "any thread non-zero", not the OR of the values: one reduction per flag)
      const int changed = __syncthreads_or(flagsw & 1);
      const int outer = __syncthreads_or(flagsw & 2);
      if (!changed && (r == dist || !outer)) break;
      if (++sweeps > 8 * W2 + 64) {   // cannot happen (labels only ever decrease); never spin on a broken state
        if (tid == 0) atomicAdd(&cnt[5], 1ull);
        break;
      }
    }
    const u64 best = s_best;
    if (best == LC_INF) {
      if (tid == 0) atomicAdd(&cnt[2], 1ull);
      continue;
    }
    // the target: least label, then smallest window index
    {
      const int side = 2 * r + 1;
      for (int i = tid; i < side * side; i += LC_THREADS) {
        const int dr = i / side - r, dc = i - (i / side) * side - r;
        const int wi = (dr + dist) * W + (dc + dist);
        if (tflag[wi] && cutv[wi] != LC_NOCELL && key[wi] == best) atomicMin(&s_tidx, (u32)wi);
      }
    }
    __syncthreads();
    if (tid == 0) {
      atomicAdd(&cnt[1], 1ull);
      int wi = (int)s_tidx;
      u64 cost; u32 k;
      lc_unpack<MIN_DIST>(best, cost, k);
      const int tr = wi / W - dist, tc = wi % W - dist;
      const float zt = w[(size_t)(pr + tr) * nx + (pc + tc)];
      const bool lower = zt < zp;
      double delta = eps;
      if (lower) { const double lim = __ddiv_rn(__dsub_rn((double)zp, (double)zt), (double)k); if (lim < delta) delta = lim; }
      bool first = true;
      while (wi != centre) {
        const int dr = wi / W - dist, dc = wi % W - dist;
        u64 ccost; u32 csteps;
        lc_unpack<MIN_DIST>(key[wi], ccost, csteps);
        if (!first || !lower) {
          const float ch = __double2float_rn(__dsub_rn((double)zp, __dmul_rn((double)csteps, delta)));
          lc_atomic_min_float(out + (size_t)(pr + dr) * nx + (pc + dc), ch);
        }
        first = false;
        int pred = -1;
        for (int a = -1; a <= 1 && pred < 0; a++)
          for (int b = -1; b <= 1; b++) {
            if (!a && !b) continue;
            const int ndr = dr + a, ndc = dc + b;
            if (ndr < -r || ndr > r || ndc < -r || ndc > r) continue;
            const int ni = wi + a * W + b;
            if (cutv[ni] == LC_NOCELL || tflag[ni] || key[ni] == LC_INF) continue;
            u64 ncost; u32 nsteps;
            lc_unpack<MIN_DIST>(key[ni], ncost, nsteps);
            if (ncost + cutv[wi] == ccost && nsteps + 1u == csteps) { pred = ni; break; }
          }
        if (pred < 0) { atomicAdd(&cnt[5], 1ull); break; }   // cannot happen: every finite label has a predecessor
        wi = pred;
      }
    }
  }
}

__global__ void k_lc_count_cut(const float *__restrict__ w, const float *__restrict__ out, size_t n,
                               unsigned long long *__restrict__ cnt) {
  unsigned long long local = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    local += out[i] < w[i];
  if (local) atomicAdd(&cnt[4], local);
}

// w -> out: the cut surface (run hc_fill_dev on it afterwards for the tool's fill=True / pydrodem's closing WL fill).
// dist: search radius in cells (1 .. 65: the window has to fit in shared memory); max_cost <= 0 or not finite: no limit;
// flat_increment <= 0: level channels.  stats (4 x int64, nullable): pits, breached, left, cells lowered.
extern "C" int hc_breach_least_cost_dev(hc_ctx *ctx, const float *w, float *out, int64_t ny, int64_t nx, float nodata,
                                        int64_t dist, double max_cost, int min_dist, double flat_increment,
                                        int64_t *stats, void *stream) {
  HC_ENTER(ctx);
  if (stats) memset(stats, 0, sizeof(int64_t) * 4);
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (!w || !out || w == out || (uint64_t)ny * (uint64_t)nx > 0xFFFFFFF0ull) { hc_set_error("hc_breach_least_cost_dev: bad argument"); return HC_ERR_ARG; }
  if (dist < 1 || dist > 65) { hc_set_error("hc_breach_least_cost_dev: dist must be 1..65 cells (got %lld)", (long long)dist); return HC_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)ny * nx;
  HC_TRY(arena_reset(ctx));
  const u32 cap = (u32)(n / 4 + (size_t)nx + (size_t)ny + 16);
  uint8_t *flags = (uint8_t *)arena_take(ctx, n);
  u32 *pits = (u32 *)arena_take(ctx, (size_t)cap * 4);
  unsigned long long *cnt = (unsigned long long *)arena_take(ctx, 256);
  if (!flags || !pits || !cnt) return HC_ERR_NOMEM;
  HC_CUDA(cudaMemsetAsync(cnt, 0, 64, st));
  HC_CUDA(cudaMemcpyAsync(out, w, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  dim3 g2((unsigned)((nx + 255) / 256), (unsigned)ny);
  HC_PROF(ctx, st, "k_lc_flags");
  k_lc_flags<<<g2, 256, 0, st>>>(w, flags, (int)ny, (int)nx, nodata);
  HC_LAUNCH_CHECK(ctx);
  HC_PROF(ctx, st, "k_lc_pits");
  k_lc_pits<<<g2, 256, 0, st>>>(w, flags, (int)ny, (int)nx, pits, cnt, cap);
  HC_LAUNCH_CHECK(ctx);
  u64 maxcost_u = LC_COST_CAP;
  if (std::isfinite(max_cost) && max_cost > 0.0) {
    const double u = floor((double)(float)max_cost * 1024.0);
    if (u < (double)LC_COST_CAP) maxcost_u = (u64)u;
  }
  const double eps = (std::isfinite(flat_increment) && flat_increment > 0.0) ? flat_increment : 0.0;
  const int W = 2 * (int)dist + 1;
  const size_t smem = (size_t)W * W * 13 + 16;
  int per_sm = 0;
  if (min_dist) {
    HC_CUDA(cudaFuncSetAttribute(k_lc_search<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_lc_search<true>, LC_THREADS, smem));
  } else {
    HC_CUDA(cudaFuncSetAttribute(k_lc_search<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_lc_search<false>, LC_THREADS, smem));
  }
  if (per_sm < 1) { hc_set_error("hc_breach_least_cost_dev: window does not fit in shared memory"); return HC_ERR_ARG; }
  const unsigned grid = (unsigned)(per_sm * ctx->sm_count);
  HC_PROF(ctx, st, "k_lc_search");
  if (min_dist) k_lc_search<true><<<grid, LC_THREADS, smem, st>>>(w, flags, out, (int)ny, (int)nx, pits, cnt, cap, (int)dist, maxcost_u, eps);
  else k_lc_search<false><<<grid, LC_THREADS, smem, st>>>(w, flags, out, (int)ny, (int)nx, pits, cnt, cap, (int)dist, maxcost_u, eps);
  HC_LAUNCH_CHECK(ctx);
  HC_PROF(ctx, st, "k_lc_count_cut");
  k_lc_count_cut<<<ctx->sm_count * 8, 256, 0, st>>>(w, out, n, cnt);
  HC_LAUNCH_CHECK(ctx);
  unsigned long long h[8];
  HC_CUDA(cudaMemcpyAsync(h, cnt, 64, cudaMemcpyDeviceToHost, st));
  HC_CUDA(cudaStreamSynchronize(st));
  if (h[0] > cap) { hc_set_error("hc_breach_least_cost_dev: pit list overflow"); return HC_ERR_ARG; }
  if (h[5]) { hc_set_error("hc_breach_least_cost_dev: broken label chain"); return HC_ERR_ARG; }
  if (stats) { stats[0] = (int64_t)h[0]; stats[1] = (int64_t)h[1]; stats[2] = (int64_t)h[2]; stats[3] = (int64_t)h[4]; }
  return HC_OK;
}
