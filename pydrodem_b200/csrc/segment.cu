// segment.cu — the pit catalog / solution selection path of pydrodem as segmented reductions.
//
// Replaces the nested Python loops of tools/depressions.py: build_pit_catalog (:272-509),
// select_solution (:517-810), apply_solution (:1140-1228), fill_shallow_pits (:1471-1531) and
// raise_lake_connected_pits (:1536-1590).  A "pit" is a label of the fill-difference map, a
// "carve" a label of the negative carve-difference map; the reference walks per-pit index lists
// (np.flatnonzero / np.intersect1d / np.isin per pit), here every quantity is a reduction keyed by
// the label rasters:
//   per pit    max / sum / count of diff_fill, fill-behind sum / max / count
//   per carve  first argmax / argmin of the carved elevations, min / sum / count of diff_carve
//   pairs      the distinct (pit, carve) pairs with a common cell (hash-deduplicated), each with the
//              reference's inclusion test "carve's highest cell inside the pit OR lowest outside"
//   per pit    min / sum / count over its included carves
// Sums are accumulated in float64 (the reference sums float32 pairwise): tolerance class, 1e-5 rel.
#include "common.cuh"

#define SEG_CHUNK 8  // consecutive cells per thread: labels are spatially coherent, so a thread
                     // pre-reduces runs of equal ids and issues one atomic per run

__device__ __forceinline__ void atomic_max_f32(float *addr, float v) {
  // ordered-uint trick on the raw bits (values are never NaN here)
  if (v >= 0.f) atomicMax((int *)addr, __float_as_int(v));
  else atomicMin((unsigned int *)addr, __float_as_uint(v));
}
__device__ __forceinline__ void atomic_min_f32(float *addr, float v) {
  if (v >= 0.f) atomicMin((int *)addr, __float_as_int(v));
  else atomicMax((unsigned int *)addr, __float_as_uint(v));
}

__global__ void k_fill_f32(float *p, float v, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// per-segment max, min, sum (f64), count of vals over cells whose id is in [lo_id, nseg] and whose
// optional selector sel[i] != 0
__global__ void __launch_bounds__(256)
k_seg_stats(const int32_t *__restrict__ ids, const float *__restrict__ vals, const int32_t *__restrict__ sel,
            size_t n, int nseg, int lo_id, float *vmax, float *vmin, double *vsum, long long *cnt) {
  size_t base = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * SEG_CHUNK;
  if (base >= n) return;
  int cur = -1;
  float mx = -INFINITY, mn = INFINITY;
  double sm = 0.0;
  long long c = 0;
  auto flush = [&]() {
    if (cur >= lo_id && cur <= nseg && c) {
      if (vmax) atomic_max_f32(&vmax[cur], mx);
      if (vmin) atomic_min_f32(&vmin[cur], mn);
      if (vsum) atomicAdd(&vsum[cur], sm);
      if (cnt) atomicAdd((unsigned long long *)&cnt[cur], (unsigned long long)c);
    }
  };
  size_t end = base + SEG_CHUNK < n ? base + SEG_CHUNK : n;
  for (size_t i = base; i < end; i++) {
    int id = ids[i];
    if (sel && !sel[i]) id = -1;
    if (id != cur) { flush(); cur = id; mx = -INFINITY; mn = INFINITY; sm = 0.0; c = 0; }
    if (id >= lo_id) {
      float v = vals ? vals[i] : 0.f;
      mx = fmaxf(mx, v); mn = fminf(mn, v); sm += (double)v; c++;
    }
  }
  flush();
}

// first index of the maximum / minimum per segment: key = (orderable(v), index) packed in 64 bits
__global__ void __launch_bounds__(256)
k_seg_argext(const int32_t *__restrict__ ids, const float *__restrict__ vals, size_t n, int nseg,
             u64 *kmax, u64 *kmin) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int id = ids[i];
  if (id < 1 || id > nseg) return;
  u32 o = d_ford(vals[i]);
  // max value, smallest index on ties  <=> max of (o, ~index)
  u64 k1 = ((u64)o << 32) | (u64)(0xFFFFFFFFu - (u32)i);
  if (k1 > kmax[id]) atomicMax(&kmax[id], k1);
  u64 k2 = ((u64)o << 32) | (u64)(u32)i;
  if (k2 < kmin[id]) atomicMin(&kmin[id], k2);
}
__global__ void k_argext_decode(const u64 *kmax, const u64 *kmin, int nseg, long long *amax, long long *amin) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > nseg) return;
  amax[s] = kmax[s] == 0ull ? -1ll : (long long)(0xFFFFFFFFu - (u32)(kmax[s] & 0xFFFFFFFFu));
  amin[s] = kmin[s] == ~0ull ? -1ll : (long long)(u32)(kmin[s] & 0xFFFFFFFFu);
}

// carveID_bound (:330): the carve id where the DEM is valid, else 0
__global__ void k_carve_bound(const int32_t *__restrict__ carveID, const float *__restrict__ dem, float nodata, size_t n,
                              int32_t *__restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = carveID[i];
  out[i] = (c > 0 && dem[i] != nodata) ? c : 0;
}

// distinct (pit, carve) pairs: a cell proposes its pair when it differs from the previous cell's,
// an open-addressing table (64-bit CAS) removes the remaining duplicates
__global__ void __launch_bounds__(256)
k_pairs_insert(const int32_t *__restrict__ a, const int32_t *__restrict__ b, size_t n, u64 *table, u32 tmask,
               u32 *overflow) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int p = a[i], c = b[i];
  if (p <= 0 || c <= 0) return;
  if (i > 0 && a[i - 1] == p && b[i - 1] == c) return;
  u64 key = ((u64)(u32)p << 32) | (u64)(u32)c;
  u32 h = (u32)((key * 0x9E3779B97F4A7C15ull) >> 32) & tmask;
  for (u32 probe = 0; probe <= tmask; probe++) {
    u64 old = atomicCAS(&table[h], 0ull, key);
    if (old == 0ull || old == key) return;
    h = (h + 1) & tmask;
  }
  *overflow = 1u;
}
__global__ void k_pairs_compact(const u64 *table, u32 tsize, long long *out, long long cap, u32 *count) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= tsize) return;
  u64 k = table[i];
  if (!k) return;
  u32 o = atomicAdd(count, 1u);
  if ((long long)o < cap) out[o] = (long long)k;
}

// pair-level logic of build_pit_catalog (:412-441)
__global__ void k_pair_flags(const long long *pairs, u32 np, const int32_t *__restrict__ fillID,
                             const long long *carve_amax, const long long *carve_amin, int carveout, int carveout_id,
                             uint8_t *included, uint8_t *pit_out) {
  u32 j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= np) return;
  int p = (int)((u64)pairs[j] >> 32), c = (int)((u64)pairs[j] & 0xFFFFFFFFu);
  // "does carve leave basin?": the carve that contains the nodata pixels touches this pit
  if (!carveout && c == carveout_id) pit_out[p] = 1;
  long long imax = carve_amax[c], imin = carve_amin[c];
  bool inc = false;
  if (imax >= 0) inc = (fillID[imax] == p) || (fillID[imin] != p);
  included[j] = inc ? 1 : 0;
}
// per pit: aggregate its included carves
__global__ void k_pair_reduce(const long long *pairs, u32 np, const uint8_t *included, const uint8_t *pit_out,
                              const float *carve_min, const double *carve_sum, const long long *carve_n,
                              const long long *carve_water, float *pit_mincarve, double *pit_sumcarve,
                              long long *pit_ncarve, long long *pit_water) {
  u32 j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= np || !included[j]) return;
  int p = (int)((u64)pairs[j] >> 32), c = (int)((u64)pairs[j] & 0xFFFFFFFFu);
  if (pit_out && pit_out[p]) return;
  if (pit_mincarve) atomic_min_f32(&pit_mincarve[p], carve_min[c]);
  if (pit_sumcarve) atomicAdd(&pit_sumcarve[p], carve_sum[c]);
  if (pit_ncarve) atomicAdd((unsigned long long *)&pit_ncarve[p], (unsigned long long)carve_n[c]);
  if (pit_water) atomicAdd((unsigned long long *)&pit_water[p], (unsigned long long)carve_water[c]);
}
__global__ void k_catalog_finish(int npits, const uint8_t *pit_out, const float *pit_mincarve, const double *pit_sumcarve,
                                 const long long *pit_ncarve, const double *fb_sum, const float *fb_max,
                                 const long long *fb_n, float *maxcarve, float *volcarve, float *volfb, float *maxfb,
                                 int8_t *complete, const double *volfill64, float *volfill) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > npits) return;
  volfill[p] = (float)volfill64[p];
  bool has = !pit_out[p] && pit_ncarve[p] > 0;
  if (has) {
    maxcarve[p] = fabsf(pit_mincarve[p]);
    volcarve[p] = fabsf((float)pit_sumcarve[p]);
    volfb[p] = fabsf((float)fb_sum[p]);
    maxfb[p] = fb_n[p] > 0 ? fb_max[p] : 0.0f;
    complete[p] = 1;
  } else {
    maxcarve[p] = 0.f; volcarve[p] = 0.f; volfb[p] = 0.f; maxfb[p] = 0.f; complete[p] = 0;
  }
}

static inline unsigned nblk(size_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

// ------------------------------------------------------------------------------------------------
extern "C" int hc_seg_stats_dev(hc_ctx *ctx, const int32_t *ids, const float *vals, const int32_t *sel, int64_t n,
                                int32_t nseg, int32_t lo_id, float *vmax, float *vmin, double *vsum, int64_t *count,
                                void *stream) {
  if (!ctx) { hc_set_error("null context"); return HC_ERR_ARG; }
  hc_guard hc_guard__(ctx);
  HC_CUDA(cudaSetDevice(ctx->device));
  if (n < 0 || nseg < 0 || !ids) { hc_set_error("hc_seg_stats: bad argument"); return HC_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  size_t m = (size_t)nseg + 1;
  if (vmax) { k_fill_f32<<<nblk(m), 256, 0, st>>>(vmax, -INFINITY, m); HC_LAUNCH_CHECK(ctx); }
  if (vmin) { k_fill_f32<<<nblk(m), 256, 0, st>>>(vmin, INFINITY, m); HC_LAUNCH_CHECK(ctx); }
  if (vsum) HC_CUDA(cudaMemsetAsync(vsum, 0, m * 8, st));
  if (count) HC_CUDA(cudaMemsetAsync(count, 0, m * 8, st));
  if (n) {
    HC_PROF(ctx, st, "k_seg_stats");
    k_seg_stats<<<nblk(((size_t)n + SEG_CHUNK - 1) / SEG_CHUNK), 256, 0, st>>>(ids, vals, sel, (size_t)n, nseg, lo_id, vmax, vmin, vsum, (long long *)count);
    HC_LAUNCH_CHECK(ctx);
  }
  return HC_OK;
}

// build_pit_catalog (tools/depressions.py:272-509).  All pointers are device pointers; per-pit
// outputs have npits+1 entries (entry 0 unused), per-carve outputs ncarves+1.  pairs_out receives the
// distinct (pit << 32 | carve) pairs, pair_included the reference's inclusion test per pair;
// *npairs (host) their number.  carveout_id = carveID at the first nodata cell (:335), ignored when
// carveout != 0.
extern "C" int hc_pit_catalog_dev(hc_ctx *ctx, const float *dem, float nodata, const float *carve, const float *diff_carve,
                                  const float *diff_fill, const int32_t *fillID, const int32_t *carveID,
                                  const int32_t *carvefillID, int64_t n, int32_t npits, int32_t ncarves, int carveout,
                                  int32_t carveout_id, int32_t *carve_bound, float *maxfill, float *volfill,
                                  int64_t *nfill, float *maxcarve, float *volcarve, int64_t *ncarvecells, float *volfb,
                                  float *maxfb, int8_t *complete, uint8_t *pit_out, int64_t *carve_amax,
                                  int64_t *carve_amin, float *carve_min, double *carve_sum, int64_t *carve_n,
                                  int64_t *pairs_out, int64_t pair_cap, uint8_t *pair_included, int64_t *npairs,
                                  void *stream) {
  if (!ctx) { hc_set_error("null context"); return HC_ERR_ARG; }
  hc_guard hc_guard__(ctx);
  HC_CUDA(cudaSetDevice(ctx->device));
  if (n <= 0 || n >= 0x7FFFFFF0ll || npits < 0 || ncarves < 0) { hc_set_error("hc_pit_catalog: bad shape"); return HC_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  HC_TRY(arena_reset(ctx));
  const size_t P1 = (size_t)npits + 1, C1 = (size_t)ncarves + 1;
  double *volfill64 = (double *)arena_take(ctx, P1 * 8);
  double *fb_sum = (double *)arena_take(ctx, P1 * 8);
  float *fb_max = (float *)arena_take(ctx, P1 * 4);
  long long *fb_n = (long long *)arena_take(ctx, P1 * 8);
  float *pit_mincarve = (float *)arena_take(ctx, P1 * 4);
  double *pit_sumcarve = (double *)arena_take(ctx, P1 * 8);
  u64 *kmax = (u64 *)arena_take(ctx, C1 * 8);
  u64 *kmin = (u64 *)arena_take(ctx, C1 * 8);
  u32 *cnt = (u32 *)arena_take(ctx, 256);
  if (!volfill64 || !fb_sum || !fb_max || !fb_n || !pit_mincarve || !pit_sumcarve || !kmax || !kmin || !cnt) return HC_ERR_NOMEM;

  HC_PROF(ctx, st, "k_carve_bound");
  k_carve_bound<<<nblk((size_t)n), 256, 0, st>>>(carveID, dem, nodata, (size_t)n, carve_bound);
  HC_LAUNCH_CHECK(ctx);
  // per pit: max / sum / count of diff_fill (:392-394)
  HC_TRY(hc_seg_stats_dev(ctx, fillID, diff_fill, nullptr, n, npits, 1, maxfill, nullptr, volfill64, nfill, stream));
  // per pit: fill behind the carve = pit cells with carvefillID > 0 (:395,449,457)
  HC_TRY(hc_seg_stats_dev(ctx, fillID, diff_carve, carvefillID, n, npits, 1, fb_max, nullptr, fb_sum, (int64_t *)fb_n, stream));
  // per carve: min / sum / count of diff_carve over its (bounded) cells, first argmax / argmin of carve_vec (:421-424)
  HC_TRY(hc_seg_stats_dev(ctx, carve_bound, diff_carve, nullptr, n, ncarves, 1, nullptr, carve_min, carve_sum, carve_n, stream));
  HC_CUDA(cudaMemsetAsync(kmax, 0, C1 * 8, st));
  HC_CUDA(cudaMemsetAsync(kmin, 0xFF, C1 * 8, st));
  HC_PROF(ctx, st, "k_seg_argext");
  k_seg_argext<<<nblk((size_t)n), 256, 0, st>>>(carve_bound, carve, (size_t)n, ncarves, kmax, kmin);
  HC_LAUNCH_CHECK(ctx);
  k_argext_decode<<<nblk(C1), 256, 0, st>>>(kmax, kmin, ncarves, (long long *)carve_amax, (long long *)carve_amin);
  HC_LAUNCH_CHECK(ctx);
  // distinct (pit, carve) pairs
  u32 tsize = 1024;
  while ((long long)tsize < 2 * pair_cap) tsize <<= 1;
  u64 *table = (u64 *)arena_take(ctx, (size_t)tsize * 8);
  if (!table) return HC_ERR_NOMEM;
  HC_CUDA(cudaMemsetAsync(table, 0, (size_t)tsize * 8, st));
  HC_CUDA(cudaMemsetAsync(cnt, 0, 16, st));
  HC_PROF(ctx, st, "k_pairs_insert");
  k_pairs_insert<<<nblk((size_t)n), 256, 0, st>>>(fillID, carve_bound, (size_t)n, table, tsize - 1, cnt + 1);
  HC_LAUNCH_CHECK(ctx);
  k_pairs_compact<<<nblk(tsize), 256, 0, st>>>(table, tsize, (long long *)pairs_out, pair_cap, cnt);
  HC_LAUNCH_CHECK(ctx);
  HC_CUDA(cudaMemcpyAsync(ctx->h_mail, cnt, 8, cudaMemcpyDeviceToHost, st));
  HC_CUDA(cudaStreamSynchronize(st));
  if (ctx->h_mail[1] || (long long)ctx->h_mail[0] > pair_cap) {
    hc_set_error("hc_pit_catalog: more than %lld distinct (pit, carve) pairs", (long long)pair_cap);
    return HC_ERR_TOO_LARGE;
  }
  const u32 np = ctx->h_mail[0];
  if (npairs) *npairs = np;
  HC_CUDA(cudaMemsetAsync(pit_out, 0, P1, st));
  k_fill_f32<<<nblk(P1), 256, 0, st>>>(pit_mincarve, INFINITY, P1);
  HC_LAUNCH_CHECK(ctx);
  HC_CUDA(cudaMemsetAsync(pit_sumcarve, 0, P1 * 8, st));
  HC_CUDA(cudaMemsetAsync(ncarvecells, 0, P1 * 8, st));
  if (np) {
    k_pair_flags<<<nblk(np), 256, 0, st>>>((const long long *)pairs_out, np, fillID, (const long long *)carve_amax,
                                            (const long long *)carve_amin, carveout, carveout_id, pair_included, pit_out);
    HC_LAUNCH_CHECK(ctx);
    k_pair_reduce<<<nblk(np), 256, 0, st>>>((const long long *)pairs_out, np, pair_included, pit_out, carve_min, carve_sum,
                                             (const long long *)carve_n, nullptr, pit_mincarve, pit_sumcarve,
                                             (long long *)ncarvecells, nullptr);
    HC_LAUNCH_CHECK(ctx);
  }
  k_catalog_finish<<<nblk(P1), 256, 0, st>>>(npits, pit_out, pit_mincarve, pit_sumcarve, (const long long *)ncarvecells, fb_sum,
                                             fb_max, fb_n, maxcarve, volcarve, volfb, maxfb, complete, volfill64, volfill);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

// ------------------------------------------------------------------------------------------------
// select_solution (tools/depressions.py:517-810): the per-pit decision tree.
struct select_params {
  double fill_vol_thresh, carve_vol_thresh, max_carve_depth, maxcarve_factor;
  int apply_shallow_fill, have_wall, have_sink;
};
// R = the type numpy evaluates the reference's scalar expressions in.  The catalog row handed to the
// tree is a float32 Series; mixed with Python floats it computes in float64 under numpy < 2 (legacy
// value-based promotion: the reference's own era) and in float32 under numpy >= 2 (weak scalars).  The
// two differ exactly at the thresholds (a maxfill of float32(0.01) is "< 0.01" only in float64).
template <typename R>
__global__ void k_select(int npits, const float *maxfill, const float *volfill, const float *maxcarve, const float *volcarve,
                         const float *volfb, const float *maxfb, const int8_t *complete, const long long *nfill,
                         const long long *ncarvecells, const long long *water_in_carve, const float *wall_min,
                         const float *sink_max, select_params prm, int32_t *solution) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < 1 || p > npits) return;
  const R mf = maxfill[p], vf = volfill[p], mc = maxcarve[p], vc = volcarve[p], vfb = volfb[p], mfb = maxfb[p];
  const R ncell = (R)nfill[p];
  const R fvt = (R)prm.fill_vol_thresh, cvt = (R)prm.carve_vol_thresh, mcd = (R)prm.max_carve_depth;
  const R c001 = (R)0.01, c1 = (R)1.0, c09 = (R)0.9;
  int sol;
  if (prm.have_wall && wall_min[p] == 1.0f) { solution[p] = 1; return; }         // pit within wall (:618-623)
  if (prm.have_sink && sink_max[p] > 0.0f) { solution[p] = 7; return; }          // sink within pit (:625-631)
  R limit = (R)prm.maxcarve_factor * mf;                                         // :637
  if (water_in_carve[p] > 0) limit += (R)5;                                      // :641-642
  if (prm.apply_shallow_fill) {                                                  // :644-653
    if (mf / ncell < c001 && vf / ncell < c1) { solution[p] = 2006; return; }
  }
  if (mf < c001) { solution[p] = 2005; return; }                                 // :655-659
  const R max_change = (mc < 0 ? -mc : mc) + (mfb < 0 ? -mfb : mfb);             // :662
  const R vtot = (vc < 0 ? -vc : vc) + (vfb < 0 ? -vfb : vfb);                   // :663
  const R avfb = vfb < 0 ? -vfb : vfb;
  if (complete[p] == 1 && max_change <= mcd && mc <= limit) {                     // valid carve (:665)
    if (water_in_carve[p] >= ncarvecells[p]) sol = 1002;                         // carve inside the water mask (:668)
    else if (vtot < vf) {                                                        // carve preferred
      if (vtot > fvt) {
        if (mf / ncell < c001 && vf / ncell < c1) sol = 1006;
        else if (avfb >= c09 * vf) sol = 2004;
        else sol = 4000;
      } else if (vc < cvt) sol = 1200;
      else if (vf < fvt) sol = 2000;
      else sol = 4000;
    } else {                                                                     // fill preferred (:727)
      if (vf < fvt) sol = vc < cvt ? 2100 : 2000;
      else sol = avfb >= c09 * vf ? 2004 : 4000;
    }
  } else {                                                                       // :740-744
    sol = vf < fvt ? 2009 : 4000;
  }
  solution[p] = sol;
}

__global__ void k_nonzero_sel(const float *w, size_t n, int32_t *o) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = w[i] != 0.f ? 1 : 0;
}

// wall / water / sink rasters arrive as float32 (host converts); any of them may be NULL
extern "C" int hc_pit_select_dev(hc_ctx *ctx, const int32_t *fillID, const int32_t *carve_bound, int64_t n, int32_t npits,
                                 int32_t ncarves, const float *maxfill, const float *volfill, const float *maxcarve,
                                 const float *volcarve, const float *volfb, const float *maxfb, const int8_t *complete,
                                 const int64_t *nfill, const int64_t *ncarvecells, const uint8_t *pit_out,
                                 const int64_t *pairs, const uint8_t *pair_included, int64_t npairs, const float *wall,
                                 const float *water, const float *sink, double fill_vol_thresh, double carve_vol_thresh,
                                 double max_carve_depth, double maxcarve_factor, int apply_shallow_fill,
                                 int float32_scalars, int32_t *solution, void *stream) {
  if (!ctx) { hc_set_error("null context"); return HC_ERR_ARG; }
  hc_guard hc_guard__(ctx);
  HC_CUDA(cudaSetDevice(ctx->device));
  if (!water) { hc_set_error("select_solution: water_mask_array is None (the reference raises NameError here, tools/depressions.py:606-647)"); return HC_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  HC_TRY(arena_reset(ctx));
  const size_t P1 = (size_t)npits + 1, C1 = (size_t)ncarves + 1;
  float *wall_min = (float *)arena_take(ctx, P1 * 4);
  float *sink_max = (float *)arena_take(ctx, P1 * 4);
  long long *carve_water = (long long *)arena_take(ctx, C1 * 8);
  long long *pit_water = (long long *)arena_take(ctx, P1 * 8);
  int32_t *wsel = (int32_t *)arena_take(ctx, (size_t)n * 4);
  if (!wall_min || !sink_max || !carve_water || !pit_water || !wsel) return HC_ERR_NOMEM;
  if (wall) HC_TRY(hc_seg_stats_dev(ctx, fillID, wall, nullptr, n, npits, 1, nullptr, wall_min, nullptr, nullptr, stream));
  if (sink) HC_TRY(hc_seg_stats_dev(ctx, fillID, sink, nullptr, n, npits, 1, sink_max, nullptr, nullptr, nullptr, stream));
  // np.count_nonzero(water_mask_vec[carve_cells]) per carve, then summed over a pit's included carves
  k_nonzero_sel<<<nblk((size_t)n), 256, 0, st>>>(water, (size_t)n, wsel);
  HC_LAUNCH_CHECK(ctx);
  HC_TRY(hc_seg_stats_dev(ctx, carve_bound, nullptr, wsel, n, ncarves, 1, nullptr, nullptr, nullptr, (int64_t *)carve_water, stream));
  HC_CUDA(cudaMemsetAsync(pit_water, 0, P1 * 8, st));
  if (npairs) {
    k_pair_reduce<<<nblk((size_t)npairs), 256, 0, st>>>((const long long *)pairs, (u32)npairs, pair_included, pit_out, nullptr,
                                                         nullptr, nullptr, carve_water, nullptr, nullptr, nullptr, pit_water);
    HC_LAUNCH_CHECK(ctx);
  }
  select_params prm = {fill_vol_thresh, carve_vol_thresh, max_carve_depth, maxcarve_factor, apply_shallow_fill,
                       wall ? 1 : 0, sink ? 1 : 0};
  HC_CUDA(cudaMemsetAsync(solution, 0, 4, st));
  HC_PROF(ctx, st, "k_select");
  if (float32_scalars)
    k_select<float><<<nblk(P1), 256, 0, st>>>(npits, maxfill, volfill, maxcarve, volcarve, volfb, maxfb, complete,
                                              (const long long *)nfill, (const long long *)ncarvecells, pit_water, wall_min,
                                              sink_max, prm, solution);
  else
    k_select<double><<<nblk(P1), 256, 0, st>>>(npits, maxfill, volfill, maxcarve, volcarve, volfb, maxfb, complete,
                                               (const long long *)nfill, (const long long *)ncarvecells, pit_water, wall_min,
                                               sink_max, prm, solution);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}

// ------------------------------------------------------------------------------------------------
// apply_solution (tools/depressions.py:1140-1228).  The reference loops over the pits in index order
// and later pits overwrite earlier ones; a cell's final value therefore comes from the HIGHEST pit
// index that writes it: its own pit (fill solution -> fill value; carve solution and fill-behind cell
// -> carved value) or, for a cell of an applied carve, the highest carve-solution pit owning that carve.
__global__ void k_carve_owner(const long long *pairs, u32 np, const uint8_t *included, const uint8_t *pit_out,
                              const int32_t *solution, const uint8_t *excluded, const long long *ncarvecells,
                              int32_t *carve_owner) {
  u32 j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= np || !included[j]) return;
  int p = (int)((u64)pairs[j] >> 32), c = (int)((u64)pairs[j] & 0xFFFFFFFFu);
  if (pit_out[p] || ncarvecells[p] <= 0) return;      // carvecells_dct[pit] is empty
  if (excluded && excluded[p]) return;
  int s = solution[p];
  if (s >= 1000 && s < 2000) atomicMax(&carve_owner[c], p);
}
__global__ void k_apply_solution(const float *__restrict__ dem, const float *__restrict__ fillv, const float *__restrict__ carvev,
                                 const int32_t *__restrict__ fillID, const int32_t *__restrict__ carve_bound,
                                 const int32_t *__restrict__ carvefillID, size_t n, const int32_t *__restrict__ solution,
                                 const uint8_t *__restrict__ excluded, const uint8_t *__restrict__ pit_out,
                                 const long long *__restrict__ ncarvecells, const int32_t *__restrict__ carve_owner,
                                 float *__restrict__ out, int16_t *__restrict__ mask) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = dem[i];
  int p = fillID[i];
  int writer = 0;      // highest pit index that has written this cell so far
  if (p > 0 && solution[p] >= 0) {   // a negative code = the pit is not in the table handed to apply_solution
    int s = solution[p];
    mask[i] = (int16_t)s;
    if (!(excluded && excluded[p])) {
      if (s >= 2000 && s < 3000) { v = fillv[i]; writer = p; }
      else if (s >= 1000 && s < 2000) {
        // fill behind the carve: only when the pit has carve cells at all (:445-452)
        if (carvefillID[i] > 0 && !pit_out[p] && ncarvecells[p] > 0) { v = carvev[i]; writer = p; }
      }
    }
  }
  int c = carve_bound[i];
  if (c > 0) {
    int q = carve_owner[c];
    if (q > 0 && q >= writer) v = carvev[i];
  }
  out[i] = v;
}

extern "C" int hc_pit_apply_dev(hc_ctx *ctx, const float *dem, const float *fillv, const float *carvev,
                                const int32_t *fillID, const int32_t *carve_bound, const int32_t *carvefillID, int64_t n,
                                int32_t npits, int32_t ncarves, const int32_t *solution, const uint8_t *excluded,
                                const uint8_t *pit_out, const int64_t *ncarvecells, const int64_t *pairs,
                                const uint8_t *pair_included, int64_t npairs, float *out, int16_t *mask, void *stream) {
  if (!ctx) { hc_set_error("null context"); return HC_ERR_ARG; }
  hc_guard hc_guard__(ctx);
  HC_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  HC_TRY(arena_reset(ctx));
  int32_t *owner = (int32_t *)arena_take(ctx, ((size_t)ncarves + 1) * 4);
  if (!owner) return HC_ERR_NOMEM;
  HC_CUDA(cudaMemsetAsync(owner, 0, ((size_t)ncarves + 1) * 4, st));
  if (npairs) {
    k_carve_owner<<<nblk((size_t)npairs), 256, 0, st>>>((const long long *)pairs, (u32)npairs, pair_included, pit_out, solution,
                                                         excluded, (const long long *)ncarvecells, owner);
    HC_LAUNCH_CHECK(ctx);
  }
  HC_PROF(ctx, st, "k_apply_solution");
  k_apply_solution<<<nblk((size_t)n), 256, 0, st>>>(dem, fillv, carvev, fillID, carve_bound, carvefillID, (size_t)n, solution,
                                                     excluded, pit_out, (const long long *)ncarvecells, owner, out, mask);
  HC_LAUNCH_CHECK(ctx);
  (void)npits;
  return HC_OK;
}

// ------------------------------------------------------------------------------------------------
// fill_shallow_pits (tools/depressions.py:1471-1531): a label (0 = background included, as in the
// reference's np.isin logic) keeps its cells when any of them is deeper than sfill or touches the
// water mask; every other cell takes the filled value.
__global__ void k_shallow_flags(const int32_t *__restrict__ ids, const float *__restrict__ diff, const float *__restrict__ water,
                                size_t n, float sfill, int nseg, uint8_t *keep) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int id = ids[i];
  if (id < 0 || id > nseg) return;
  if (diff[i] > sfill || (water && water[i] != 0.f)) keep[id] = 1;
}
__global__ void k_shallow_apply(const int32_t *__restrict__ ids, const float *__restrict__ fillv, size_t n, int nseg,
                                const uint8_t *__restrict__ keep, float *__restrict__ mod) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int id = ids[i];
  if (id < 0 || id > nseg) return;
  if (!keep[id]) mod[i] = fillv[i];
}
extern "C" int hc_fill_shallow_pits_dev(hc_ctx *ctx, float *mod, float sfill, const int32_t *fillID, const float *diff_fill,
                                        const float *fillv, const float *water, int64_t n, int32_t npits,
                                        int64_t *n_shallow, void *stream) {
  if (!ctx) { hc_set_error("null context"); return HC_ERR_ARG; }
  hc_guard hc_guard__(ctx);
  HC_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  HC_TRY(arena_reset(ctx));
  uint8_t *keep = (uint8_t *)arena_take(ctx, (size_t)npits + 1);
  if (!keep) return HC_ERR_NOMEM;
  HC_CUDA(cudaMemsetAsync(keep, 0, (size_t)npits + 1, st));
  HC_PROF(ctx, st, "k_shallow_flags");
  k_shallow_flags<<<nblk((size_t)n), 256, 0, st>>>(fillID, diff_fill, water, (size_t)n, sfill, npits, keep);
  HC_LAUNCH_CHECK(ctx);
  HC_PROF(ctx, st, "k_shallow_apply");
  k_shallow_apply<<<nblk((size_t)n), 256, 0, st>>>(fillID, fillv, (size_t)n, npits, keep, mod);
  HC_LAUNCH_CHECK(ctx);
  if (n_shallow) {
    // number of labels (background included) that were filled, as the reference prints it
    uint8_t *h = (uint8_t *)malloc((size_t)npits + 1);
    if (!h) return HC_ERR_NOMEM;
    cudaError_t e1 = cudaMemcpyAsync(h, keep, (size_t)npits + 1, cudaMemcpyDeviceToHost, st);
    cudaError_t e2 = e1 == cudaSuccess ? cudaStreamSynchronize(st) : e1;
    if (e2 != cudaSuccess) {
      free(h);
      hc_set_error("hc_fill_shallow_pits: read-back -> %s", cudaGetErrorString(e2));
      return HC_ERR_CUDA;
    }
    int64_t c = 0;
    for (int32_t p = 0; p <= npits; p++) c += h[p] ? 0 : 1;
    free(h);
    *n_shallow = c;
  }
  return HC_OK;
}

// ------------------------------------------------------------------------------------------------
// exact order statistics (np.percentile of endo.find_sink, tools/endo.py:286): radix select over the
// monotone uint image of the floats, 8 bits per pass (4 passes, one 256-bin histogram each)
__global__ void __launch_bounds__(256)
k_radix_hist(const float *__restrict__ x, size_t n, u32 prefix, u32 prefix_mask, int shift, unsigned long long *hist) {
  __shared__ u32 sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    u32 u = d_ford(x[i]);
    if ((u & prefix_mask) == prefix) atomicAdd(&sh[(u >> shift) & 255u], 1u);
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

// value of the k-th smallest element (0-based) of x[0..n); x must not contain NaN.  Host pointer out.
extern "C" int hc_kth_value_dev(hc_ctx *ctx, const float *x, int64_t n, int64_t k, float *value, void *stream) {
  if (!ctx) { hc_set_error("null context"); return HC_ERR_ARG; }
  hc_guard hc_guard__(ctx);
  HC_CUDA(cudaSetDevice(ctx->device));
  if (!x || !value || n <= 0 || k < 0 || k >= n) { hc_set_error("hc_kth_value: bad argument"); return HC_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  HC_TRY(arena_reset(ctx));
  unsigned long long *hist = (unsigned long long *)arena_take(ctx, 256 * 8);
  if (!hist) return HC_ERR_NOMEM;
  unsigned long long hh[256];
  u32 prefix = 0, mask = 0;
  unsigned long long rank = (unsigned long long)k;
  for (int shift = 24; shift >= 0; shift -= 8) {
    HC_CUDA(cudaMemsetAsync(hist, 0, 256 * 8, st));
    HC_PROF(ctx, st, "k_radix_hist");
    k_radix_hist<<<ctx->sm_count * 8, 256, 0, st>>>(x, (size_t)n, prefix, mask, shift, hist);
    HC_LAUNCH_CHECK(ctx);
    HC_CUDA(cudaMemcpyAsync(hh, hist, 256 * 8, cudaMemcpyDeviceToHost, st));
    HC_CUDA(cudaStreamSynchronize(st));
    int d = 0;
    for (; d < 256; d++) {
      if (rank < hh[d]) break;
      rank -= hh[d];
    }
    if (d == 256) { hc_set_error("hc_kth_value: internal (NaN in input?)"); return HC_ERR_INTERNAL; }
    prefix |= (u32)d << shift;
    mask |= 255u << shift;
  }
  u32 b = (prefix & 0x80000000u) ? (prefix & 0x7FFFFFFFu) : ~prefix;
  memcpy(value, &b, 4);
  return HC_OK;
}
