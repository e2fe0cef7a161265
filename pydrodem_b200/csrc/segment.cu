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
             u64 *kmax, u64 *kmin, u32 cell0) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int id = ids[i];
  if (id < 1 || id > nseg) return;
  u32 o = d_ford(vals[i]);
  const u32 gi = cell0 + (u32)i;   // index in the whole raster (cell0: a row band's first cell)
  // max value, smallest index on ties  <=> max of (o, ~index)
  u64 k1 = ((u64)o << 32) | (u64)(0xFFFFFFFFu - gi);
  if (k1 > kmax[id]) atomicMax(&kmax[id], k1);
  u64 k2 = ((u64)o << 32) | (u64)gi;
  if (k2 < kmin[id]) atomicMin(&kmin[id], k2);
}
// pit id at every carve's first highest / first lowest cell, as far as this band owns that cell (else 0): the two
// look-ups of the inclusion test (:435-439).  -1: the carve has no cell at all.
__global__ void k_argext_owner_ids(const u64 *kmax, const u64 *kmin, int nseg, const int32_t *__restrict__ fillID, u32 cell0,
                                   size_t n, int32_t *fid_amax, int32_t *fid_amin) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > nseg) return;
  int32_t a = -1, b = -1;
  if (kmax[s] != 0ull) {
    const u32 g = 0xFFFFFFFFu - (u32)(kmax[s] & 0xFFFFFFFFu);
    a = (g >= cell0 && (size_t)(g - cell0) < n) ? fillID[g - cell0] : 0;
  }
  if (kmin[s] != ~0ull) {
    const u32 g = (u32)(kmin[s] & 0xFFFFFFFFu);
    b = (g >= cell0 && (size_t)(g - cell0) < n) ? fillID[g - cell0] : 0;
  }
  fid_amax[s] = a;
  fid_amin[s] = b;
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
__global__ void k_pair_flags(const long long *pairs, u32 np, const int32_t *__restrict__ fid_amax,
                             const int32_t *__restrict__ fid_amin, int carveout, int carveout_id,
                             uint8_t *included, uint8_t *pit_out) {
  u32 j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= np) return;
  int p = (int)((u64)pairs[j] >> 32), c = (int)((u64)pairs[j] & 0xFFFFFFFFu);
  // "does carve leave basin?": the carve that contains the nodata pixels touches this pit
  if (!carveout && c == carveout_id) pit_out[p] = 1;
  bool inc = false;
  if (fid_amax[c] >= 0) inc = (fid_amax[c] == p) || (fid_amin[c] != p);
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

// build_pit_catalog (tools/depressions.py:272-509) in three steps, so that a raster cut into row bands can run the
// first on every band, reduce the per-label tables across the bands (max / min / sum; the arg-extreme keys carry whole-
// raster cell indices) and run the other two on the reduced tables ("AllReduce of per-label partial
// max / sum").  hc_pit_catalog_dev is the three back to back on one raster.
struct catalog_partials {
  float *maxfill; double *volfill64; long long *nfill;          // per pit: diff_fill over its cells (:392-394)
  float *fb_max; double *fb_sum; long long *fb_n;               // per pit: fill behind the carve (:395,449,457)
  float *carve_min; double *carve_sum; long long *carve_n;      // per carve: diff_carve over its bounded cells
  u64 *kmax, *kmin;                                             // per carve: first arg-max / arg-min of carve_vec (:421-424)
};

// step 1: everything that is a reduction over cells.  *npairs (host) = distinct (pit, carve) pairs of these cells.
static int catalog_local(hc_ctx *ctx, const float *dem, float nodata, const float *carve, const float *diff_carve,
                         const float *diff_fill, const int32_t *fillID, const int32_t *carveID, const int32_t *carvefillID,
                         int64_t n, u32 cell0, int32_t npits, int32_t ncarves, int32_t *carve_bound, const catalog_partials &t,
                         int64_t *pairs_out, int64_t pair_cap, int64_t *npairs, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const size_t C1 = (size_t)ncarves + 1;
  u32 *cnt = (u32 *)arena_take(ctx, 256);
  if (!cnt) return HC_ERR_NOMEM;
  HC_PROF(ctx, st, "k_carve_bound");
  k_carve_bound<<<nblk((size_t)n), 256, 0, st>>>(carveID, dem, nodata, (size_t)n, carve_bound);
  HC_LAUNCH_CHECK(ctx);
  HC_TRY(hc_seg_stats_dev(ctx, fillID, diff_fill, nullptr, n, npits, 1, t.maxfill, nullptr, t.volfill64, (int64_t *)t.nfill, stream));
  HC_TRY(hc_seg_stats_dev(ctx, fillID, diff_carve, carvefillID, n, npits, 1, t.fb_max, nullptr, t.fb_sum, (int64_t *)t.fb_n, stream));
  HC_TRY(hc_seg_stats_dev(ctx, carve_bound, diff_carve, nullptr, n, ncarves, 1, nullptr, t.carve_min, t.carve_sum, (int64_t *)t.carve_n, stream));
  HC_CUDA(cudaMemsetAsync(t.kmax, 0, C1 * 8, st));
  HC_CUDA(cudaMemsetAsync(t.kmin, 0xFF, C1 * 8, st));
  HC_PROF(ctx, st, "k_seg_argext");
  k_seg_argext<<<nblk((size_t)n), 256, 0, st>>>(carve_bound, carve, (size_t)n, ncarves, t.kmax, t.kmin, cell0);
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
  if (npairs) *npairs = ctx->h_mail[0];
  return HC_OK;
}

// step 3: pair-level logic and the per-pit carve columns from the (reduced) tables
static int catalog_finish(hc_ctx *ctx, int32_t npits, int32_t ncarves, const int64_t *pairs, int64_t np, const int32_t *fid_amax,
                          const int32_t *fid_amin, int carveout, int32_t carveout_id, const catalog_partials &t,
                          float *volfill, float *maxcarve, float *volcarve, int64_t *ncarvecells, float *volfb, float *maxfb,
                          int8_t *complete, uint8_t *pit_out, uint8_t *pair_included, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const size_t P1 = (size_t)npits + 1;
  float *pit_mincarve = (float *)arena_take(ctx, P1 * 4);
  double *pit_sumcarve = (double *)arena_take(ctx, P1 * 8);
  if (!pit_mincarve || !pit_sumcarve) return HC_ERR_NOMEM;
  HC_CUDA(cudaMemsetAsync(pit_out, 0, P1, st));
  k_fill_f32<<<nblk(P1), 256, 0, st>>>(pit_mincarve, INFINITY, P1);
  HC_LAUNCH_CHECK(ctx);
  HC_CUDA(cudaMemsetAsync(pit_sumcarve, 0, P1 * 8, st));
  HC_CUDA(cudaMemsetAsync(ncarvecells, 0, P1 * 8, st));
  if (np) {
    k_pair_flags<<<nblk((size_t)np), 256, 0, st>>>((const long long *)pairs, (u32)np, fid_amax, fid_amin, carveout, carveout_id,
                                                   pair_included, pit_out);
    HC_LAUNCH_CHECK(ctx);
    k_pair_reduce<<<nblk((size_t)np), 256, 0, st>>>((const long long *)pairs, (u32)np, pair_included, pit_out, t.carve_min, t.carve_sum,
                                                    t.carve_n, nullptr, pit_mincarve, pit_sumcarve, (long long *)ncarvecells, nullptr);
    HC_LAUNCH_CHECK(ctx);
  }
  k_catalog_finish<<<nblk(P1), 256, 0, st>>>(npits, pit_out, pit_mincarve, pit_sumcarve, (const long long *)ncarvecells, t.fb_sum,
                                             t.fb_max, t.fb_n, maxcarve, volcarve, volfb, maxfb, complete, t.volfill64, volfill);
  HC_LAUNCH_CHECK(ctx);
  (void)ncarves;
  return HC_OK;
}

// All pointers are device pointers; per-pit outputs have npits+1 entries (entry 0 unused), per-carve outputs
// ncarves+1.  pairs_out receives the distinct (pit << 32 | carve) pairs, pair_included the reference's inclusion
// test per pair; *npairs (host) their number.  carveout_id = carveID at the first nodata cell (:335), ignored when
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
  catalog_partials t;
  t.maxfill = maxfill; t.nfill = (long long *)nfill;
  t.carve_min = carve_min; t.carve_sum = carve_sum; t.carve_n = (long long *)carve_n;
  t.volfill64 = (double *)arena_take(ctx, P1 * 8);
  t.fb_sum = (double *)arena_take(ctx, P1 * 8);
  t.fb_max = (float *)arena_take(ctx, P1 * 4);
  t.fb_n = (long long *)arena_take(ctx, P1 * 8);
  t.kmax = (u64 *)arena_take(ctx, C1 * 8);
  t.kmin = (u64 *)arena_take(ctx, C1 * 8);
  int32_t *fid_amax = (int32_t *)arena_take(ctx, C1 * 4), *fid_amin = (int32_t *)arena_take(ctx, C1 * 4);
  if (!t.volfill64 || !t.fb_sum || !t.fb_max || !t.fb_n || !t.kmax || !t.kmin || !fid_amax || !fid_amin) return HC_ERR_NOMEM;
  int64_t np = 0;
  HC_TRY(catalog_local(ctx, dem, nodata, carve, diff_carve, diff_fill, fillID, carveID, carvefillID, n, 0u, npits, ncarves,
                       carve_bound, t, pairs_out, pair_cap, &np, stream));
  if (npairs) *npairs = np;
  k_argext_decode<<<nblk(C1), 256, 0, st>>>(t.kmax, t.kmin, ncarves, (long long *)carve_amax, (long long *)carve_amin);
  HC_LAUNCH_CHECK(ctx);
  k_argext_owner_ids<<<nblk(C1), 256, 0, st>>>(t.kmax, t.kmin, ncarves, fillID, 0u, (size_t)n, fid_amax, fid_amin);
  HC_LAUNCH_CHECK(ctx);
  return catalog_finish(ctx, npits, ncarves, pairs_out, np, fid_amax, fid_amin, carveout, carveout_id, t, volfill, maxcarve,
                        volcarve, ncarvecells, volfb, maxfb, complete, pit_out, pair_included, stream);
}

// ---- the same over row bands (pydrodem_b200/bands.py pit_catalog_banded) ------------------------------------------------
// step 1 on one band: `cell0` = whole-raster index of the band's first cell, label rasters hold WHOLE-RASTER ids
// (bands.label_banded).  Outputs are this band's partial tables, to be reduced across the bands by the caller: max for
// maxfill / fb_max / kmax (unsigned), min for carve_min / kmin (unsigned), sum for the float64 sums and the counts; the
// pair lists are concatenated and made unique.
extern "C" int hc_band_catalog_local_dev(hc_ctx *ctx, const float *dem, float nodata, const float *carve, const float *diff_carve,
                                         const float *diff_fill, const int32_t *fillID, const int32_t *carveID,
                                         const int32_t *carvefillID, int64_t n, int64_t cell0, int32_t npits, int32_t ncarves,
                                         int32_t *carve_bound, float *maxfill, double *volfill64, int64_t *nfill, float *fb_max,
                                         double *fb_sum, int64_t *fb_n, float *carve_min, double *carve_sum, int64_t *carve_n,
                                         uint64_t *kmax, uint64_t *kmin, int64_t *pairs_out, int64_t pair_cap, int64_t *npairs,
                                         void *stream) {
  HC_ENTER(ctx);
  if (n <= 0 || cell0 < 0 || cell0 + n >= 0x7FFFFFF0ll || npits < 0 || ncarves < 0) { hc_set_error("hc_band_catalog_local: bad shape"); return HC_ERR_ARG; }
  HC_TRY(arena_reset(ctx));
  catalog_partials t = {maxfill, volfill64, (long long *)nfill, fb_max, fb_sum, (long long *)fb_n, carve_min, carve_sum,
                        (long long *)carve_n, (u64 *)kmax, (u64 *)kmin};
  return catalog_local(ctx, dem, nodata, carve, diff_carve, diff_fill, fillID, carveID, carvefillID, n, (u32)cell0, npits, ncarves,
                       carve_bound, t, pairs_out, pair_cap, npairs, stream);
}
// step 2 on one band, after kmax / kmin were reduced: the pit ids at the carves' extreme cells this band owns (0
// elsewhere, -1 for an empty carve): max-reduce across the bands.
extern "C" int hc_band_catalog_owner_ids_dev(hc_ctx *ctx, const int32_t *fillID, int64_t n, int64_t cell0, int32_t ncarves,
                                             const uint64_t *kmax, const uint64_t *kmin, int32_t *fid_amax, int32_t *fid_amin,
                                             void *stream) {
  HC_ENTER(ctx);
  if (n <= 0 || cell0 < 0 || ncarves < 0 || !fillID || !kmax || !kmin || !fid_amax || !fid_amin) { hc_set_error("hc_band_catalog_owner_ids: bad argument"); return HC_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  k_argext_owner_ids<<<nblk((size_t)ncarves + 1), 256, 0, st>>>((const u64 *)kmax, (const u64 *)kmin, ncarves, fillID, (u32)cell0, (size_t)n,
                                                               fid_amax, fid_amin);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}
// step 3 on the reduced tables (every rank, redundantly: the tables are small): the catalog's columns
extern "C" int hc_catalog_finish_dev(hc_ctx *ctx, int32_t npits, int32_t ncarves, const int64_t *pairs, int64_t npairs,
                                     const int32_t *fid_amax, const int32_t *fid_amin, int carveout, int32_t carveout_id,
                                     const double *volfill64, const float *fb_max, const double *fb_sum, const int64_t *fb_n,
                                     const float *carve_min, const double *carve_sum, const int64_t *carve_n, float *volfill,
                                     float *maxcarve, float *volcarve, int64_t *ncarvecells, float *volfb, float *maxfb,
                                     int8_t *complete, uint8_t *pit_out, uint8_t *pair_included, void *stream) {
  HC_ENTER(ctx);
  if (npits < 0 || ncarves < 0 || npairs < 0) { hc_set_error("hc_catalog_finish: bad shape"); return HC_ERR_ARG; }
  HC_TRY(arena_reset(ctx));
  catalog_partials t = {nullptr, (double *)volfill64, nullptr, (float *)fb_max, (double *)fb_sum, (long long *)fb_n,
                        (float *)carve_min, (double *)carve_sum, (long long *)carve_n, nullptr, nullptr};
  return catalog_finish(ctx, npits, ncarves, pairs, npairs, fid_amax, fid_amin, carveout, carveout_id, t, volfill, maxcarve, volcarve,
                        ncarvecells, volfb, maxfb, complete, pit_out, pair_included, stream);
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

// select_solution in two steps, split like the catalog: the reductions over cells (per pit: is it inside the wall, does
// it hold a sink; per carve: how many of its cells lie in the water mask), then the decision tree on the tables.
static int select_partials(hc_ctx *ctx, const int32_t *fillID, const int32_t *carve_bound, int64_t n, int32_t npits,
                           int32_t ncarves, const float *wall, const float *water, const float *sink, float *wall_min,
                           float *sink_max, int64_t *carve_water, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  int32_t *wsel = (int32_t *)arena_take(ctx, (size_t)n * 4);
  if (!wsel) return HC_ERR_NOMEM;
  if (wall) HC_TRY(hc_seg_stats_dev(ctx, fillID, wall, nullptr, n, npits, 1, nullptr, wall_min, nullptr, nullptr, stream));
  if (sink) HC_TRY(hc_seg_stats_dev(ctx, fillID, sink, nullptr, n, npits, 1, sink_max, nullptr, nullptr, nullptr, stream));
  // np.count_nonzero(water_mask_vec[carve_cells]) per carve
  k_nonzero_sel<<<nblk((size_t)n), 256, 0, st>>>(water, (size_t)n, wsel);
  HC_LAUNCH_CHECK(ctx);
  return hc_seg_stats_dev(ctx, carve_bound, nullptr, wsel, n, ncarves, 1, nullptr, nullptr, nullptr, carve_water, stream);
}

static int select_tables(hc_ctx *ctx, int32_t npits, const float *maxfill, const float *volfill, const float *maxcarve,
                         const float *volcarve, const float *volfb, const float *maxfb, const int8_t *complete,
                         const int64_t *nfill, const int64_t *ncarvecells, const uint8_t *pit_out, const int64_t *pairs,
                         const uint8_t *pair_included, int64_t npairs, const float *wall_min, const float *sink_max,
                         const int64_t *carve_water, double fill_vol_thresh, double carve_vol_thresh, double max_carve_depth,
                         double maxcarve_factor, int apply_shallow_fill, int float32_scalars, int32_t *solution, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const size_t P1 = (size_t)npits + 1;
  long long *pit_water = (long long *)arena_take(ctx, P1 * 8);
  if (!pit_water) return HC_ERR_NOMEM;
  HC_CUDA(cudaMemsetAsync(pit_water, 0, P1 * 8, st));
  if (npairs) {   // ... summed over a pit's included carves
    k_pair_reduce<<<nblk((size_t)npairs), 256, 0, st>>>((const long long *)pairs, (u32)npairs, pair_included, pit_out, nullptr,
                                                         nullptr, nullptr, (const long long *)carve_water, nullptr, nullptr, nullptr,
                                                         pit_water);
    HC_LAUNCH_CHECK(ctx);
  }
  select_params prm = {fill_vol_thresh, carve_vol_thresh, max_carve_depth, maxcarve_factor, apply_shallow_fill,
                       wall_min ? 1 : 0, sink_max ? 1 : 0};
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
  HC_TRY(arena_reset(ctx));
  const size_t P1 = (size_t)npits + 1, C1 = (size_t)ncarves + 1;
  float *wall_min = (float *)arena_take(ctx, P1 * 4);
  float *sink_max = (float *)arena_take(ctx, P1 * 4);
  int64_t *carve_water = (int64_t *)arena_take(ctx, C1 * 8);
  if (!wall_min || !sink_max || !carve_water) return HC_ERR_NOMEM;
  HC_TRY(select_partials(ctx, fillID, carve_bound, n, npits, ncarves, wall, water, sink, wall_min, sink_max, carve_water, stream));
  return select_tables(ctx, npits, maxfill, volfill, maxcarve, volcarve, volfb, maxfb, complete, nfill, ncarvecells, pit_out, pairs,
                       pair_included, npairs, wall ? wall_min : nullptr, sink ? sink_max : nullptr, carve_water, fill_vol_thresh,
                       carve_vol_thresh, max_carve_depth, maxcarve_factor, apply_shallow_fill, float32_scalars, solution, stream);
}

// over row bands: the per-band partial tables (min-reduce wall_min, max-reduce sink_max, sum carve_water across the
// bands), then the tree on the reduced tables (every rank, redundantly).  wall / sink may be NULL (then the table is
// left untouched and NULL must be passed to hc_pit_select_tables_dev too).
extern "C" int hc_band_select_partials_dev(hc_ctx *ctx, const int32_t *fillID, const int32_t *carve_bound, int64_t n, int32_t npits,
                                           int32_t ncarves, const float *wall, const float *water, const float *sink,
                                           float *wall_min, float *sink_max, int64_t *carve_water, void *stream) {
  HC_ENTER(ctx);
  if (!water) { hc_set_error("select_solution: water_mask_array is None (the reference raises NameError here, tools/depressions.py:606-647)"); return HC_ERR_ARG; }
  if (n <= 0 || npits < 0 || ncarves < 0 || !fillID || !carve_bound || !carve_water || (wall && !wall_min) || (sink && !sink_max)) {
    hc_set_error("hc_band_select_partials: bad argument");
    return HC_ERR_ARG;
  }
  HC_TRY(arena_reset(ctx));
  return select_partials(ctx, fillID, carve_bound, n, npits, ncarves, wall, water, sink, wall_min, sink_max, carve_water, stream);
}
extern "C" int hc_pit_select_tables_dev(hc_ctx *ctx, int32_t npits, const float *maxfill, const float *volfill, const float *maxcarve,
                                        const float *volcarve, const float *volfb, const float *maxfb, const int8_t *complete,
                                        const int64_t *nfill, const int64_t *ncarvecells, const uint8_t *pit_out,
                                        const int64_t *pairs, const uint8_t *pair_included, int64_t npairs, const float *wall_min,
                                        const float *sink_max, const int64_t *carve_water, double fill_vol_thresh,
                                        double carve_vol_thresh, double max_carve_depth, double maxcarve_factor,
                                        int apply_shallow_fill, int float32_scalars, int32_t *solution, void *stream) {
  HC_ENTER(ctx);
  if (npits < 0 || npairs < 0 || !carve_water || !solution) { hc_set_error("hc_pit_select_tables: bad argument"); return HC_ERR_ARG; }
  HC_TRY(arena_reset(ctx));
  return select_tables(ctx, npits, maxfill, volfill, maxcarve, volcarve, volfb, maxfb, complete, nfill, ncarvecells, pit_out, pairs,
                       pair_included, npairs, wall_min, sink_max, carve_water, fill_vol_thresh, carve_vol_thresh, max_carve_depth,
                       maxcarve_factor, apply_shallow_fill, float32_scalars, solution, stream);
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
