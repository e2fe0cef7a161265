/*
 * hydrocore.h — C ABI of libhydrocore.so (hand-written sm_100a CUDA, no torch types).
 *
 * This is the drop-in boundary for pydrodem's depression-fill / D8 / flat-resolution /
 * accumulation / labelling path.  pydrodem itself reaches that arithmetic by
 * writing GeoTIFFs and spawning WhiteboxTools / TauDEM executables; each entry point below
 * names the reference call site (file:line under /root/reference) whose native work it
 * replaces.  Python binds these with ctypes (pydrodem_b200/_lib.py); INTEGRATION.md shows the
 * stub a pydrodem maintainer would add.
 *
 * Conventions
 *   - rasters are row-major, north-up, ny rows x nx columns, ny*nx < 2^31;
 *   - a cell is "nodata" when z == nodata or z is NaN;
 *   - every function returns 0 on success, a negative hc_status otherwise, and
 *     hc_last_error() then returns a static thread-local message;
 *   - *_dev entry points take DEVICE pointers and a cudaStream_t (as void*) and are
 *     asynchronous apart from a few 4-byte read-backs; *_host entry points take HOST pointers
 *     and do the H2D / D2H copies themselves;
 *   - there is no CPU fallback: without a CUDA device hc_create fails.
 */
#ifndef HYDROCORE_H
#define HYDROCORE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hc_ctx hc_ctx;

enum hc_status {
  HC_OK = 0,
  HC_ERR_CUDA = -1,
  HC_ERR_ARG = -2,
  HC_ERR_NOMEM = -3,
  HC_ERR_TOO_LARGE = -4,
  HC_ERR_INTERNAL = -5
};

/* seed conventions of the two fills pydrodem calls */
#define HC_SEED_WBT 0    /* wbt.fill_depressions_wang_and_liu: tools/depressions.py:1284 */
#define HC_SEED_TAUDEM 1 /* TauDEM PitRemove: run_aws.py:582-588 */
/* D8 code tables */
#define HC_TABLE_WBT 0    /* wbt.d8_pointer: run_aws.py:539 (NE=1,E=2,SE=4,...,N=128; 0 none) */
#define HC_TABLE_TAUDEM 1 /* TauDEM D8FlowDir -p: run_aws.py:594-603 (1=E .. 8=SE, ccw)      */
#define HC_DIR_NOFLOW 0
#define HC_DIR_NODATA 255

const char *hc_last_error(void);
int hc_version(void);

/* A context owns the workspace arena, a pinned mailbox, the *_host staging buffers and the band state.
 * Threading contract: every entry point that takes a context holds the context's (recursive) lock for the
 * duration of the call, so concurrent callers of ONE context serialise; the device work a call enqueues runs on the
 * stream it was given.  Calls on different streams through one context are therefore safe but not concurrent (each
 * public call rewinds the one arena): use one context per thread / stream for concurrency (hc_create is cheap).
 * hc_fill_host / hc_label_host share staging buffers with hc_fill_d8_accum_host_begin and wait for its copies. */
int hc_create(hc_ctx **ctx, int device);
int hc_destroy(hc_ctx *ctx);
/* bytes of device workspace currently held */
int64_t hc_workspace_bytes(const hc_ctx *ctx);
/* number of kernels this context has launched since creation (bench.py "gpu_launches") */
int64_t hc_launch_count(const hc_ctx *ctx);

/* Depression fill, epsilon = 0: out = max(z, lowest spill level to a seed).
 * Replaces wbt.fill_depressions_wang_and_liu / wbt.fill_depressions behind dep.fill_pits
 * (tools/depressions.py:1236-1316, :1277, :1284; final fill models/depression_handling.py:480)
 * and TauDEM PitRemove (run_aws.py:582-588).  z and out may not alias. */
int hc_fill_dev(hc_ctx *ctx, const float *z, float *out, int64_t ny, int64_t nx, float nodata,
                int seed_mode, void *stream);

/* Priority-Flood + epsilon (WhiteboxTools fill with fix_flats / flat_increment > 0; dep.fill_pits' second fill,
 * tools/depressions.py:195-198, :1271-1286): float64 surface W = the unique solution of
 * W(c) = max(z(c), min over neighbours W(n) + eps), W = z on the seeds, reached by monotone sweeps from the eps = 0
 * fill.  `sweeps` (optional) receives the number of raster sweeps.  Tolerance class: the serial flood makes the same
 * float64 additions, WhiteboxTools' own heap order is unknown (parity unpinned). */
int hc_fill_eps_dev(hc_ctx *ctx, const float *z, double *out, int64_t ny, int64_t nx, float nodata, int seed_mode,
                    double eps, int64_t *sweeps, void *stream);

/* D8 direction (uint8 code, HC_DIR_NODATA on nodata / TauDEM-contaminated cells) and optional
 * slope (float32, may be NULL).  Replaces wbt.d8_pointer (run_aws.py:539) and the steepest-
 * descent half of TauDEM D8FlowDir -p -sd8 (run_aws.py:594-603). */
int hc_d8_dev(hc_ctx *ctx, const float *z, uint8_t *dir, float *slope, int64_t ny, int64_t nx,
              float nodata, double dx, double dy, int table, void *stream);

/* Flat resolution (Barnes et al. 2014 two-gradient mask; the job TauDEM D8FlowDir does on
 * filled pits, run_aws.py:594-603).  Rewrites HC_DIR_NOFLOW cells of drainable flats in place. */
int hc_resolve_flats_dev(hc_ctx *ctx, const float *z, uint8_t *dir, int64_t ny, int64_t nx,
                         float nodata, int table, void *stream);

/* D8 flow accumulation in cells, self-inclusive, uint32.  Replaces wbt.d8_flow_accumulation
 * (pntr=True, run_aws.py:544) and TauDEM AreaD8 -nc (run_aws.py:608-617). */
int hc_accum_dev(hc_ctx *ctx, const uint8_t *dir, uint32_t *acc, int64_t ny, int64_t nx,
                 int table, void *stream);

/* The north-star chain fill -> D8(+slope) -> flats -> accumulation on one raster.
 * filled / dir / acc are required, slope may be NULL; resolve_flats is a flag. */
int hc_fill_d8_accum_dev(hc_ctx *ctx, const float *z, float *filled, uint8_t *dir, float *slope,
                         uint32_t *acc, int64_t ny, int64_t nx, float nodata, double dx, double dy,
                         int seed_mode, int table, int resolve_flats, void *stream);

/* Pitched ("leading dimension") form of hc_fill_d8_accum_dev: every raster has ld >= nx elements per row.  With ld a
 * multiple of 16 and 16-byte aligned bases every row is vector- and TMA-addressable whatever nx is; the library
 * overwrites the padding columns [nx, ld) of z with nodata (hence non-const) and runs the chain on the ny x ld raster,
 * which gives the dense call's result in the first nx columns (a nodata column acts like the raster edge for seeds,
 * contamination and counts).  Output padding columns are unspecified.  Same reference call sites as
 * hc_fill_d8_accum_dev (run_aws.py:580-617). */
int hc_fill_d8_accum_ld_dev(hc_ctx *ctx, float *z, float *filled, uint8_t *dir, float *slope, uint32_t *acc,
                            int64_t ny, int64_t nx, int64_t ld, float nodata, double dx, double dy, int seed_mode,
                            int table, int resolve_flats, void *stream);

/* Connected-component labels of mask != 0, conn = 4 or 8, ids 1..n in raster-scan order of each
 * component's first cell (exactly scipy.ndimage.label, tools/depressions.py:1646-1669,
 * tools/endo.py:323).  *nlab receives n (host pointer, may be NULL). */
int hc_label_dev(hc_ctx *ctx, const uint8_t *mask, int32_t *lab, int64_t ny, int64_t nx, int conn,
                 int32_t *nlab, void *stream);

/* WhiteboxTools' single-cell-pit tools (models/depression_handling.py:259-261,
 * models/find_sinks.py:118).  fill: an interior cell whose 8 neighbours are all valid and strictly
 * higher is raised to the lowest of them.  breach: for such a pit (two cells from the edge at
 * least), every lower 2-ring cell cuts the bridging 1-ring cell to the mean of pit and target
 * (serial last-write-wins order reproduced).  z and out may not alias. */
int hc_fill_single_cell_pits_dev(hc_ctx *ctx, const float *z, float *out, int64_t ny, int64_t nx,
                                 float nodata, void *stream);
int hc_breach_single_cell_pits_dev(hc_ctx *ctx, const float *z, float *out, int64_t ny, int64_t nx,
                                   float nodata, void *stream);

/* TaudemCheck (models/taudem_check.py:143-147): mask = D8 is nodata (255) on a valid cell whose
 * flatten mask is in 2..5 and whose EDM value is < 4; *n_problem (host pointer) gets the count.
 * hc_taudem_fix (fix_basin, :183-194): cells with flat > 0 are raised by `bump` (0.10 m). */
int hc_taudem_check_dev(hc_ctx *ctx, const uint8_t *p, const float *fel, const uint8_t *flat,
                        const uint8_t *edm, uint8_t *mask, int64_t *n_problem, int64_t ny, int64_t nx,
                        float nodata, void *stream);
int hc_taudem_fix_dev(hc_ctx *ctx, const float *dem, const uint8_t *flat, float *out, int64_t ny,
                      int64_t nx, float nodata, float bump, void *stream);
/* np.round(a - b, decimals) on float32 rasters, numpy's float32 rule rint(x * 10^d) / 10^d (the fill / carve
 * difference maps of tools/depressions.py:1308, :1395); out may alias a or b. */
int hc_diff_round_dev(hc_ctx *ctx, const float *a, const float *b, float *out, int64_t n, int decimals, void *stream);

/* Smoothing stencils of CreateDTM (models/dem_noise_removal.py:397-487) and dep.create_clump_array.
 * conv2d_symm = scipy.signal.convolve2d(boundary='symm', mode='same') as pdt.apply_convolve_filter
 * calls it (tools/derive.py:94); kernel is a HOST pointer to kh*kw float32 (odd sizes <= 31).
 * slope: pdt.slope_D2 (d4=0) / slope_D4 (d4=1) -> int8 degrees (deg) and/or float32 magnitude (mag).
 * curvature_d2: pdt.curvature_D2.  mean_std: np.mean/np.std in float64.  dtm_select: the filter
 * selection block (:414-471); masks may be NULL; cbreak0/1 = float32(curvebreak * sigma). */
int hc_conv2d_symm_dev(hc_ctx *ctx, const float *in, const float *kernel_host, int kh, int kw, float *out,
                       int64_t ny, int64_t nx, void *stream);
/* the four DW smoothers of CreateDTM (models/dem_noise_removal.py:399-402) in one pass over the raster:
 * kernels_host = the 9x9, 7x7, 5x5, 3x3 float32 kernels back to back (164 floats, host pointer) */
int hc_conv2d_symm_dw4_dev(hc_ctx *ctx, const float *in, const float *kernels_host, float *out9, float *out7,
                           float *out5, float *out3, int64_t ny, int64_t nx, void *stream);
int hc_slope_dev(hc_ctx *ctx, const float *z, int8_t *deg, float *mag, int64_t ny, int64_t nx, double dx,
                 double dy, int d4, void *stream);
/* Optional: the 90 float32 slopes at which np.rad2deg(np.arctan(x)).astype(int8) steps to 1, 2, ..., 90 degrees, as the
 * caller's own arctan places them (pydrodem_b200.stencils.degree_thresholds derives them from numpy by bisection).
 * With them hc_slope_dev's degree output is a binary search and bit-identical to that numpy; n = 0 restores atanf.
 * (tools/derive.py:128,203: `.astype(np.int8)` of the degree array.) */
int hc_slope_thresholds_set(hc_ctx *ctx, const float *thr, int32_t n);
int hc_curvature_d2_dev(hc_ctx *ctx, const float *z, float *out, int64_t ny, int64_t nx, double dx, double dy,
                        int geometric, void *stream);
int hc_mean_std_dev(hc_ctx *ctx, const float *x, int64_t n, double *mean, double *std, void *stream);
/* partial sums of a row band for the global sigma: sums[0] = sum(x - shift), sums[1] = sum((x - shift)^2) (host) */
int hc_sum_shifted_dev(hc_ctx *ctx, const float *x, int64_t n, double shift, double *sums, void *stream);
/* TIFF LZW codec (host; TIFF 6.0 flavour with libtiff's early change) for the GeoTIFF wire format the reference uses
 * (COMPRESS=LZW, tools/convert.py:202,312).  decode: *nout <= ndst bytes written.  encode: cap >= 2 * nsrc + 16. */
int hc_lzw_decode_host(const uint8_t *src, int64_t nsrc, uint8_t *dst, int64_t ndst, int64_t *nout);
int hc_lzw_encode_host(const uint8_t *src, int64_t nsrc, uint8_t *dst, int64_t cap, int64_t *nout);
/* Depression breaching, Lindsay (2016): wbt.breach_depressions(dem, out, max_length=radius, max_depth=None,
 * fill_pits=True, flat_increment=0.0) as dep.breach_pits calls it (tools/depressions.py:1376-1380, :1012-1014).
 * HOST arithmetic (one global priority order, not a bandwidth-bound pass; no context needed):
 * z/out are host rasters (out != z).  max_length <= 0 and max_depth <= 0 / inf mean unconstrained.  do_fill = 1 also
 * fills what could not be breached (host priority flood); with do_fill = 0 the caller runs hc_fill_* (seed mode WBT)
 * on the result.  flat_increment is 0 (level channels).  stats (5 x int64, nullable): pits met, breached, left to the
 * fill, cells lowered, cells raised.  Specification and its PARITY-UNPINNED status: csrc/breach.cu. */
int hc_breach_depressions_host(const float *z, float *out, int64_t ny, int64_t nx, float nodata, int64_t max_length,
                               double max_depth, int fill_pits, int do_fill, int64_t *stats);
/* EXPERIMENTAL -- the pointer-chasing half of the order-independent flow-path breach (specification
 * oracle/breach_flowpath.py): w = surface to breach, filled = hc_fill(w, WBT seeds), dir = hc_d8 + hc_resolve_flats
 * (WBT table) of `filled` with every seed (hc_wbt_seed_mask_host) lowered by one float32 ulp, so that flats at a
 * seed's level drain through it (all host rasters here); out = cut surface, to be filled afterwards.  stats (4 x
 * int64, nullable): pits, breached, left, cells lowered.  Not yet what wbt.breach_depressions runs. */
int hc_breach_flowpath_host(const float *w, const float *filled, const uint8_t *dir, float *out, int64_t ny, int64_t nx,
                            float nodata, int64_t max_length, double max_depth, int64_t *stats);
int hc_wbt_seed_mask_host(const float *w, uint8_t *seed_out, int64_t ny, int64_t nx, float nodata);
/* The whole flow-path breach on the device (the default of wbt.breach_depressions; GPU-checked against its specification): seeds, fill, seeds lowered by one
 * ulp, D8 + flats, one walk per pit (atomic min on the cut surface).  w -> out (fill it afterwards); filled / dir / seed
 * are caller-provided device rasters that receive the intermediates; dx, dy as for hc_d8_dev.  stats as above (host). */
int hc_breach_flowpath_dev(hc_ctx *ctx, const float *w, float *out, float *filled, uint8_t *dir, uint8_t *seed,
                           int64_t ny, int64_t nx, float nodata, double dx, double dy, int64_t max_length,
                           double max_depth, int64_t *stats, void *stream);
/* Least-cost breaching on the device (csrc/leastcost.cu): replaces WhiteboxTools breach_depressions_least_cost as
 * tools/depressions.py:1370-1374 (breach_pits, method 'LC') and :996-1000 (combined_solution, carve_method 'LC') call
 * it -- wbt.breach_depressions_least_cost(dem, output, dist, max_cost=, min_dist=, flat_increment=, fill=False).
 * Order-independent restatement of Lindsay & Dhun (2015), specification oracle/breach_leastcost.py (parity unpinned:
 * the tool itself is not in the reference tree).  w -> out = the cut surface (fill it afterwards with hc_fill_dev for
 * fill=True / the caller's closing Wang & Liu fill).  dist 1..65 cells; max_cost <= 0: no limit; min_dist != 0: shortest
 * breach first, then cheapest; flat_increment <= 0: level channels.  stats (4 x int64, host, nullable): pits, breached,
 * left, cells lowered. */
int hc_breach_least_cost_dev(hc_ctx *ctx, const float *w, float *out, int64_t ny, int64_t nx, float nodata, int64_t dist,
                             double max_cost, int min_dist, double flat_increment, int64_t *stats, void *stream);
/* Water stencils of tools/water.py (called from models/depression_handling.py:647,889,898 and
 * models/burn_rivers.py:841,856).  wm = water_mask_array as uint8, swo = SWO_array as int16, dem/burn float32;
 * `burn` (DEM_burn_arr) is updated in place and may alias `dem`, as in the reference's own calls.
 * minmax_filter: scipy.ndimage.minimum_filter / maximum_filter(in, size=size) (mode 'reflect'), elem_bytes 1 (uint8)
 *   or 4 (float32).  conv2d_symm_f64: apply_convolve_filter with a float64 kernel (host pointer, odd sizes <= 15):
 *   float64 accumulation, float32 result.
 * river_minimize :786-847 (out may not alias dem); raise_small_rivers :1010-1045; raise_small_streams :960-1007;
 * smooth_rivers :928-957.  The scalar parameters are the reference's keyword arguments, narrowed to float32 the
 * way numpy narrows a Python float against a float32 array. */
int hc_minmax_filter_dev(hc_ctx *ctx, const void *in, void *out, int64_t ny, int64_t nx, int size, int is_max,
                         int elem_bytes, void *stream);
int hc_conv2d_symm_f64_dev(hc_ctx *ctx, const float *in, const double *kernel_host, int kh, int kw, float *out,
                           int64_t ny, int64_t nx, void *stream);
int hc_river_minimize_dev(hc_ctx *ctx, const float *dem, const uint8_t *wm, const int16_t *swo, float nodata,
                          float max_burn, int filter_size, float *out, int64_t ny, int64_t nx, void *stream);
int hc_raise_small_rivers_dev(hc_ctx *ctx, const float *dem, float *burn, const uint8_t *wm, const int16_t *swo,
                              float nodata, int swo_high, int swo_low, float extra_raise, float max_raise,
                              int filter_size, int64_t ny, int64_t nx, void *stream);
int hc_raise_small_streams_dev(hc_ctx *ctx, const float *dem, float *burn, const uint8_t *wm, float nodata,
                               float extra_raise, float max_raise, int fsize, int64_t ny, int64_t nx, void *stream);
int hc_smooth_rivers_dev(hc_ctx *ctx, const float *dem, float *burn, const uint8_t *wm, const int16_t *swo,
                         float nodata, int swo_cut, float max_lower, int filter_size, int64_t ny, int64_t nx,
                         void *stream);
/* wbt.fill_missing_data(i, output, filter, weight) (tools/water.py:1396,1774; models/burn_rivers.py:1283,1582;
 * models/dem_noise_removal.py:739): a nodata cell takes the inverse-distance-weighted (1 / d^weight) mean of the valid
 * cells that border nodata within filter / 2 cells; stays nodata when none is in reach.  out != z.  Restated from the
 * tool's documentation (WhiteboxTools unavailable: parity unpinned). */
int hc_fill_missing_data_dev(hc_ctx *ctx, const float *z, float *out, int64_t ny, int64_t nx, float nodata, int filter,
                             double weight, void *stream);
/* Raster halves of wtr.enforce_monotonic_rivers (tools/water.py:1407-1779) for one water level h.  hc_mono_mask_dev: the
 * candidate cells (:1488-1496): river classes 4..7 with left != 0, above h or 8-adjacent to a river cell above h, not
 * exactly on h.  hc_mono_stats_dev: per label 1..nlab of the 4-connected clumps of that mask, stats[6 * id + k] =
 * {cells (with inner == 1 when inner is given), touches donut == 1, row min, row max, col min, col max} (:1505-1538);
 * the caller presets stats to {0, 0, INT_MAX, -1, INT_MAX, -1}.  donut / inner may be NULL. */
int hc_mono_mask_dev(hc_ctx *ctx, const float *dem, const uint8_t *wm, const uint8_t *left, float h, uint8_t *out,
                     int64_t ny, int64_t nx, void *stream);
int hc_mono_stats_dev(hc_ctx *ctx, const int32_t *lab, const uint8_t *donut, const uint8_t *inner, int32_t nlab,
                      int32_t *stats, int64_t ny, int64_t nx, void *stream);
/* Vegetation-bias lookup of CreateDTM.compute_veg_bias_3D (models/dem_noise_removal.py:764-805; the 2-D table of
 * compute_veg_bias_2D :620-629 is n1 == 1 with i1 NULL).
 * box3_index: inputs first capped at in_max (:777) and zeroed where in_zero != 0 (nullable; the ocean mask, :687);
 *   3x3 symmetric-border mean in float64, narrowed to float32, values > zero_above set to 0, astype(int)
 *   (:766-772, :778-781); `smooth` (nullable) receives the float32 array (self.TC_array).  Pass INFINITY to
 *   disable in_max / zero_above.
 * lut3d_gather: bias = cube[i0, i1, min(i2, cap2)] (float64, C order n0 x n1 x n2); i2 is capped in place (:792);
 *   bias is 0 where zero_mask != 0 (nullable; :361); dem_out = dem - bias in float64 when dem/dem_out are given
 *   (:364); bias or dem_out may be NULL.  An index outside the table is an error (*oob_host = 1, HC_ERR_ARG). */
int hc_box3_index_dev(hc_ctx *ctx, const float *in, const uint8_t *in_zero, float in_max, int64_t ny, int64_t nx,
                      float zero_above, int32_t *idx, float *smooth, void *stream);
int hc_lut3d_gather_dev(hc_ctx *ctx, const int32_t *i0, const int32_t *i1, int8_t *i2, int cap2, const double *cube,
                        int n0, int n1, int n2, const uint8_t *zero_mask, double *bias, const float *dem,
                        double *dem_out, int64_t n, int *oob_host, void *stream);
int hc_dtm_select_dev(hc_ctx *ctx, const float *dem, const float *z9, const float *z7, const float *z5,
                      const float *z3, const int8_t *slope0, const float *curv, const uint8_t *swo,
                      const uint8_t *edm, const uint8_t *tx, const uint8_t *artifact, float *zf, int8_t *filt,
                      int64_t n, float cbreak0, float cbreak1, const int *slopebreaks5, void *stream);
/* The same block fused (models/dem_noise_removal.py:397-471): two tile passes over the DEM, no smoothed raster in HBM.
 * hc_dtm_pre_dev: curvature_D2 (geometric, float32) and slope_D4 / slope_D2 (int8 degrees) of the 5x5-smoothed DEM
 * (:405-410) plus sums[0] = sum, sums[1] = sum of squares (float64, HOST) of the curvature over rows [row_lo, row_hi)
 * -- np.std(curvature) of :413; curv / slope0 may be NULL (sums only).  hc_dtm_apply_dev: the four DW smoothings
 * (:399-402) and the selection (:414-471) -> Zf and the filter code.  kernels_host = the 9x9, 7x7, 5x5 and 3x3 kernels
 * back to back (81 + 49 + 25 + 9 float32, HOST pointer); cbreak0/1 = c * sigma as np.float32.  Results are bit-identical
 * to the unfused calls above. */
int hc_dtm_pre_dev(hc_ctx *ctx, const float *dem, const float *kernels_host, float *curv, int8_t *slope0, int64_t ny,
                   int64_t nx, double dx, double dy, int d4, int64_t row_lo, int64_t row_hi, double *sums, void *stream);
int hc_dtm_apply_dev(hc_ctx *ctx, const float *dem, const float *kernels_host, const float *curv, const int8_t *slope0,
                     const uint8_t *swo, const uint8_t *edm, const uint8_t *tx, const uint8_t *artifact, float *zf,
                     int8_t *filt, int64_t ny, int64_t nx, float cbreak0, float cbreak1, const int *slopebreaks5,
                     void *stream);

/* ---- Row-band sharding (BASELINE.json configs[2]: one 40000^2 DEM over 2/4/8 GPUs) ----
 * A band owns rows [0, nb) of its slice; its rasters carry ONE halo row above (row -1) when
 * flags & 1 (there is a band above) and below (row nb) when flags & 2.  Pointers passed here
 * point at own row 0.  Every pass is split into a band-local phase, an exchange the HOST performs
 * (torch.distributed all_gather over NCCL in pydrodem_b200/bands.py — the library never talks to
 * another rank) and a finish phase every rank runs on the gathered seam records.  Results are
 * bit-identical to the whole-raster entry points above.  The equivalent in the reference is the
 * MPI decomposition inside TauDEM (`mpiexec -n N PitRemove/D8FlowDir/AreaD8`, run_aws.py:584-612).
 *
 * fill   local : watershed labels with every seam cell a terminal of its own; publishes <= cap
 *                (a, b, float-bits w) edges in global node ids (0 = outside,
 *                1 + (band*2 + side)*nx + col = seam cell): the minimum spanning forest of the
 *                band's terminal graph + the cross-seam edges of its bottom seam; unused entries 0.
 *        finish: all_edges = nbands * cap triples (band order); solves the seam graph, writes out.
 *        HC_SEED_WBT: `ext` = uint8 raster (halo rows like z) that is 1 on nodata cells connected to the raster
 *        frame through nodata — a band cannot know that by itself; bands.py derives it from the banded labelling
 *        of the nodata mask.  ext == NULL means every nodata cell is exterior.  Ignored for HC_SEED_TAUDEM. */
int64_t hc_band_fill_edge_cap(int64_t nx);
int hc_band_fill_local_dev(hc_ctx *ctx, const float *z, int64_t nb, int64_t nx, float nodata, int seed_mode,
                           int flags, int64_t band_index, const uint8_t *ext, uint32_t *edges_out, void *stream);
int hc_band_fill_finish_dev(hc_ctx *ctx, const uint32_t *all_edges, int64_t nbands, float *out, void *stream);
/* D8 on a band (z with halo rows; seam rows are not raster border). */
int hc_band_d8_dev(hc_ctx *ctx, const float *z, uint8_t *dir, float *slope, int64_t nb, int64_t nx,
                   float nodata, double dx, double dy, int table, int flags, void *stream);
/* flats: begin (fast tier + slow-tier init; z and dirA carry halo rows), then repeat
 * { relax; get_seams -> exchange -> put_halos } until no band reports `changed`, then finish.
 * seam record: uint32 [2 sides][3 fields nf,dlo,dhi][nx]; put_halos takes the upper neighbour's
 * side-1 record as `top` and the lower neighbour's side-0 record as `bot` (NULL where none). */
int hc_band_flats_begin_dev(hc_ctx *ctx, const float *z, const uint8_t *dirA, uint8_t *dirB, int64_t nb,
                            int64_t nx, float nodata, int table, int flags, void *stream);
int hc_band_flats_get_seams_dev(hc_ctx *ctx, uint32_t *out, void *stream);
int hc_band_flats_put_halos_dev(hc_ctx *ctx, const uint32_t *top, const uint32_t *bot, int *changed, void *stream);
/* Same, without any read-back: `flag_dev` (device int32) is set to 1 when a halo row changed (never cleared here), and the
 * seam tile rows are re-queued on the device.  The caller all-reduces the ranks' flag words once per exchange round. */
int hc_band_flats_put_halos_flag_dev(hc_ctx *ctx, const uint32_t *top, const uint32_t *bot, int32_t *flag_dev, void *stream);
int hc_band_flats_relax_dev(hc_ctx *ctx, void *stream);
int hc_band_flats_finish_dev(hc_ctx *ctx, void *stream);
/* accumulation: local writes the band-local counts into acc and publishes uint32
 * [2 sides][2 fields exit_of,bout][nx]; finish takes all bands' records (band order), solves the seam
 * forest and completes the SAME acc raster. dir carries halo rows. */
int hc_band_accum_local_dev(hc_ctx *ctx, const uint8_t *dir, uint32_t *acc, int64_t nb, int64_t nx, int table,
                            int flags, uint32_t *seam_out, void *stream);
int hc_band_accum_finish_dev(hc_ctx *ctx, const uint32_t *all_seams, int64_t nbands, int64_t band_index,
                             uint32_t *acc, void *stream);
/* labels of ONE mask over row bands (scipy numbering over the whole raster): local union-find, then rounds of
 * { all_gather the uint32 [2][nx] seam records; relax } until no band reports `changed`, rank (returns the number of
 * components whose first cell lies in this band), all_gather the counts (offset = sum over earlier bands), publish the
 * (root, id) pairs of owned seam components [2][nx][2], all_gather + sort them by root (host side), finish. */
int hc_band_label_local_dev(hc_ctx *ctx, const uint8_t *mask, int64_t nb, int64_t nx, int conn, int64_t row0,
                            uint32_t *seam_out, void *stream);
int hc_band_label_relax_dev(hc_ctx *ctx, const uint32_t *all_seams, int64_t nbands, int64_t band_index,
                            uint32_t *seam_out, int *changed, void *stream);
int hc_band_label_rank_dev(hc_ctx *ctx, int64_t *n_owned, void *stream);
int hc_band_label_pairs_dev(hc_ctx *ctx, int64_t offset, uint32_t *pairs_out, void *stream);
int hc_band_label_finish_dev(hc_ctx *ctx, int64_t offset, const uint32_t *keys, const uint32_t *vals, int64_t nkeys,
                             int32_t *lab, void *stream);
/* out[i] = flag[lab[i]] (flag[0] = 0): turns per-component flags into a raster (exterior nodata of the WBT seed rule) */
int hc_gather_flag_dev(hc_ctx *ctx, const int32_t *lab, const uint8_t *flag, uint8_t *out, int64_t n, void *stream);
/* diagnostics of the last banded fill: terminal-graph edges emitted, MSF rounds, seam-graph rounds */
int hc_band_stats(const hc_ctx *ctx, int64_t *term_edges, int32_t *msf_rounds, int32_t *global_rounds);

/* ---- Pit catalog / solution selection (tools/depressions.py:272-810, 1140-1228, 1471-1531) ----
 * The reference's per-pit Python loops as segmented reductions keyed by the label rasters.  All
 * pointers are DEVICE pointers; per-pit tables have npits+1 entries (entry 0 unused), per-carve tables
 * ncarves+1.  Sums are accumulated in float64 and returned as float32 (the reference sums float32
 * pairwise): tolerance 1e-5 relative on volfill / volcarve / volfillbehind, everything else exact. */
/* generic: max / min / float64 sum / count of vals (may be NULL) per label in [lo_id, nseg], restricted
 * to cells with sel[i] != 0 when sel is given; outputs may be NULL */
int hc_seg_stats_dev(hc_ctx *ctx, const int32_t *ids, const float *vals, const int32_t *sel, int64_t n,
                     int32_t nseg, int32_t lo_id, float *vmax, float *vmin, double *vsum, int64_t *count,
                     void *stream);
/* dep.build_pit_catalog (:272-509).  carve_bound receives carveID restricted to valid DEM cells (:330);
 * pairs_out the distinct (pit << 32 | carve) pairs sharing a cell, pair_included the reference's test
 * "carve's highest cell inside the pit OR lowest cell outside" (:435-439); *npairs is a host pointer.
 * carveout_id = carveID at the first nodata cell (:335), ignored when carveout != 0. */
int hc_pit_catalog_dev(hc_ctx *ctx, const float *dem, float nodata, const float *carve, const float *diff_carve,
                       const float *diff_fill, const int32_t *fillID, const int32_t *carveID,
                       const int32_t *carvefillID, int64_t n, int32_t npits, int32_t ncarves, int carveout,
                       int32_t carveout_id, int32_t *carve_bound, float *maxfill, float *volfill, int64_t *nfill,
                       float *maxcarve, float *volcarve, int64_t *ncarvecells, float *volfb, float *maxfb,
                       int8_t *complete, uint8_t *pit_out, int64_t *carve_amax, int64_t *carve_amin,
                       float *carve_min, double *carve_sum, int64_t *carve_n, int64_t *pairs_out, int64_t pair_cap,
                       uint8_t *pair_included, int64_t *npairs, void *stream);
/* The catalog over ROW BANDS of one raster (SURVEY.md 8e "Catalog / select / apply: AllReduce of per-label partial
 * max / sum"; pydrodem_b200/bands.py pit_catalog_banded).  Label rasters hold whole-raster ids (hc_band_label_*).
 * (1) hc_band_catalog_local_dev on every band: cell0 = whole-raster index of the band's first cell; partial tables out
 *     (per pit: maxfill, volfill64, nfill, fb_max, fb_sum, fb_n; per carve: carve_min, carve_sum, carve_n and the
 *     arg-extreme keys kmax / kmin, which carry whole-raster cell indices) + the band's distinct (pit, carve) pairs.
 *     The caller reduces across the bands: max (maxfill, fb_max, kmax as UNSIGNED), min (carve_min, kmin as unsigned),
 *     sum (the float64 sums, the counts); pair lists are concatenated and made unique.
 * (2) hc_band_catalog_owner_ids_dev on every band with the reduced keys: the pit id at each carve's first highest /
 *     first lowest cell where this band owns that cell (0 elsewhere, -1 for an empty carve): max-reduce.
 * (3) hc_catalog_finish_dev on the reduced tables (every rank, redundantly): the catalog's columns, pair_included,
 *     pit_out.  hc_pit_catalog_dev is exactly (1)-(3) on one raster. */
int hc_band_catalog_local_dev(hc_ctx *ctx, const float *dem, float nodata, const float *carve, const float *diff_carve,
                              const float *diff_fill, const int32_t *fillID, const int32_t *carveID,
                              const int32_t *carvefillID, int64_t n, int64_t cell0, int32_t npits, int32_t ncarves,
                              int32_t *carve_bound, float *maxfill, double *volfill64, int64_t *nfill, float *fb_max,
                              double *fb_sum, int64_t *fb_n, float *carve_min, double *carve_sum, int64_t *carve_n,
                              uint64_t *kmax, uint64_t *kmin, int64_t *pairs_out, int64_t pair_cap, int64_t *npairs,
                              void *stream);
int hc_band_catalog_owner_ids_dev(hc_ctx *ctx, const int32_t *fillID, int64_t n, int64_t cell0, int32_t ncarves,
                                  const uint64_t *kmax, const uint64_t *kmin, int32_t *fid_amax, int32_t *fid_amin,
                                  void *stream);
int hc_catalog_finish_dev(hc_ctx *ctx, int32_t npits, int32_t ncarves, const int64_t *pairs, int64_t npairs,
                          const int32_t *fid_amax, const int32_t *fid_amin, int carveout, int32_t carveout_id,
                          const double *volfill64, const float *fb_max, const double *fb_sum, const int64_t *fb_n,
                          const float *carve_min, const double *carve_sum, const int64_t *carve_n, float *volfill,
                          float *maxcarve, float *volcarve, int64_t *ncarvecells, float *volfb, float *maxfb,
                          int8_t *complete, uint8_t *pit_out, uint8_t *pair_included, void *stream);
/* select_solution over row bands: per-band partial tables (min-reduce wall_min, max-reduce sink_max, sum carve_water),
 * then the decision tree on the reduced tables.  wall / sink (and then wall_min / sink_max) may be NULL. */
int hc_band_select_partials_dev(hc_ctx *ctx, const int32_t *fillID, const int32_t *carve_bound, int64_t n, int32_t npits,
                                int32_t ncarves, const float *wall, const float *water, const float *sink, float *wall_min,
                                float *sink_max, int64_t *carve_water, void *stream);
int hc_pit_select_tables_dev(hc_ctx *ctx, int32_t npits, const float *maxfill, const float *volfill, const float *maxcarve,
                             const float *volcarve, const float *volfb, const float *maxfb, const int8_t *complete,
                             const int64_t *nfill, const int64_t *ncarvecells, const uint8_t *pit_out, const int64_t *pairs,
                             const uint8_t *pair_included, int64_t npairs, const float *wall_min, const float *sink_max,
                             const int64_t *carve_water, double fill_vol_thresh, double carve_vol_thresh,
                             double max_carve_depth, double maxcarve_factor, int apply_shallow_fill, int float32_scalars,
                             int32_t *solution, void *stream);
/* dep.select_solution (:517-810): decision tree per pit -> 4-digit codes.  wall / sink may be NULL; water
 * must not be (the reference raises NameError without a water mask, :606-647).  float32_scalars selects
 * the arithmetic numpy gives the reference's scalar expressions: 1 = float32 (numpy >= 2, weak Python
 * scalars; what the golden vectors were produced with), 0 = float64 (numpy < 2 value-based promotion). */
int hc_pit_select_dev(hc_ctx *ctx, const int32_t *fillID, const int32_t *carve_bound, int64_t n, int32_t npits,
                      int32_t ncarves, const float *maxfill, const float *volfill, const float *maxcarve,
                      const float *volcarve, const float *volfb, const float *maxfb, const int8_t *complete,
                      const int64_t *nfill, const int64_t *ncarvecells, const uint8_t *pit_out, const int64_t *pairs,
                      const uint8_t *pair_included, int64_t npairs, const float *wall, const float *water,
                      const float *sink, double fill_vol_thresh, double carve_vol_thresh, double max_carve_depth,
                      double maxcarve_factor, int apply_shallow_fill, int float32_scalars, int32_t *solution,
                      void *stream);
/* dep.apply_solution (:1140-1228): scatter fill / carve values by pit solution (later pits win, as in the
 * reference's loop order) and the int16 solution mask (in/out).  excluded: u8[npits+1] or NULL. */
int hc_pit_apply_dev(hc_ctx *ctx, const float *dem, const float *fillv, const float *carvev, const int32_t *fillID,
                     const int32_t *carve_bound, const int32_t *carvefillID, int64_t n, int32_t npits,
                     int32_t ncarves, const int32_t *solution, const uint8_t *excluded, const uint8_t *pit_out,
                     const int64_t *ncarvecells, const int64_t *pairs, const uint8_t *pair_included, int64_t npairs,
                     float *out, int16_t *mask, void *stream);
/* dep.fill_shallow_pits (:1471-1531): mod (in/out) takes the filled value wherever the cell's label
 * (0 = background included) has no cell deeper than sfill and no water-mask cell (water may be NULL). */
int hc_fill_shallow_pits_dev(hc_ctx *ctx, float *mod, float sfill, const int32_t *fillID, const float *diff_fill,
                             const float *fillv, const float *water, int64_t n, int32_t npits, int64_t *n_shallow,
                             void *stream);

/* k-th smallest element (0-based) of x[0..n), exact (radix select; no NaN in x).  The order statistics behind
 * np.percentile(raster_array, 1) in endo.find_sink (tools/endo.py:286).  value is a host pointer. */
int hc_kth_value_dev(hc_ctx *ctx, const float *x, int64_t n, int64_t k, float *value, void *stream);

/* HOST-buffer variants (H2D + kernels + D2H inside the call; synchronous). */
int hc_fill_host(hc_ctx *ctx, const float *z, float *out, int64_t ny, int64_t nx, float nodata,
                 int seed_mode);
int hc_fill_d8_accum_host(hc_ctx *ctx, const float *z, float *filled, uint8_t *dir, float *slope,
                          uint32_t *acc, int64_t ny, int64_t nx, float nodata, double dx, double dy,
                          int seed_mode, int table, int resolve_flats);
int hc_label_host(hc_ctx *ctx, const uint8_t *mask, int32_t *lab, int64_t ny, int64_t nx, int conn,
                  int32_t *nlab);
/* The same chain as a begin / end pair, for a stream of units (the reference processes a region as many tiles /
 * basins, run_local.py:636-677): `begin` uploads, computes and queues the result copies of one unit into PINNED
 * host buffers; `end` waits for them.  slot = 0 or 1 selects one of two device staging sets, so begin(1) may be
 * called before end(0): one unit's results travel over PCIe while the next is uploaded and computed.  The host
 * buffers of a slot must not be touched between its begin and end. */
int hc_fill_d8_accum_host_begin(hc_ctx *ctx, int slot, const float *z, float *filled, uint8_t *dir, float *slope,
                                uint32_t *acc, int64_t ny, int64_t nx, float nodata, double dx, double dy,
                                int seed_mode, int table, int resolve_flats);
int hc_fill_d8_accum_host_end(hc_ctx *ctx, int slot);

/* Per-kernel timing with CUDA events on the launching stream (what bench.py's roofline line is
 * computed from).  enable(on=1) starts logging every launch; read() synchronises, writes one line
 * "kernel_name launches total_ms max_ms" per kernel into buf and clears the log. */
int hc_profile_enable(hc_ctx *ctx, int on, int max_records);
int hc_profile_read(hc_ctx *ctx, char *buf, int64_t buflen);
/* 16 internal work counters (sweeps, tile visits ...) since the last read; diagnostics only. */
int hc_debug_read(hc_ctx *ctx, unsigned long long *out16);

/* Diagnostics of the last fill on this context: basins found, Boruvka rounds run. */
int hc_fill_stats(const hc_ctx *ctx, int64_t *n_basins, int32_t *n_rounds);

#ifdef __cplusplus
}
#endif
#endif /* HYDROCORE_H */
