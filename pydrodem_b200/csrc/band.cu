// band.cu — extern "C" surface of the row-band phases (include/hydrocore.h, "row-band sharding").
// The phases themselves live next to the whole-raster passes they extend (fill.cu, d8.cu, flats.cu,
// accum.cu); the exchanges between phases are the host's job (torch.distributed over NCCL, see
// pydrodem_b200/bands.py) — the library never talks to another rank.
#include "common.cuh"

#define CHECK_CTX(ctx) HC_ENTER(ctx)

static int check_band(int64_t nb, int64_t nx, int flags) {
  if (nb <= 0 || nx <= 0) { hc_set_error("band: empty shape"); return HC_ERR_ARG; }
  if ((double)(nb + 2) * (double)nx >= 2147483632.0) { hc_set_error("band exceeds the 31-bit cell index"); return HC_ERR_TOO_LARGE; }
  if (flags & ~3) { hc_set_error("band: bad flags"); return HC_ERR_ARG; }
  if ((flags & 3) && nb < 2) { hc_set_error("a band with a neighbour needs at least 2 rows"); return HC_ERR_ARG; }
  return HC_OK;
}

extern "C" int64_t hc_band_fill_edge_cap(int64_t nx) { return band_fill_edge_cap(nx); }

extern "C" int hc_band_fill_local_dev(hc_ctx *ctx, const float *z, int64_t nb, int64_t nx, float nodata, int seed_mode,
                                      int flags, int64_t band_index, const uint8_t *ext, uint32_t *edges_out, void *stream) {
  CHECK_CTX(ctx);
  HC_TRY(check_band(nb, nx, flags));
  if (!z || !edges_out || band_index < 0) { hc_set_error("hc_band_fill_local: bad argument"); return HC_ERR_ARG; }
  if (seed_mode != HC_SEED_WBT && seed_mode != HC_SEED_TAUDEM) { hc_set_error("hc_band_fill_local: bad seed_mode"); return HC_ERR_ARG; }
  return band_fill_local(ctx, z, nb, nx, nodata, seed_mode, flags, band_index, edges_out, (cudaStream_t)stream, ext);
}

extern "C" int hc_band_fill_finish_dev(hc_ctx *ctx, const uint32_t *all_edges, int64_t nbands, float *out, void *stream) {
  CHECK_CTX(ctx);
  if (!all_edges || !out || nbands < 1) { hc_set_error("hc_band_fill_finish: bad argument"); return HC_ERR_ARG; }
  return band_fill_finish(ctx, all_edges, nbands, out, (cudaStream_t)stream);
}

extern "C" int hc_band_d8_dev(hc_ctx *ctx, const float *z, uint8_t *dir, float *slope, int64_t nb, int64_t nx,
                              float nodata, double dx, double dy, int table, int flags, void *stream) {
  CHECK_CTX(ctx);
  HC_TRY(check_band(nb, nx, flags));
  if (!z || !dir) { hc_set_error("hc_band_d8: null pointer"); return HC_ERR_ARG; }
  return d8_run_rows(ctx, z, dir, slope, nb, nx, nodata, dx, dy, table, hc_make_rows(nb, flags), (cudaStream_t)stream);
}

extern "C" int hc_band_flats_begin_dev(hc_ctx *ctx, const float *z, const uint8_t *dirA, uint8_t *dirB, int64_t nb,
                                       int64_t nx, float nodata, int table, int flags, void *stream) {
  CHECK_CTX(ctx);
  HC_TRY(check_band(nb, nx, flags));
  if (!z || !dirA || !dirB) { hc_set_error("hc_band_flats_begin: null pointer"); return HC_ERR_ARG; }
  return band_flats_begin(ctx, z, dirA, dirB, nb, nx, nodata, table, flags, (cudaStream_t)stream);
}
extern "C" int hc_band_flats_get_seams_dev(hc_ctx *ctx, uint32_t *out, void *stream) {
  CHECK_CTX(ctx);
  if (!out) { hc_set_error("hc_band_flats_get_seams: null pointer"); return HC_ERR_ARG; }
  return band_flats_get_seams(ctx, out, (cudaStream_t)stream);
}
extern "C" int hc_band_flats_put_halos_dev(hc_ctx *ctx, const uint32_t *top, const uint32_t *bot, int *changed, void *stream) {
  CHECK_CTX(ctx);
  return band_flats_put_halos(ctx, top, bot, changed, (cudaStream_t)stream);
}
extern "C" int hc_band_flats_put_halos_flag_dev(hc_ctx *ctx, const uint32_t *top, const uint32_t *bot, int32_t *flag_dev, void *stream) {
  CHECK_CTX(ctx);
  if (!flag_dev) { hc_set_error("hc_band_flats_put_halos_flag: null flag"); return HC_ERR_ARG; }
  return band_flats_put_halos_flag(ctx, top, bot, (int *)flag_dev, (cudaStream_t)stream);
}
extern "C" int hc_band_flats_relax_dev(hc_ctx *ctx, void *stream) {
  CHECK_CTX(ctx);
  return band_flats_relax(ctx, (cudaStream_t)stream);
}
extern "C" int hc_band_flats_finish_dev(hc_ctx *ctx, void *stream) {
  CHECK_CTX(ctx);
  return band_flats_finish(ctx, (cudaStream_t)stream);
}

extern "C" int hc_band_accum_local_dev(hc_ctx *ctx, const uint8_t *dir, uint32_t *acc, int64_t nb, int64_t nx, int table,
                                       int flags, uint32_t *seam_out, void *stream) {
  CHECK_CTX(ctx);
  HC_TRY(check_band(nb, nx, flags));
  if (!dir || !acc || !seam_out) { hc_set_error("hc_band_accum_local: null pointer"); return HC_ERR_ARG; }
  return band_accum_local(ctx, dir, acc, nb, nx, table, flags, seam_out, (cudaStream_t)stream);
}
extern "C" int hc_band_accum_finish_dev(hc_ctx *ctx, const uint32_t *all_seams, int64_t nbands, int64_t band_index,
                                        uint32_t *acc, void *stream) {
  CHECK_CTX(ctx);
  if (!all_seams || !acc || nbands < 1 || band_index < 0 || band_index >= nbands) { hc_set_error("hc_band_accum_finish: bad argument"); return HC_ERR_ARG; }
  return band_accum_finish(ctx, all_seams, nbands, band_index, acc, (cudaStream_t)stream);
}

extern "C" int hc_band_stats(const hc_ctx *ctx, int64_t *term_edges, int32_t *msf_rounds, int32_t *global_rounds) {
  if (!ctx) return HC_ERR_ARG;
  if (term_edges) *term_edges = ctx->band_term_edges;
  if (msf_rounds) *msf_rounds = ctx->band_msf_rounds;
  if (global_rounds) *global_rounds = ctx->band_global_rounds;
  return HC_OK;
}
