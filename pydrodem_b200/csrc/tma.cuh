// tma.cuh — Tensor Memory Accelerator staging of raster tiles (sm_90+ cp.async.bulk.tensor, SASS UTMALDG).
//
// A tile kernel that stages a (tile + halo) window of a raster spends a large part of its instruction stream on
// the loader: per element an index split, four bounds tests, an address and a predicated load.  With a 2-D tensor
// map over the raster ONE elected thread issues one bulk-tensor copy per window; the copy engine does the address
// generation, fills everything outside the raster (negative coordinates included) with NaN (float maps) or zero
// (integer maps) and signals an mbarrier when the bytes have landed.  The kernels' window pitch is the box width.
//
// Requirements of the hardware path (hc_tmap_ok): raster base 16-byte aligned and row pitch a multiple of 16 bytes
// (float32 / uint32 rasters: nx % 4 == 0; byte rasters: nx % 16 == 0).  Rasters that do not qualify (the caller's
// dense 10801-wide arrays) take the kernels' scalar loaders; pydrodem_b200.core.empty_raster() hands out padded
// device rasters that always qualify.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

// host: encode a 2-D tiled map (innermost dimension first).  Returns 0 on success; on failure the caller falls back
// to the scalar loader (`ok` stays 0).
struct hc_tmap {
  CUtensorMap map;
  int ok;
};
int hc_tmap_2d(hc_tmap *out, const void *base, int elem_bytes, int is_float, uint64_t width, uint64_t height,
               uint64_t pitch_bytes, uint32_t box_w, uint32_t box_h);
static inline bool hc_tmap_ok(const void *base, uint64_t pitch_bytes) {
  return (((uintptr_t)base) & 15u) == 0 && (pitch_bytes & 15u) == 0 && pitch_bytes > 0;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t hc_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void hc_mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(hc_smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void hc_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(hc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void hc_mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "HC_MBAR_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra HC_MBAR_DONE_%=;\n"
      "bra HC_MBAR_WAIT_%=;\n"
      "HC_MBAR_DONE_%=:\n"
      "}\n" ::"r"(hc_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// box (x, y) = coordinates of the box's first element (may be negative / past the end: filled).  x * element size
// MUST be a multiple of 16 bytes (measured on B200: a float box at x = -1 raises "illegal instruction", x = -8 or
// x = 100 work, tests/probes/tma_probe.cu) — halo windows therefore start 4 floats / 16 bytes left of the tile.
__device__ __forceinline__ void hc_tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          hc_smem_u32(smem_dst)),
      "l"(map), "r"(hc_smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void hc_tmap_prefetch(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
#endif
