// tiffcodec.cu -- host-side TIFF LZW codec for raster_io (the GeoTIFF wire format): the reference
// compresses its rasters with COMPRESS=LZW (tools/convert.py:202 tiled, :312 striped), so reading a pydrodem output
// and writing one it can read back needs this codec.  Plain host C++ (byte-serial by nature; strips / tiles are the
// unit of parallelism and raster_io hands them over one at a time).  TIFF 6.0 flavour: MSB-first codes, 9..12 bits,
// ClearCode 256, EOI 257, "early change" of the code width as libtiff writes it.
#include <cstdint>
#include <cstring>

#include "common.cuh"

namespace {
const int LZW_CLEAR = 256, LZW_EOI = 257, LZW_FIRST = 258, LZW_MAX = 4096;
}

// returns HC_OK and *nout = bytes produced (<= ndst); a truncated / corrupt stream is HC_ERR_ARG
extern "C" int hc_lzw_decode_host(const uint8_t *src, int64_t nsrc, uint8_t *dst, int64_t ndst, int64_t *nout) {
  if (nout) *nout = 0;
  if ((!src && nsrc > 0) || (!dst && ndst > 0) || nsrc < 0 || ndst < 0) {
    hc_set_error("hc_lzw_decode: bad argument");
    return HC_ERR_ARG;
  }
  // Every table string is a run of bytes that already sits in the output: entry `next` = string(old) + the byte that
  // followed it, i.e. dst[pos_old .. pos_old + len_old].  So the table is (offset, length) pairs into dst and a code
  // is expanded by a forward copy (which also handles the code == next case, where the last byte copied is the first
  // one written).
  static thread_local int64_t pos[LZW_MAX];
  static thread_local int32_t length[LZW_MAX];
  int next = LZW_FIRST, width = 9, old = -1;
  int64_t old_pos = 0;
  int32_t old_len = 0;
  uint64_t acc = 0;
  int nbits = 0;
  int64_t ip = 0, op = 0;
  while (op < ndst) {
    while (nbits < width && ip < nsrc) { acc = (acc << 8) | src[ip++]; nbits += 8; }
    if (nbits < width) break;                       // ran out of input without EOI: accept what was decoded
    int code = (int)((acc >> (nbits - width)) & ((1u << width) - 1));
    nbits -= width;
    if (code == LZW_EOI) break;
    if (code == LZW_CLEAR) { next = LZW_FIRST; width = 9; old = -1; continue; }
    if (old < 0) {
      if (code > 255) { hc_set_error("hc_lzw_decode: corrupt stream (first code after clear)"); return HC_ERR_ARG; }
      dst[op] = (uint8_t)code;
      old = code; old_pos = op; old_len = 1;
      op++;
      continue;
    }
    if (code > next || (code >= 256 && code < LZW_FIRST) || next >= LZW_MAX) {
      hc_set_error("hc_lzw_decode: corrupt stream (code %d, table %d)", code, next);
      return HC_ERR_ARG;
    }
    pos[next] = old_pos;
    length[next] = old_len + 1;
    int64_t start = op;
    int32_t len;
    if (code < 256) {
      dst[op++] = (uint8_t)code;
      len = 1;
    } else {
      len = length[code];
      const uint8_t *from = dst + pos[code];
      int64_t room = ndst - op;
      int32_t ncopy = len < room ? len : (int32_t)room;
      for (int32_t t = 0; t < ncopy; t++) dst[op + t] = from[t];
      op += ncopy;
    }
    next++;
    if (next >= (1 << width) - 1 && width < 12) width++;
    old = code; old_pos = start; old_len = len;
  }
  if (nout) *nout = op < ndst ? op : ndst;
  return HC_OK;
}

// worst case output: every byte a 12-bit code + clears: cap >= nsrc * 2 + 16 is always enough
extern "C" int hc_lzw_encode_host(const uint8_t *src, int64_t nsrc, uint8_t *dst, int64_t cap, int64_t *nout) {
  if (nout) *nout = 0;
  if ((!src && nsrc > 0) || !dst || nsrc < 0 || cap < nsrc * 2 + 16) {
    hc_set_error("hc_lzw_encode: bad argument (cap must be >= 2 * nsrc + 16)");
    return HC_ERR_ARG;
  }
  const int HSIZE = 16384;                       // open addressing on (prefix << 8 | byte)
  static thread_local int32_t hkey[HSIZE];
  static thread_local uint16_t hval[HSIZE];
  uint64_t acc = 0;
  int nbits = 0;
  int64_t op = 0;
  int width = 9, next = LZW_FIRST;
  auto put = [&](int code) {
    acc = (acc << width) | (uint32_t)code;
    nbits += width;
    while (nbits >= 8) { dst[op++] = (uint8_t)(acc >> (nbits - 8)); nbits -= 8; }
  };
  auto reset = [&]() { memset(hkey, 0xff, sizeof(int32_t) * HSIZE); width = 9; next = LZW_FIRST; };
  reset();
  put(LZW_CLEAR);
  int64_t ip = 0;
  int omega = -1;
  while (ip < nsrc) {
    int k = src[ip++];
    if (omega < 0) { omega = k; continue; }
    int32_t key = (omega << 8) | k;
    int h = (int)(((uint32_t)key * 2654435761u) >> 18) & (HSIZE - 1);
    bool found = false;
    while (hkey[h] != -1) {
      if (hkey[h] == key) { omega = hval[h]; found = true; break; }
      h = (h + 1) & (HSIZE - 1);
    }
    if (found) continue;
    put(omega);
    hkey[h] = key;
    hval[h] = (uint16_t)next;
    next++;
    omega = k;
    if (next >= LZW_MAX - 2) {
      put(LZW_CLEAR);
      reset();
    } else if (next >= (1 << width) && width < 12) {
      width++;
    }
  }
  if (omega >= 0) {
    put(omega);
    next++;
    if (next >= LZW_MAX - 2) {
      put(LZW_CLEAR);
      width = 9;
    } else if (next >= (1 << width) && width < 12) {
      width++;
    }
  }
  put(LZW_EOI);
  if (nbits > 0) dst[op++] = (uint8_t)(acc << (8 - nbits));
  if (nout) *nout = op;
  return HC_OK;
}
