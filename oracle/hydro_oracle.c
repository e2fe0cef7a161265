/*
 * hydro_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call this file.  The product path (pydrodem_b200/) never does.
 *
 * PARITY UNPINNED for fill / D8 / flats / accumulation: pydrodem has no arithmetic of its
 * own for these passes — it shells out to WhiteboxTools (Rust) and TauDEM (C++/MPI), neither
 * of which is vendored, version-pinned or installable here, and the
 * reference ships no tests or golden rasters.  Every function below restates the PUBLISHED
 * algorithm named in its header under the call parameters pydrodem uses, citing the
 * reference call site it stands in for.  Tool-specific conventions that could not be checked
 * against a binary (seed definition, neighbour search order, code tables, self-inclusive
 * counts) are oracle PARAMETERS, not hard-coded facts.
 *
 * Plain C99, single-threaded on purpose: WhiteboxTools' fill is a single-threaded binary
 * heap, so this is also the algorithmic class of the CPU baseline.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_SEED_WBT 0    /* seeds = valid cells 8-adjacent to exterior-connected nodata / raster frame */
#define ORC_SEED_TAUDEM 1 /* seeds = valid cells on the raster border or 8-adjacent to ANY nodata      */
#define ORC_TABLE_WBT 0
#define ORC_TABLE_TAUDEM 1
#define ORC_DIR_NOFLOW 0
#define ORC_DIR_NODATA 255

static inline int is_nodata(float v, float nodata) { return v == nodata || v != v; }

/* ---- neighbour tables ------------------------------------------------------------------
 * WBT  (D8Pointer, call site run_aws.py:539): search order NE,E,SE,S,SW,W,NW,N with codes
 *      1,2,4,8,16,32,64,128; 0 = no lower neighbour.
 * TauDEM (D8FlowDir, call site run_aws.py:594-603): k=1..8 = E,NE,N,NW,W,SW,S,SE, code = k.
 * Row index grows southwards (north-up raster, tools/convert.py:128-169). */
static const int WBT_DR[8] = {-1, 0, 1, 1, 1, 0, -1, -1};
static const int WBT_DC[8] = {1, 1, 1, 0, -1, -1, -1, 0};
static const int WBT_CODE[8] = {1, 2, 4, 8, 16, 32, 64, 128};
static const int TAU_DR[8] = {0, -1, -1, -1, 0, 1, 1, 1};
static const int TAU_DC[8] = {1, 1, 0, -1, -1, -1, 0, 1};
static const int TAU_CODE[8] = {1, 2, 3, 4, 5, 6, 7, 8};

static void table_get(int table, const int **dr, const int **dc, const int **code) {
  if (table == ORC_TABLE_WBT) { *dr = WBT_DR; *dc = WBT_DC; *code = WBT_CODE; }
  else { *dr = TAU_DR; *dc = TAU_DC; *code = TAU_CODE; }
}

/* code -> neighbour slot k (0..7) or -1 */
static int code_to_k(int table, int code) {
  const int *dr, *dc, *cd;
  table_get(table, &dr, &dc, &cd);
  for (int k = 0; k < 8; k++) if (cd[k] == code) return k;
  return -1;
}

/* ---- seed mask ---------------------------------------------------------------------------
 * WBT_WL: FillDepressionsWangAndLiu (call site tools/depressions.py:1284) floods inwards from
 * a one-cell frame around the raster through 8-connected nodata; the first valid cells met
 * are the priority-queue seeds.  Interior nodata holes are neither seeds nor crossed.
 * TAUDEM: PitRemove (run_aws.py:582-588), Planchon & Darboux 2001: border cells and cells next
 * to any nodata keep their elevation. */
int orc_seed_mask(const float *z, uint8_t *seed, int64_t ny, int64_t nx, float nodata, int mode) {
  int64_t n = ny * nx;
  memset(seed, 0, (size_t)n);
  if (mode == ORC_SEED_TAUDEM) {
    for (int64_t r = 0; r < ny; r++)
      for (int64_t c = 0; c < nx; c++) {
        int64_t i = r * nx + c;
        if (is_nodata(z[i], nodata)) continue;
        int s = (r == 0 || c == 0 || r == ny - 1 || c == nx - 1);
        for (int dr = -1; dr <= 1 && !s; dr++)
          for (int dc = -1; dc <= 1; dc++) {
            int64_t rr = r + dr, cc = c + dc;
            if (rr < 0 || cc < 0 || rr >= ny || cc >= nx) continue;
            if (is_nodata(z[rr * nx + cc], nodata)) { s = 1; break; }
          }
        seed[i] = (uint8_t)s;
      }
    return 0;
  }
  /* WBT: BFS over exterior nodata starting from the virtual frame */
  uint8_t *ext = (uint8_t *)calloc((size_t)n, 1);
  int64_t *stack = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  if (!ext || !stack) { free(ext); free(stack); return -1; }
  int64_t sp = 0;
  for (int64_t r = 0; r < ny; r++)
    for (int64_t c = 0; c < nx; c++) {
      if (!(r == 0 || c == 0 || r == ny - 1 || c == nx - 1)) continue;
      int64_t i = r * nx + c;
      if (is_nodata(z[i], nodata)) { if (!ext[i]) { ext[i] = 1; stack[sp++] = i; } }
      else seed[i] = 1;
    }
  while (sp > 0) {
    int64_t i = stack[--sp];
    int64_t r = i / nx, c = i % nx;
    for (int dr = -1; dr <= 1; dr++)
      for (int dc = -1; dc <= 1; dc++) {
        if (!dr && !dc) continue;
        int64_t rr = r + dr, cc = c + dc;
        if (rr < 0 || cc < 0 || rr >= ny || cc >= nx) continue;
        int64_t j = rr * nx + cc;
        if (is_nodata(z[j], nodata)) { if (!ext[j]) { ext[j] = 1; stack[sp++] = j; } }
        else seed[j] = 1;
      }
  }
  free(ext); free(stack);
  return 0;
}

/* ---- binary min-heap on (key, seq) --------------------------------------------------------*/
typedef struct { double key; uint64_t seq; int64_t idx; } hnode;
typedef struct { hnode *a; int64_t n, cap; } heap;
static int heap_push(heap *h, double key, uint64_t seq, int64_t idx) {
  if (h->n == h->cap) {
    int64_t nc = h->cap ? h->cap * 2 : 1024;
    hnode *na = (hnode *)realloc(h->a, sizeof(hnode) * (size_t)nc);
    if (!na) return -1;
    h->a = na; h->cap = nc;
  }
  int64_t i = h->n++;
  while (i > 0) {
    int64_t p = (i - 1) >> 1;
    if (h->a[p].key < key || (h->a[p].key == key && h->a[p].seq < seq)) break;
    h->a[i] = h->a[p]; i = p;
  }
  h->a[i].key = key; h->a[i].seq = seq; h->a[i].idx = idx;
  return 0;
}
static hnode heap_pop(heap *h) {
  hnode top = h->a[0];
  hnode last = h->a[--h->n];
  int64_t i = 0;
  for (;;) {
    int64_t l = 2 * i + 1, r = l + 1, m;
    if (l >= h->n) break;
    m = l;
    if (r < h->n && (h->a[r].key < h->a[l].key || (h->a[r].key == h->a[l].key && h->a[r].seq < h->a[l].seq))) m = r;
    if (last.key < h->a[m].key || (last.key == h->a[m].key && last.seq < h->a[m].seq)) break;
    h->a[i] = h->a[m]; i = m;
  }
  if (h->n > 0) h->a[i] = last;
  return top;
}

/* ---- depression fill: priority-flood (Wang & Liu 2006; Barnes, Lehman & Mulla 2014 Alg. 1) ----
 * Stands in for wbt.fill_depressions_wang_and_liu / wbt.fill_depressions with fix_flats=False
 * (tools/depressions.py:1277-1286, default flat_increment=0.0 config_local.ini:161) and for
 * TauDEM PitRemove (run_aws.py:582-588).  out = max(z, lowest spill level to a seed); every
 * output value is an input float32 value, so the result is independent of visiting order.
 * Valid cells never reached (walled in by interior nodata) keep their input elevation. */
int orc_fill(const float *z, float *out, int64_t ny, int64_t nx, float nodata, int seed_mode) {
  int64_t n = ny * nx;
  uint8_t *seed = (uint8_t *)malloc((size_t)(n > 0 ? n : 1));
  uint8_t *done = (uint8_t *)calloc((size_t)(n > 0 ? n : 1), 1);
  heap h = {0, 0, 0};
  uint64_t seq = 0;
  if (!seed || !done) { free(seed); free(done); return -1; }
  if (orc_seed_mask(z, seed, ny, nx, nodata, seed_mode)) { free(seed); free(done); return -1; }
  for (int64_t i = 0; i < n; i++) {
    out[i] = z[i];
    if (seed[i]) { done[i] = 1; if (heap_push(&h, z[i], seq++, i)) return -1; }
  }
  while (h.n > 0) {
    hnode t = heap_pop(&h);
    int64_t r = t.idx / nx, c = t.idx % nx;
    float lvl = out[t.idx];
    for (int dr = -1; dr <= 1; dr++)
      for (int dc = -1; dc <= 1; dc++) {
        if (!dr && !dc) continue;
        int64_t rr = r + dr, cc = c + dc;
        if (rr < 0 || cc < 0 || rr >= ny || cc >= nx) continue;
        int64_t j = rr * nx + cc;
        if (done[j] || is_nodata(z[j], nodata)) continue;
        done[j] = 1;
        out[j] = z[j] > lvl ? z[j] : lvl;
        if (heap_push(&h, out[j], seq++, j)) return -1;
      }
  }
  free(h.a); free(seed); free(done);
  return 0;
}

/* Priority-Flood+epsilon in float64, the fix_flats=True variant used by the (disabled) WBT test
 * block run_aws.py:528-533 (flat_increment=1e-5).  Order dependent (ties pop in insertion order
 * here; WBT's BinaryHeap order on ties is implementation defined) => tolerance-class only. */
int orc_fill_eps(const float *z, double *out, int64_t ny, int64_t nx, float nodata, int seed_mode, double eps) {
  int64_t n = ny * nx;
  uint8_t *seed = (uint8_t *)malloc((size_t)(n > 0 ? n : 1));
  uint8_t *done = (uint8_t *)calloc((size_t)(n > 0 ? n : 1), 1);
  heap h = {0, 0, 0};
  uint64_t seq = 0;
  if (!seed || !done) { free(seed); free(done); return -1; }
  if (orc_seed_mask(z, seed, ny, nx, nodata, seed_mode)) { free(seed); free(done); return -1; }
  for (int64_t i = 0; i < n; i++) {
    out[i] = z[i];
    if (seed[i]) { done[i] = 1; if (heap_push(&h, z[i], seq++, i)) return -1; }
  }
  while (h.n > 0) {
    hnode t = heap_pop(&h);
    int64_t r = t.idx / nx, c = t.idx % nx;
    double lvl = out[t.idx] + eps;
    for (int dr = -1; dr <= 1; dr++)
      for (int dc = -1; dc <= 1; dc++) {
        if (!dr && !dc) continue;
        int64_t rr = r + dr, cc = c + dc;
        if (rr < 0 || cc < 0 || rr >= ny || cc >= nx) continue;
        int64_t j = rr * nx + cc;
        if (done[j] || is_nodata(z[j], nodata)) continue;
        done[j] = 1;
        out[j] = (double)z[j] > lvl ? (double)z[j] : lvl;
        if (heap_push(&h, out[j], seq++, j)) return -1;
      }
  }
  free(h.a); free(seed); free(done);
  return 0;
}

/* ---- single-cell pits (WBT FillSingleCellPits / BreachSingleCellPits) -----------------------
 * call sites models/depression_handling.py:259-261, models/find_sinks.py:118.
 * fill: an interior cell whose 8 neighbours are all valid and strictly higher is raised to the
 * lowest neighbour.  breach: for such a pit, every 2-ring cell lower than the pit lowers the
 * bridging 1-ring cell to the mean of pit and target.  Both read the INPUT raster only. */
int orc_fill_single_cell_pits(const float *z, float *out, int64_t ny, int64_t nx, float nodata) {
  memcpy(out, z, sizeof(float) * (size_t)(ny * nx));
  for (int64_t r = 1; r < ny - 1; r++)
    for (int64_t c = 1; c < nx - 1; c++) {
      float v = z[r * nx + c];
      if (is_nodata(v, nodata)) continue;
      float mn = INFINITY; int pit = 1;
      for (int k = 0; k < 8 && pit; k++) {
        float zn = z[(r + WBT_DR[k]) * nx + c + WBT_DC[k]];
        if (is_nodata(zn, nodata) || !(zn > v)) pit = 0;
        else if (zn < mn) mn = zn;
      }
      if (pit) out[r * nx + c] = mn;
    }
  return 0;
}

int orc_breach_single_cell_pits(const float *z, float *out, int64_t ny, int64_t nx, float nodata) {
  static const int DC2[16] = {2, 2, 2, 2, 2, 1, 0, -1, -2, -2, -2, -2, -2, -1, 0, 1};
  static const int DR2[16] = {-2, -1, 0, 1, 2, 2, 2, 2, 2, 1, 0, -1, -2, -2, -2, -2};
  static const int BRIDGE[16] = {0, 1, 1, 1, 2, 3, 3, 3, 4, 5, 5, 5, 6, 7, 7, 7}; /* slot in WBT_DR/DC */
  memcpy(out, z, sizeof(float) * (size_t)(ny * nx));
  for (int64_t r = 2; r < ny - 2; r++)
    for (int64_t c = 2; c < nx - 2; c++) {
      float v = z[r * nx + c];
      if (is_nodata(v, nodata)) continue;
      int pit = 1;
      for (int k = 0; k < 8 && pit; k++) {
        float zn = z[(r + WBT_DR[k]) * nx + c + WBT_DC[k]];
        if (is_nodata(zn, nodata) || !(zn > v)) pit = 0;
      }
      if (!pit) continue;
      for (int k = 0; k < 16; k++) {
        float z2 = z[(r + DR2[k]) * nx + c + DC2[k]];
        if (is_nodata(z2, nodata) || !(z2 < v)) continue;
        int b = BRIDGE[k];
        out[(r + WBT_DR[b]) * nx + c + WBT_DC[b]] = (float)(((double)v + (double)z2) / 2.0);
      }
    }
  return 0;
}

/* ---- D8 direction + slope (O'Callaghan & Mark 1984) -------------------------------------------
 * table WBT: slope_k = (double)(z - zk) / len_k, first strict maximum > 0 in WBT search order;
 *            nodata neighbours skipped; no lower neighbour -> 0; nodata cell -> 255.
 * table TAUDEM: slope_k = (float)(1/len_k) * (z - zk) in float32, first strict maximum > 0 for
 *            k = 1..8; a cell on the raster border or with any nodata neighbour is
 *            "contaminated" and gets the nodata code 255 (what models/taudem_check.py:143-147
 *            hunts for on water cells).
 * slope output (sd8, run_aws.py:601) = max drop/distance as float32, 0 where none, -1 at 255. */
int orc_d8(const float *z, uint8_t *dir, float *slope, int64_t ny, int64_t nx, float nodata,
           double dx, double dy, int table) {
  const int *dr, *dc, *code;
  double len[8];
  float fact[8];
  table_get(table, &dr, &dc, &code);
  for (int k = 0; k < 8; k++) {
    len[k] = sqrt((double)(dc[k] * dc[k]) * dx * dx + (double)(dr[k] * dr[k]) * dy * dy);
    fact[k] = (float)(1.0 / len[k]);
  }
  for (int64_t r = 0; r < ny; r++)
    for (int64_t c = 0; c < nx; c++) {
      int64_t i = r * nx + c;
      float v = z[i];
      if (is_nodata(v, nodata)) { dir[i] = ORC_DIR_NODATA; if (slope) slope[i] = -1.0f; continue; }
      int best = -1, contaminated = 0;
      double smax = 0.0; float smaxf = 0.0f;
      for (int k = 0; k < 8; k++) {
        int64_t rr = r + dr[k], cc = c + dc[k];
        if (rr < 0 || cc < 0 || rr >= ny || cc >= nx) { contaminated = 1; continue; }
        float zn = z[rr * nx + cc];
        if (is_nodata(zn, nodata)) { contaminated = 1; continue; }
        if (table == ORC_TABLE_WBT) {
          double s = ((double)v - (double)zn) / len[k];
          if (s > smax) { smax = s; best = k; }
        } else {
          float s = fact[k] * (v - zn);
          if (s > smaxf) { smaxf = s; best = k; }
        }
      }
      if (table == ORC_TABLE_TAUDEM && contaminated) { dir[i] = ORC_DIR_NODATA; if (slope) slope[i] = -1.0f; continue; }
      dir[i] = best < 0 ? ORC_DIR_NOFLOW : (uint8_t)code[best];
      if (slope) slope[i] = table == ORC_TABLE_WBT ? (float)smax : smaxf;
    }
  return 0;
}

/* ---- flat resolution (Barnes, Lehman & Mulla 2014, "An efficient assignment of drainage
 * direction over flat surfaces", Algorithms 1-7; the two-gradient idea of Garbrecht & Martz 1997
 * that TauDEM D8FlowDir applies, run_aws.py:594-603).  Fills in directions for NOFLOW cells of
 * drainable flats; undrainable flats stay NOFLOW.  A valid cell carrying code 255 (TauDEM
 * contaminated cell) counts as "has a direction" (it drains off the domain).
 * Tie rule on the flat mask: lowest mask strictly below own; among equals the first cardinal
 * neighbour in table order, else the first diagonal.  */
int orc_resolve_flats(const float *z, uint8_t *dir, int64_t ny, int64_t nx, float nodata, int table) {
  const int *dr, *dc, *code;
  int64_t n = ny * nx;
  table_get(table, &dr, &dc, &code);
  int32_t *label = (int32_t *)calloc((size_t)(n > 0 ? n : 1), sizeof(int32_t));
  int32_t *mask = (int32_t *)calloc((size_t)(n > 0 ? n : 1), sizeof(int32_t));
  int64_t *q = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
  int64_t *low = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
  int64_t *high = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
  int32_t *fheight = NULL;
  int64_t nlow = 0, nhigh = 0;
  if (!label || !mask || !q || !low || !high) return -1;
#define VALID(i) (!is_nodata(z[i], nodata))
  /* Alg. 3 FlatEdges */
  for (int64_t r = 0; r < ny; r++)
    for (int64_t c = 0; c < nx; c++) {
      int64_t i = r * nx + c;
      if (!VALID(i)) continue;
      for (int k = 0; k < 8; k++) {
        int64_t rr = r + dr[k], cc = c + dc[k];
        if (rr < 0 || cc < 0 || rr >= ny || cc >= nx) continue;
        int64_t j = rr * nx + cc;
        if (!VALID(j)) continue;
        if (dir[i] != ORC_DIR_NOFLOW && dir[j] == ORC_DIR_NOFLOW && z[i] == z[j]) { low[nlow++] = i; break; }
        if (dir[i] == ORC_DIR_NOFLOW && z[i] < z[j]) { high[nhigh++] = i; break; }
      }
    }
  /* Alg. 4 LabelFlats: flood equal elevations from every low edge */
  int32_t nlab = 0;
  for (int64_t s = 0; s < nlow; s++) {
    if (label[low[s]]) continue;
    nlab++;
    float e = z[low[s]];
    int64_t qh = 0, qt = 0;
    q[qt++] = low[s]; label[low[s]] = nlab;
    while (qh < qt) {
      int64_t i = q[qh++], r = i / nx, c = i % nx;
      for (int k = 0; k < 8; k++) {
        int64_t rr = r + dr[k], cc = c + dc[k];
        if (rr < 0 || cc < 0 || rr >= ny || cc >= nx) continue;
        int64_t j = rr * nx + cc;
        if (!VALID(j) || label[j] || z[j] != e) continue;
        label[j] = nlab; q[qt++] = j;
      }
    }
  }
  fheight = (int32_t *)calloc((size_t)nlab + 1, sizeof(int32_t));
  if (!fheight) return -1;
  /* Alg. 5 AwayFromHigher (level-synchronous BFS through NOFLOW cells of the same label) */
  {
    int64_t qh = 0, qt = 0;
    for (int64_t s = 0; s < nhigh; s++) if (label[high[s]]) q[qt++] = high[s];
    int32_t loops = 1;
    int64_t level_end = qt;
    int64_t *q2 = q; /* q holds at most n entries because each cell is enqueued once below */
    uint8_t *inq = (uint8_t *)calloc((size_t)(n > 0 ? n : 1), 1);
    if (!inq) return -1;
    for (int64_t s = 0; s < qt; s++) inq[q[s]] = 1;
    while (qh < qt) {
      if (qh == level_end) { loops++; level_end = qt; }
      int64_t i = q2[qh++], r = i / nx, c = i % nx;
      mask[i] = loops;
      if (loops > fheight[label[i]]) fheight[label[i]] = loops;
      for (int k = 0; k < 8; k++) {
        int64_t rr = r + dr[k], cc = c + dc[k];
        if (rr < 0 || cc < 0 || rr >= ny || cc >= nx) continue;
        int64_t j = rr * nx + cc;
        if (inq[j] || label[j] != label[i] || dir[j] != ORC_DIR_NOFLOW) continue;
        inq[j] = 1; q2[qt++] = j;
      }
    }
    free(inq);
  }
  /* Alg. 6 TowardsLower */
  {
    int64_t qh = 0, qt = 0;
    uint8_t *inq = (uint8_t *)calloc((size_t)(n > 0 ? n : 1), 1);
    if (!inq) return -1;
    for (int64_t i = 0; i < n; i++) mask[i] = -mask[i];
    for (int64_t s = 0; s < nlow; s++) if (!inq[low[s]]) { inq[low[s]] = 1; q[qt++] = low[s]; }
    int32_t loops = 1;
    int64_t level_end = qt;
    while (qh < qt) {
      if (qh == level_end) { loops++; level_end = qt; }
      int64_t i = q[qh++], r = i / nx, c = i % nx;
      if (mask[i] < 0) mask[i] = fheight[label[i]] + mask[i] + 2 * loops;
      else mask[i] = 2 * loops;
      for (int k = 0; k < 8; k++) {
        int64_t rr = r + dr[k], cc = c + dc[k];
        if (rr < 0 || cc < 0 || rr >= ny || cc >= nx) continue;
        int64_t j = rr * nx + cc;
        if (inq[j] || label[j] != label[i] || dir[j] != ORC_DIR_NOFLOW) continue;
        inq[j] = 1; q[qt++] = j;
      }
    }
    free(inq);
  }
  /* Alg. 7 D8MaskedFlowDirs — computed into a copy so that the NOFLOW tests stay on the input */
  uint8_t *nd = (uint8_t *)malloc((size_t)(n > 0 ? n : 1));
  if (!nd) return -1;
  memcpy(nd, dir, (size_t)n);
  for (int64_t r = 0; r < ny; r++)
    for (int64_t c = 0; c < nx; c++) {
      int64_t i = r * nx + c;
      if (!VALID(i) || dir[i] != ORC_DIR_NOFLOW || !label[i]) continue;
      int32_t mn = mask[i]; int best = -1;
      for (int k = 0; k < 8; k++) {
        int64_t rr = r + dr[k], cc = c + dc[k];
        if (rr < 0 || cc < 0 || rr >= ny || cc >= nx) continue;
        int64_t j = rr * nx + cc;
        if (label[j] != label[i]) continue;
        int kdiag = dr[k] != 0 && dc[k] != 0;
        int bdiag = best >= 0 && dr[best] != 0 && dc[best] != 0;
        if (mask[j] < mn || (mask[j] == mn && best >= 0 && bdiag && !kdiag)) { mn = mask[j]; best = k; }
      }
      if (best >= 0) nd[i] = (uint8_t)code[best];
    }
  memcpy(dir, nd, (size_t)n);
#undef VALID
  free(nd); free(fheight); free(label); free(mask); free(q); free(low); free(high);
  return 0;
}

/* ---- D8 flow accumulation, cell counts (WBT D8FlowAccumulation pntr=True, run_aws.py:544;
 * TauDEM AreaD8 -nc, run_aws.py:608-617): in-degree count, then topological push.  Every cell
 * with a code other than 255 counts itself (1); 255 cells hold 0 and pass nothing on.  A pointer
 * leaving the raster or landing on a 255 cell ends the path there.  Counts are uint32. */
int orc_accum(const uint8_t *dir, uint32_t *acc, int64_t ny, int64_t nx, int table) {
  const int *dr, *dc, *code;
  int64_t n = ny * nx;
  table_get(table, &dr, &dc, &code);
  int8_t lut[256];
  for (int i = 0; i < 256; i++) lut[i] = (int8_t)code_to_k(table, i);
  int64_t *down = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  uint8_t *indeg = (uint8_t *)calloc((size_t)(n > 0 ? n : 1), 1);
  int64_t *q = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  if (!down || !indeg || !q) return -1;
  for (int64_t r = 0; r < ny; r++)
    for (int64_t c = 0; c < nx; c++) {
      int64_t i = r * nx + c;
      down[i] = -1;
      acc[i] = dir[i] == ORC_DIR_NODATA ? 0u : 1u;
      if (dir[i] == ORC_DIR_NODATA || dir[i] == ORC_DIR_NOFLOW) continue;
      int k = lut[dir[i]];
      if (k < 0) continue;
      int64_t rr = r + dr[k], cc = c + dc[k];
      if (rr < 0 || cc < 0 || rr >= ny || cc >= nx) continue;
      int64_t j = rr * nx + cc;
      if (dir[j] == ORC_DIR_NODATA) continue;
      down[i] = j;
    }
  for (int64_t i = 0; i < n; i++) if (down[i] >= 0) indeg[down[i]]++;
  int64_t qh = 0, qt = 0;
  for (int64_t i = 0; i < n; i++) if (dir[i] != ORC_DIR_NODATA && indeg[i] == 0) q[qt++] = i;
  while (qh < qt) {
    int64_t i = q[qh++], j = down[i];
    if (j < 0) continue;
    acc[j] += acc[i];
    if (--indeg[j] == 0) q[qt++] = j;
  }
  free(down); free(indeg); free(q);
  return 0;
}
