// accum.cu — D8 flow accumulation (cell counts, self-inclusive).
//
// Replaces wbt.d8_flow_accumulation(pntr=True) (run_aws.py:544) and TauDEM AreaD8 -nc
// (run_aws.py:608-617): in-degree counting, then dependency-driven propagation.
//
// Per cell one 64-bit state word  [ acc : 56 | hasdown : 1 | k : 3 | cnt : 4 ]
//   cnt  = number of upstream neighbours that have not delivered yet (15 = source marker)
//   k    = neighbour slot this cell drains to, hasdown = that target exists and is not 255
//   acc  = cells accumulated so far (starts at 1)
// A thread that owns a finished cell delivers with ONE atomicAdd of ((acc << 8) - 1) on the
// target's word: the add credits the count and retires one dependency at once, and the returned
// old word tells the thread whether it was the last contributor (old cnt == 1), in which case it
// owns the target (its total is old.acc + acc, its k is in the same word) and walks on.  No
// fences, no second load per hop.
#include "common.cuh"

#define ACC_SRC 15ull

template <int TABLE>
__global__ void __launch_bounds__(256)
k_acc_init(const uint8_t *__restrict__ dir, u64 *__restrict__ state, int ny, int nx) {
  // 32 x 64 tile per block with halo in smem
  const int TR = 32, TC = 64, W = TC + 2;
  __shared__ uint8_t sd[(TR + 2) * (TC + 2)];
  const int r0 = blockIdx.y * TR, c0 = blockIdx.x * TC;
  for (int i = threadIdx.x; i < (TR + 2) * W; i += 256) {
    int lr = i / W, lc = i - lr * W;
    int r = r0 + lr - 1, c = c0 + lc - 1;
    uint8_t d = HC_DIR_NODATA;
    if (r >= 0 && r < ny && c >= 0 && c < nx) d = __ldg(&dir[(size_t)r * nx + c]);
    sd[i] = d;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TR * TC; i += 256) {
    int lr = i / TC, lc = i - lr * TC;
    int r = r0 + lr, c = c0 + lc;
    if (r >= ny || c >= nx) continue;
    int ci = (lr + 1) * W + lc + 1;
    int d = sd[ci];
    u64 w = 0;
    if (d != HC_DIR_NODATA) {
      int indeg = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        // neighbour in slot k drains into me iff its code is the opposite slot's code
        int nd = sd[ci + c_dr[TABLE][k] * W + c_dc[TABLE][k]];
        int opp = (k + 4) & 7;
        if (nd == c_code[TABLE][opp]) indeg++;
      }
      int k = d_code_to_k(TABLE, d);
      u64 hasdown = 0;
      if (k >= 0) {
        int td = sd[ci + c_dr[TABLE][k] * W + c_dc[TABLE][k]];  // halo holds 255 outside the raster
        hasdown = td != HC_DIR_NODATA;
      } else {
        k = 0;
      }
      w = (1ull << 8) | (hasdown << 7) | ((u64)k << 4) | (indeg ? (u64)indeg : ACC_SRC);
    }
    state[(size_t)r * nx + c] = w;
  }
}

template <int TABLE>
__global__ void __launch_bounds__(256)
k_acc_walk(u64 *__restrict__ state, int ny, int nx) {
  size_t n = (size_t)ny * nx;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u64 w = state[i];
  if ((w & 15ull) != ACC_SRC) return;
  int r = (int)(i / (size_t)nx), c = (int)(i - (size_t)r * nx);
  u64 a = w >> 8;
  for (;;) {
    if (!(w & 0x80ull)) return;
    int k = (int)((w >> 4) & 7ull);
    r += c_dr[TABLE][k];
    c += c_dc[TABLE][k];
    u64 old = atomicAdd(&state[(size_t)r * nx + c], (a << 8) - 1ull);
    if ((old & 15ull) != 1ull) return;  // others still owe this cell
    a += old >> 8;
    w = old;
  }
}

__global__ void __launch_bounds__(256)
k_acc_out(const u64 *__restrict__ state, uint32_t *__restrict__ acc, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) acc[i] = (uint32_t)(state[i] >> 8);
}

size_t accum_workspace(int64_t ny, int64_t nx) { return hc_align((size_t)ny * nx * 8); }

int accum_run(hc_ctx *ctx, const uint8_t *dir, uint32_t *acc, int64_t ny, int64_t nx, int table,
              cudaStream_t st) {
  if (ny <= 0 || nx <= 0) return HC_OK;
  if (table != HC_TABLE_WBT && table != HC_TABLE_TAUDEM) { hc_set_error("accum: bad table %d", table); return HC_ERR_ARG; }
  size_t n = (size_t)ny * nx;
  HC_TRY(arena_reset(ctx));
  u64 *state = (u64 *)arena_take(ctx, n * 8);
  if (!state) return HC_ERR_NOMEM;
  dim3 tg((unsigned)((nx + 63) / 64), (unsigned)((ny + 31) / 32));
  unsigned wb = (unsigned)((n + 255) / 256);
  if (table == HC_TABLE_WBT) {
    HC_PROF(ctx, st, "k_acc_init");
    k_acc_init<HC_TABLE_WBT><<<tg, 256, 0, st>>>(dir, state, (int)ny, (int)nx);
    HC_LAUNCH_CHECK(ctx);
    HC_PROF(ctx, st, "k_acc_walk");
    k_acc_walk<HC_TABLE_WBT><<<wb, 256, 0, st>>>(state, (int)ny, (int)nx);
    HC_LAUNCH_CHECK(ctx);
  } else {
    HC_PROF(ctx, st, "k_acc_init");
    k_acc_init<HC_TABLE_TAUDEM><<<tg, 256, 0, st>>>(dir, state, (int)ny, (int)nx);
    HC_LAUNCH_CHECK(ctx);
    HC_PROF(ctx, st, "k_acc_walk");
    k_acc_walk<HC_TABLE_TAUDEM><<<wb, 256, 0, st>>>(state, (int)ny, (int)nx);
    HC_LAUNCH_CHECK(ctx);
  }
  HC_PROF(ctx, st, "k_acc_out");
  k_acc_out<<<ctx->sm_count * 8, 256, 0, st>>>(state, acc, n);
  HC_LAUNCH_CHECK(ctx);
  return HC_OK;
}
