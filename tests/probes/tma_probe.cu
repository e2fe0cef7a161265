// Stand-alone probe of the TMA helpers in csrc/tma.cuh (build: see tests/probes/run_tma_probe.sh).
// usage: tma_probe box_w box_h x y
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../pydrodem_b200/csrc/tma.cuh"

__global__ void k_probe(const __grid_constant__ CUtensorMap m, float *out, int x, int y, int n) {
  extern __shared__ __align__(128) unsigned char smem[];
  float *sz = (float *)smem;
  uint64_t *bar = (uint64_t *)(smem + ((n * 4 + 127) & ~127));
  if (threadIdx.x == 0) {
    hc_mbar_init(bar, 1);
    hc_mbar_expect_tx(bar, n * 4);
    hc_tma_load_2d(sz, &m, bar, x, y);
  }
  __syncthreads();
  hc_mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = sz[i];
}

int main(int argc, char **argv) {
  const int BW = atoi(argv[1]), BH = atoi(argv[2]), x = atoi(argv[3]), y = atoi(argv[4]);
  const int ny = 100, nx = 128;
  std::vector<float> h(ny * nx);
  for (int i = 0; i < ny * nx; i++) h[i] = (float)i;
  float *d, *o;
  cudaMalloc(&d, ny * nx * 4);
  cudaMalloc(&o, BW * BH * 4);
  cudaMemcpy(d, h.data(), ny * nx * 4, cudaMemcpyHostToDevice);
  hc_tmap tm;
  int rc = hc_tmap_2d(&tm, d, 4, 1, nx, ny, (uint64_t)nx * 4, BW, BH);
  printf("box %dx%d at (%d,%d): encode rc=%d ok=%d; ", BW, BH, x, y, rc, tm.ok);
  if (!tm.ok) { printf("\n"); return 2; }
  int bad = 0;
  k_probe<<<1, 256, BW * BH * 4 + 256>>>(tm.map, o, x, y, BW * BH);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s; ", cudaGetErrorString(e));
  if (e != cudaSuccess) { printf("\n"); return 3; }
  std::vector<float> r(BW * BH);
  cudaMemcpy(r.data(), o, BW * BH * 4, cudaMemcpyDeviceToHost);
  for (int i = 0; i < BH; i++)
    for (int j = 0; j < BW; j++) {
      int rr = y + i, cc = x + j;
      float v = r[i * BW + j];
      bool in = rr >= 0 && rr < ny && cc >= 0 && cc < nx;
      if (in ? (v != h[rr * nx + cc]) : !std::isnan(v)) bad++;
    }
  printf("bad=%d\n", bad);
  return bad ? 1 : 0;
}
