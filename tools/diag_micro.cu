#include <cstdio>
#include "../mvil_fusion_b200/csrc/ba_device.cuh"
using namespace vb;
__global__ void k(double* out, int reps, int variant) {
  __shared__ double tile[TSZ + 16 * TLD + 16]; __shared__ double dinv[16]; __shared__ double Li[256]; __shared__ int flag;
  const int lane = threadIdx.x & 31;
  long long tot = 0;
  for (int rep = 0; rep < reps; rep++) {
    for (int e = lane; e < 256; e += 32) { int i = e / 16, j = e % 16; tile[i * TLD + j] = (i == j ? 20.0 + i : 1.0 / (1 + abs(i - j))); }
    __syncwarp();
    long long t0 = clock64();
    if (variant == 0) chol_diag_factor(tile, dinv, &flag); else chol_diag_inverse(tile, dinv, Li);
    long long t1 = clock64();
    tot += t1 - t0;
    __syncwarp();
  }
  if (lane == 0) { out[0] = (double)tot / reps; out[1] = tile[5 * TLD + 3]; out[2] = dinv[7]; }
}
int main() {
  double* d; cudaMalloc(&d, 64); double h[4];
  for (int v = 0; v < 2; v++) {
    k<<<1, 32>>>(d, 50, v); cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("variant %d (%s): %.0f cycles per tile  (check %.6f %.6f)\n", v, v ? "inverse" : "factor", h[0], h[1], h[2]);
  }
  return 0;
}
