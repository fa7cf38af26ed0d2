// Micro-benchmark of the 16x16 diagonal tile factorisation (chol_diag_factor) of the solve kernel: alone in its CTA, and next to
// warps that run the trailing update's inner loop (LDS + DFMA on other tiles) the way phase B of cholesky_tiles does.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a [-DVILS_DIAG_PIPELINED=0] -o tools/diag_micro tools/diag_micro.cu
#include <cstdio>
#include "../mvil_fusion_b200/csrc/ba_device.cuh"
using namespace vb;

__device__ void fill_tile(double* tile, int lane) {
  for (int e = lane; e < 256; e += 32) { int i = e / 16, j = e % 16; tile[i * TLD + j] = (i == j ? 20.0 + i : 1.0 / (1 + abs(i - j))); }
}

// mode 0: only the diagonal warp works.  mode 1: warps of the other three sub-partitions run update_tile-like loops meanwhile.
// mode 2: every other warp (the diagonal warp's own sub-partition included) does.
__global__ void __launch_bounds__(SOLVE_THREADS, 1) k(double* out, int reps, int mode, int variant) {
  extern __shared__ double sm[];
  double* tile = sm;                       // the diagonal tile
  double* dinv = sm + TSZ + 16 * TLD;      // (reads past the tile stay inside the buffer)
  double* Li = dinv + 16;
  double* work = Li + 256;                 // 3 tiles per updater warp
  __shared__ int flag; __shared__ volatile int done;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, dw = SOLVE_WARPS - 1;
  for (int e = t; e < SOLVE_WARPS * 3 * TSZ; e += blockDim.x) work[e] = 1e-3 * (e % 97);
  if (t == 0) done = 0;
  __syncthreads();
  if (warp == dw) {
    long long tot = 0;
    for (int rep = 0; rep < reps; rep++) {
      fill_tile(tile, lane);
      __syncwarp();
      const long long t0 = clock64();
      if (variant == 0) chol_diag_factor<true>(tile, dinv, &flag, lane); else chol_diag_inverse(tile, dinv, Li);
      const long long t1 = clock64();
      tot += t1 - t0;
      __syncwarp();
    }
    if (lane == 0) { out[0] = (double)tot / reps; out[1] = tile[5 * TLD + 3]; out[2] = dinv[7]; done = 1; }
  } else if (mode == 3 && (warp & 3) != (dw & 3)) {          // DFMA only, no shared-memory traffic
    double a0 = lane, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7; const double bb = 1.0000001, cc = 1e-9;
    while (!done) {
#pragma unroll
      for (int q = 0; q < 8; q++) { a0 = fma(a0, bb, cc); a1 = fma(a1, bb, cc); a2 = fma(a2, bb, cc); a3 = fma(a3, bb, cc); a4 = fma(a4, bb, cc); a5 = fma(a5, bb, cc); a6 = fma(a6, bb, cc); a7 = fma(a7, bb, cc); }
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 1.2345) out[3] = a0;
  } else if (mode == 4 && (warp & 3) != (dw & 3)) {          // shared-memory loads only
    const double* A = work + (size_t)warp * 3 * TSZ; double acc = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    while (!done) {
#pragma unroll
      for (int q = 0; q < 16; q += 4) { acc += A[(lane & 15) * TLD + q]; acc1 += A[(lane & 15) * TLD + q + 1]; acc2 += A[(lane & 15) * TLD + q + 2]; acc3 += A[(lane & 15) * TLD + q + 3]; }
    }
    if (acc + acc1 + acc2 + acc3 == 1.2345) out[3] = acc;
  } else if (mode == 2 || ((mode == 1 || mode == 5) && (warp & 3) != (dw & 3))) {
    double* A = work + (size_t)warp * 3 * TSZ; const double* Lik = A + TSZ; const double* Ljk = A + 2 * TSZ;
    const int sub = lane & 15, r0 = (sub >> 2) * 4, c0 = (sub & 3) * 4;
    while (!done) {
      double acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[a][c] = 0;
#pragma unroll 4
      for (int m = 0; m < 16; m++) {
        double av[4], bv[4];
#pragma unroll
        for (int a = 0; a < 4; a++) av[a] = Lik[(r0 + a) * TLD + m];
#pragma unroll
        for (int c = 0; c < 4; c++) bv[c] = Ljk[(c0 + c) * TLD + m];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
          for (int c = 0; c < 4; c++) acc[a][c] = fma(av[a], bv[c], acc[a][c]);
      }
      if (mode != 5 && lane < 16) {
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
          for (int c = 0; c < 4; c++) A[(r0 + a) * TLD + c0 + c] -= 1e-9 * acc[a][c];
      }
      if (mode == 5 && acc[0][0] + acc[1][1] + acc[2][2] + acc[3][3] + acc[0][3] + acc[3][0] == 1.2345) out[3] = 1;
    }
  }
}

int main() {
  double* d; cudaMalloc(&d, 64); double h[4];
  const size_t smem = (size_t)(TSZ + 16 * TLD + 16 + 256 + SOLVE_WARPS * 3 * TSZ) * 8;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int v = 0; v < 2; v++)
    for (int mode = 0; mode < 6; mode++) {
      k<<<1, SOLVE_THREADS, smem>>>(d, 50, mode, v);
      cudaError_t e = cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
      printf("%s, %s: %.0f cycles per tile  (check %.9f %.9f)%s\n", v ? "inverse" : "factor ",
             mode == 0 ? "alone                               " : mode == 1 ? "3 sub-partitions run trailing update" : mode == 2 ? "4 sub-partitions run trailing update" : mode == 3 ? "3 sub-partitions run DFMA only      " : mode == 4 ? "3 sub-partitions run LDS (+DADD)    " : "3 sub-partitions: update, no stores ", h[0], h[1], h[2],
             e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
