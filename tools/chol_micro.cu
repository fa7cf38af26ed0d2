// Micro-benchmark of the solve kernel's dense step on its own: cholesky_tiles + backsub_tiles on a 157-dimensional SPD system held in shared
// memory by ONE CTA of SOLVE_THREADS threads (what every Gauss-Newton iteration of a config-2 window does), cycles by clock64, result checked
// against the residual |H x - b|.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -diag-suppress 177 -o tools/chol_micro.bin tools/chol_micro.cu
#include <cstdio>
#include <vector>
#include <cmath>
#include "../mvil_fusion_b200/csrc/ba_device.cuh"
using namespace vb;

__global__ void __launch_bounds__(SOLVE_THREADS, 1) k(const double* Hin, const double* bin, double* xout, double* out, int D, int nb, int reps, long long* prof) {
  extern __shared__ double sm[];
  double* H = sm;                               // tri(nb) tiles
  double* b = H + tri(nb) * TSZ + 16 * TLD;     // (reads past the last tile stay inside the buffer)
  double* dinv = b + nb * 16 + 16;
  double* dx = dinv + nb * 16;
  double* linv = dx + nb * 16;
  double* tmp = linv + nb * 256;
  __shared__ int flag;
  long long tc = 0, tb = 0;
  for (int rep = 0; rep < reps; rep++) {
    for (int e = threadIdx.x; e < tri(nb) * TSZ; e += blockDim.x) H[e] = 0.0;
    __syncthreads();
    for (int e = threadIdx.x; e < nb * 16 * nb * 16; e += blockDim.x) {
      const int i = e / (nb * 16), j = e % (nb * 16);
      if (i >= j) H[tidx(i, j)] = (i < D && j < D) ? Hin[(size_t)i * D + j] : (i == j ? 1.0 : 0.0);
    }
    for (int e = threadIdx.x; e < nb * 16; e += blockDim.x) b[e] = e < D ? bin[e] : 0.0;
    if (threadIdx.x == 0) flag = 0;
    __syncthreads();
    const long long t0 = clock64();
    cholesky_tiles<true>(H, b, linv, dinv, nb, &flag, prof);
    __syncthreads();
    const long long t1 = clock64();
    backsub_tiles(H, b, linv, dx, nb, tmp);
    __syncthreads();
    const long long t2 = clock64();
    tc += t1 - t0; tb += t2 - t1;
  }
  for (int e = threadIdx.x; e < D; e += blockDim.x) xout[e] = dx[e];
  if (threadIdx.x == 0) { out[0] = (double)tc / reps; out[1] = (double)tb / reps; out[2] = flag; }
}

int main(int argc, char** argv) {
  const int D = argc > 1 ? atoi(argv[1]) : 157, nb = (D + 15) / 16;
  std::vector<double> A((size_t)D * D), Hm((size_t)D * D, 0.0), b(D), x(D);
  unsigned s = 12345u; auto rnd = [&] { s = s * 1664525u + 1013904223u; return (double)(s >> 8) / (1 << 24) - 0.5; };
  for (auto& v : A) v = rnd();
  for (int i = 0; i < D; i++) for (int j = 0; j < D; j++) { double t = 0; for (int m = 0; m < D; m++) t += A[(size_t)i * D + m] * A[(size_t)j * D + m]; Hm[(size_t)i * D + j] = t + (i == j ? 1.0 : 0.0); }
  for (auto& v : b) v = rnd();
  double *dH, *db, *dxo, *dout; cudaMalloc(&dH, Hm.size() * 8); cudaMalloc(&db, D * 8); cudaMalloc(&dxo, D * 8); cudaMalloc(&dout, 64);
  cudaMemcpy(dH, Hm.data(), Hm.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), D * 8, cudaMemcpyHostToDevice);
  const size_t smem = (size_t)(tri(nb) * TSZ + 16 * TLD + nb * 16 * 3 + 16 + nb * 256 + 32) * 8;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long* dprof; cudaMalloc(&dprof, 16 * 8); cudaMemset(dprof, 0, 16 * 8);
  k<<<1, SOLVE_THREADS, smem>>>(dH, db, dxo, dout, D, nb, 20, dprof);
  double out[4]; cudaError_t e = cudaMemcpy(out, dout, 24, cudaMemcpyDeviceToHost); cudaMemcpy(x.data(), dxo, D * 8, cudaMemcpyDeviceToHost);
  double res = 0, bn = 0;
  for (int i = 0; i < D; i++) { double t = -b[i]; for (int j = 0; j < D; j++) t += Hm[(size_t)i * D + j] * x[j]; res += t * t; bn += b[i] * b[i]; }
  printf("D = %d (%d tile rows): cholesky %.0f cycles (%.0f per tile row), back-substitution %.0f cycles, flag %.0f, |Hx - b| / |b| = %.2e %s\n", D, nb, out[0],
         out[0] / nb, out[1], out[2], std::sqrt(res / bn), e == cudaSuccess ? "" : cudaGetErrorString(e));
  long long pr[16]; cudaMemcpy(pr, dprof, sizeof(pr), cudaMemcpyDeviceToHost);
  printf("  per factorisation, thread 0: phase B (own work) %lld, wait for the diagonal warp %lld, panel %lld, phase A %lld\n", pr[12] / 20, pr[13] / 20, pr[14] / 20, pr[15] / 20);
  return 0;
}
