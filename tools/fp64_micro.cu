// FP64 latency / throughput probe for B200 (what the fused solve kernel is bound by).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat_dfma(double* o, int n) { double a = o[0], b = 1.0000001, c = 1e-9; long long t0 = clock64();
  for (int i = 0; i < n; i++) { a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); }
  long long t1 = clock64(); if (threadIdx.x == 0) { o[1] = a; o[2] = (double)(t1 - t0) / (4.0 * n); } }
__global__ void lat_rsqrt(double* o, int n) { double a = o[0] + 2.0; long long t0 = clock64();
  for (int i = 0; i < n; i++) { a = rsqrt(a) + 1.5; a = rsqrt(a) + 1.5; }
  long long t1 = clock64(); if (threadIdx.x == 0) { o[1] = a; o[2] = (double)(t1 - t0) / (2.0 * n); } }
__global__ void lat_div(double* o, int n) { double a = o[0] + 2.0; long long t0 = clock64();
  for (int i = 0; i < n; i++) { a = 3.0 / a + 1.5; a = 3.0 / a + 1.5; }
  long long t1 = clock64(); if (threadIdx.x == 0) { o[1] = a; o[2] = (double)(t1 - t0) / (2.0 * n); } }
__global__ void lat_sqrt(double* o, int n) { double a = o[0] + 2.0; long long t0 = clock64();
  for (int i = 0; i < n; i++) { a = sqrt(a) + 1.5; a = sqrt(a) + 1.5; }
  long long t1 = clock64(); if (threadIdx.x == 0) { o[1] = a; o[2] = (double)(t1 - t0) / (2.0 * n); } }
__global__ void lat_shfl(double* o, int n) { double a = o[0] + threadIdx.x; long long t0 = clock64();
  for (int i = 0; i < n; i++) { a = __shfl_sync(0xffffffffu, a, (threadIdx.x + 1) & 31) + 1.0; a = __shfl_sync(0xffffffffu, a, 3) + 1.0; }
  long long t1 = clock64(); if (threadIdx.x == 0) { o[1] = a; o[2] = (double)(t1 - t0) / (2.0 * n); } }
__global__ void lat_lds(double* o, int n) { __shared__ double s[1024]; for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i * 7 + 1) % 1024; __syncthreads();
  int idx = threadIdx.x; double acc = 0; long long t0 = clock64();
  for (int i = 0; i < n; i++) { double v = s[idx]; idx = (int)v; acc += v; }
  long long t1 = clock64(); if (threadIdx.x == 0) { o[1] = acc; o[2] = (double)(t1 - t0) / n; } }
__global__ void thr_dfma(double* o, int n) { double a0 = o[0], a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7, b = 1.0000001, c = 1e-9;
  for (int i = 0; i < n; i++) { a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c); }
  if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.0) o[3] = a0; }
int main() {
  double* d; cudaMalloc(&d, 64); double h[4] = {1.0, 0, 0, 0}; cudaMemcpy(d, h, 32, cudaMemcpyHostToDevice);
  const int n = 20000;
#define RUN(k, name) k<<<1, 32>>>(d, n); cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost); printf("%-10s latency %.1f cycles\n", name, h[2]); h[0] = 1.0; cudaMemcpy(d, h, 32, cudaMemcpyHostToDevice);
  RUN(lat_dfma, "DFMA") RUN(lat_rsqrt, "rsqrt+add") RUN(lat_div, "div+add") RUN(lat_sqrt, "sqrt+add") RUN(lat_shfl, "shfl64+add") RUN(lat_lds, "LDS.64 chase")
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int threads : {128, 256, 512, 1024}) {
    thr_dfma<<<148, threads>>>(d, 1000); cudaEventRecord(e0); thr_dfma<<<148 * 2, threads>>>(d, 100000); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("DFMA throughput %4d thr/CTA x 296 CTAs: %.2f TFLOP/s\n", threads, 2.0 * 8 * 100000.0 * threads * 296 / (ms * 1e-3) / 1e12);
  }
  return 0;
}
