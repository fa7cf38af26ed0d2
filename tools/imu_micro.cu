#include <cstdio>
#include "../mvil_fusion_b200/csrc/factors.cuh"
using namespace vf;
__global__ void k(const double* pre, const double* poses, double* out, int variant) {
  __shared__ double J[450]; __shared__ double r[15];
  const int lane = threadIdx.x;
  for (int e = lane; e < 450; e += 32) J[e] = 0;
  __syncwarp();
  const double G[3] = {0, 0, 9.795};
  long long t0 = clock64();
  if (variant == 0) { if (lane == 0) imu_eval_raw(pre, G, poses, poses + 7, poses + 16, poses + 23, r, J); }
  else if (variant == 1) { if (lane < IMU_PARTS) imu_eval_part(lane, pre, G, poses, poses + 7, poses + 16, poses + 23, r, J); }
  else { imu_eval_raw(pre + 0, G, poses, poses + 7, poses + 16, poses + 23, r, lane == 0 ? J : nullptr); }   // all lanes same code (SIMD over factors)
  __syncwarp();
  long long t1 = clock64();
  if (lane == 0) { out[0] = (double)(t1 - t0); out[1] = J[3 * 30 + 3] + r[2]; }
}
int main() {
  double hpre[467], hposes[32];
  for (int i = 0; i < 467; i++) hpre[i] = 0.01 * (i % 17) - 0.05; hpre[3] = 0.01; hpre[4] = 0.02; hpre[5] = -0.01; hpre[6] = 0.999; hpre[16] = 0.1;
  for (int i = 0; i < 32; i++) hposes[i] = 0.1 * i; hposes[3] = 0; hposes[4] = 0; hposes[5] = 0.1; hposes[6] = 0.99; hposes[19] = 0.01; hposes[20] = 0; hposes[21] = 0.12; hposes[22] = 0.99;
  double *dpre, *dposes, *dout; cudaMalloc(&dpre, sizeof(hpre)); cudaMalloc(&dposes, sizeof(hposes)); cudaMalloc(&dout, 64);
  cudaMemcpy(dpre, hpre, sizeof(hpre), cudaMemcpyHostToDevice); cudaMemcpy(dposes, hposes, sizeof(hposes), cudaMemcpyHostToDevice);
  for (int v = 0; v < 3; v++) for (int rep = 0; rep < 2; rep++) { k<<<1, 32>>>(dpre, dposes, dout, v); double h[2]; cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost); printf("variant %d rep %d: %.0f cycles (check %.5f)\n", v, rep, h[0], h[1]); }
  return 0;
}
