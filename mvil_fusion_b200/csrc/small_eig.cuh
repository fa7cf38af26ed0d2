// small_eig.cuh — 3x3 symmetric eigen-decomposition shared by the LiDAR association and VGICP kernels.
#pragma once

namespace vils_eig {

// 3x3 symmetric eigen-decomposition (cyclic Jacobi, FP64): w ascending like Eigen::SelfAdjointEigenSolver, V columns = eigenvectors
__device__ inline void eig3(double A[3][3], double w[3], double V[3][3]) {
  #pragma unroll
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; sweep++) {
    const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    if (off < 1e-300 || off <= 1e-17 * (fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]))) break;
    #pragma unroll
    for (int p = 0; p < 2; p++)
      #pragma unroll
      for (int q = p + 1; q < 3; q++) {
        if (fabs(A[p][q]) < 1e-300) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        #pragma unroll
        for (int k = 0; k < 3; k++) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
        #pragma unroll
        for (int k = 0; k < 3; k++) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
        #pragma unroll
        for (int k = 0; k < 3; k++) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq; }
      }
  }
  #pragma unroll
  for (int k = 0; k < 3; k++) w[k] = A[k][k];
  #pragma unroll
  for (int a = 0; a < 2; a++)
    #pragma unroll
    for (int b = 0; b < 2 - a; b++)
      if (w[b] > w[b + 1]) { const double t = w[b]; w[b] = w[b + 1]; w[b + 1] = t; for (int k = 0; k < 3; k++) { const double v = V[k][b]; V[k][b] = V[k][b + 1]; V[k][b + 1] = v; } }
}

}  // namespace vils_eig
