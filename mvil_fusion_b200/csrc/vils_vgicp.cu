// vils_vgicp.cu — voxelised GICP scan matching: the producer of the LidarICPConstraint measurement (SURVEY.md §8f-3 ii).
// Reference: vils_estimator/src/estimator.cpp:263-303 builds fast_gicp::FastVGICP (resolution 0.5, DIRECT1, ADDITIVE voxels, PLANE
// regularisation, 20 neighbours), aligns scan j to scan i from a predicted guess and turns getFitnessScore() into the factor weight.
// Algorithm restated from the headers under vils_estimator/src/lidar_functions/fast_gicp/include/fast_gicp/gicp:
//   impl/fast_gicp_impl.hpp:240-298  calculate_covariances: k-NN, centred scatter / k, SVD, singular values replaced by (1, 1, 1e-3)
//   fast_vgicp_voxel.hpp:109-170     AdditiveGaussianVoxel + GaussianVoxelMap (coord = floor(x / res - 0.5), mean of points and covariances)
//   impl/fast_vgicp_impl.hpp:76-176  update_correspondences (voxel lookup, (C_B + T C_A T^T)^-1), linearize (w = sqrt(num_points),
//                                    J = [skew(T a), -I]), compute_error (correspondences and weights of the LAST linearize)
//   impl/lsq_registration_impl.hpp:52-166  computeTransformation / is_converged / step_lm; so3/so3.hpp:56-76 so3_exp
//   pcl::Registration::getFitnessScore: mean squared 1-NN distance of the transformed source in the target.
// B200 layout: clouds are float4 arrays in HBM.  Neighbour search is exhaustive over shared-memory tiles, one THREAD per query with its
// 20 best in registers (a 29 k-point scan is 227 CTAs; a 0.5 m grid comes next for larger clouds).  The voxel map is built without
// atomics on data and without a sort: the first point of every voxel (found by scanning the 64-bit voxel keys) sums its voxel's members
// in ascending point order, exactly the order of the reference's insertion loop, so the map is bit-reproducible; only the open-addressing
// table that maps key -> voxel uses atomicCAS.  linearize is one thread per source point, a fixed-shape block reduction of the 28 sums
// and a single-CTA pass over the block partials: deterministic as well.  The 6 x 6 LM step runs on the host between two launches.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/vils_cabi.h"
#include "common.h"
#include "small_eig.cuh"

namespace {

constexpr int VG_K = 20;          // FastGICP::k_correspondences_ (fast_gicp_impl.hpp:24); the reference never changes it
constexpr int VG_T = 128;         // threads per CTA of the scanning kernels
constexpr int VG_TILE = 1024;     // points per shared-memory tile
constexpr int VG_KT = 2048;       // voxel keys per shared-memory tile
constexpr int VG_RT = 256;        // threads per CTA of the linearize kernel
constexpr int VG_NS = 29;         // 21 (H lower triangle) + 6 (b) + 1 (error) + 1 (correspondence count)
constexpr unsigned long long VG_EMPTY = ~0ull;

struct Pose { double R[9]; double t[3]; };   // row-major rotation + translation of an Eigen::Isometry3d

__device__ __forceinline__ float sqdist(const float4& a, const float4& b) {
  // FLANN L2_Simple: ((dx^2 + dy^2) + dz^2) in float, no contraction
  const float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y), dz = __fsub_rn(a.z, b.z);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// calculate_covariances (fast_gicp_impl.hpp:240-298), RegularizationMethod::PLANE.  cov6 = xx xy xz yy yz zz of the regularised
// covariance U diag(1, 1, 1e-3) V^T; for the symmetric PSD scatter U = V, i.e. I - (1 - 1e-3) u3 u3^T with u3 the direction of least spread.
__global__ void __launch_bounds__(VG_T) vg_cov_kernel(const float4* __restrict__ pts, int n, double* __restrict__ cov6, int32_t* __restrict__ nn_out) {
  __shared__ float4 tile[VG_TILE];
  const int i = blockIdx.x * VG_T + threadIdx.x;
  const float4 q = pts[min(i, n - 1)];
  float d[VG_K]; int id[VG_K];
#pragma unroll
  for (int k = 0; k < VG_K; k++) { d[k] = FLT_MAX; id[k] = 0x7fffffff; }
  for (int base = 0; base < n; base += VG_TILE) {
    const int cnt = min(VG_TILE, n - base);
    __syncthreads();
    for (int k = threadIdx.x; k < cnt; k += VG_T) tile[k] = pts[base + k];
    __syncthreads();
    for (int j = 0; j < cnt; j++) {
      const float dist = sqdist(q, tile[j]);
      if (dist < d[VG_K - 1] || (dist == d[VG_K - 1] && base + j < id[VG_K - 1])) {
        d[VG_K - 1] = dist; id[VG_K - 1] = base + j;
#pragma unroll
        for (int k = VG_K - 1; k > 0; k--) {
          const bool sw = d[k] < d[k - 1] || (d[k] == d[k - 1] && id[k] < id[k - 1]);
          if (sw) { const float td = d[k]; d[k] = d[k - 1]; d[k - 1] = td; const int ti = id[k]; id[k] = id[k - 1]; id[k - 1] = ti; }
        }
      }
    }
  }
  if (i >= n) return;
  double m[3] = {0, 0, 0};
#pragma unroll
  for (int k = 0; k < VG_K; k++) { const float4 p = pts[id[k]]; m[0] += (double)p.x; m[1] += (double)p.y; m[2] += (double)p.z; }
  m[0] /= VG_K; m[1] /= VG_K; m[2] /= VG_K;
  double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
  for (int k = 0; k < VG_K; k++) {
    const float4 p = pts[id[k]];
    const double x = (double)p.x - m[0], y = (double)p.y - m[1], z = (double)p.z - m[2];
    C[0][0] += x * x; C[0][1] += x * y; C[0][2] += x * z; C[1][1] += y * y; C[1][2] += y * z; C[2][2] += z * z;
  }
  C[0][0] /= VG_K; C[0][1] /= VG_K; C[0][2] /= VG_K; C[1][1] /= VG_K; C[1][2] /= VG_K; C[2][2] /= VG_K;
  C[1][0] = C[0][1]; C[2][0] = C[0][2]; C[2][1] = C[1][2];
  double w[3], V[3][3];
  vils_eig::eig3(C, w, V);                                   // ascending: column 0 = least spread
  const double ux = V[0][0], uy = V[1][0], uz = V[2][0], s = 1.0 - 1e-3;
  double* o = cov6 + (size_t)6 * i;
  o[0] = 1.0 - s * ux * ux; o[1] = -s * ux * uy; o[2] = -s * ux * uz; o[3] = 1.0 - s * uy * uy; o[4] = -s * uy * uz; o[5] = 1.0 - s * uz * uz;
  if (nn_out) {
#pragma unroll
    for (int k = 0; k < VG_K; k++) nn_out[(size_t)VG_K * i + k] = id[k];
  }
}

// GaussianVoxelMap::voxel_coord (fast_vgicp_voxel.hpp:150-152): floor(x / res - 0.5) per axis, 21 bits each, biased
__device__ __forceinline__ unsigned long long pack_key(long long cx, long long cy, long long cz) {
  const long long B = 1 << 20, M = (1 << 21) - 1;
  cx = min(max(cx + B, 0ll), M); cy = min(max(cy + B, 0ll), M); cz = min(max(cz + B, 0ll), M);
  return ((unsigned long long)cx << 42) | ((unsigned long long)cy << 21) | (unsigned long long)cz;
}
__device__ __forceinline__ void voxel_coord(double x, double y, double z, double res, long long c[3]) {
  c[0] = (long long)floor(x / res - 0.5); c[1] = (long long)floor(y / res - 0.5); c[2] = (long long)floor(z / res - 0.5);
}
__device__ __forceinline__ unsigned int key_hash(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return (unsigned int)k;
}

__global__ void vg_key_kernel(const float4* __restrict__ pts, int n, double res, unsigned long long* __restrict__ keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  long long c[3]; voxel_coord((double)p.x, (double)p.y, (double)p.z, res, c);
  keys[i] = pack_key(c[0], c[1], c[2]);
}

// create_voxelmap (fast_vgicp_voxel.hpp:113-148), ADDITIVE.  Thread i scans the keys in ascending order: a match before i means i is not
// the first point of its voxel and it retires; otherwise it appends every member (itself included) in index order and finalises.
// vox: 10 doubles per voxel, stored at the index of the voxel's first point: mean(3) cov6(6) num_points.
__global__ void __launch_bounds__(VG_T) vg_voxel_kernel(const float4* __restrict__ pts, const double* __restrict__ cov6, const unsigned long long* __restrict__ keys, int n,
                                                        unsigned long long* __restrict__ tkeys, int32_t* __restrict__ tvals, unsigned int mask, double* __restrict__ vox,
                                                        int32_t* __restrict__ n_vox) {
  __shared__ unsigned long long tile[VG_KT];
  const int i = blockIdx.x * VG_T + threadIdx.x;
  bool alive = i < n;
  const unsigned long long key = alive ? keys[i] : 0ull;
  double m[3] = {0, 0, 0}, c[6] = {0, 0, 0, 0, 0, 0};
  int cnt = 0;
  for (int base = 0; base < n; base += VG_KT) {
    const int num = min(VG_KT, n - base);
    __syncthreads();
    for (int k = threadIdx.x; k < num; k += VG_T) tile[k] = keys[base + k];
    __syncthreads();
    if (!alive) continue;
    for (int j = 0; j < num; j++) {
      if (tile[j] != key) continue;
      const int g = base + j;
      if (g < i) { alive = false; break; }
      const float4 p = pts[g]; const double* cg = cov6 + (size_t)6 * g;
      m[0] += (double)p.x; m[1] += (double)p.y; m[2] += (double)p.z;
#pragma unroll
      for (int e = 0; e < 6; e++) c[e] += cg[e];
      cnt++;
    }
  }
  if (!alive) return;
  double* o = vox + (size_t)10 * i;
  o[0] = m[0] / cnt; o[1] = m[1] / cnt; o[2] = m[2] / cnt;
#pragma unroll
  for (int e = 0; e < 6; e++) o[3 + e] = c[e] / cnt;
  o[9] = (double)cnt;
  unsigned int slot = key_hash(key) & mask;
  while (true) {
    const unsigned long long old = atomicCAS(tkeys + slot, VG_EMPTY, key);
    if (old == VG_EMPTY) { tvals[slot] = i; break; }
    slot = (slot + 1) & mask;
  }
  atomicAdd(n_vox, 1);
}

__device__ __forceinline__ int voxel_lookup(const unsigned long long* __restrict__ tkeys, const int32_t* __restrict__ tvals, unsigned int mask, unsigned long long key) {
  unsigned int slot = key_hash(key) & mask;
  while (true) {
    const unsigned long long k = tkeys[slot];
    if (k == key) return tvals[slot];
    if (k == VG_EMPTY) return -1;
    slot = (slot + 1) & mask;
  }
}

__device__ __forceinline__ void inv3_sym(const double a[6], double o[3][3]) {
  // general 3x3 inverse by cofactors of the symmetric matrix xx xy xz yy yz zz
  const double xx = a[0], xy = a[1], xz = a[2], yy = a[3], yz = a[4], zz = a[5];
  const double c00 = yy * zz - yz * yz, c01 = xz * yz - xy * zz, c02 = xy * yz - xz * yy;
  const double det = xx * c00 + xy * c01 + xz * c02, r = 1.0 / det;
  o[0][0] = c00 * r; o[0][1] = o[1][0] = c01 * r; o[0][2] = o[2][0] = c02 * r;
  o[1][1] = (xx * zz - xz * xz) * r; o[1][2] = o[2][1] = (xy * xz - xx * yz) * r; o[2][2] = (xx * yy - xy * xy) * r;
}

// update_correspondences + linearize / compute_error (fast_vgicp_impl.hpp:76-176, 178-203).  T0: the pose the correspondences and the
// fused covariances were computed at (the last linearize), Ti: the pose the error is evaluated at (= T0 for linearize).
// n_off = 1 / 7 / 27 (NeighborSearchMethod).  part: gridDim.x x VG_NS block partials.
__global__ void __launch_bounds__(VG_RT) vg_linearize_kernel(const float4* __restrict__ src, const double* __restrict__ scov, int n, Pose T0, Pose Ti, double res, int n_off,
                                                             const unsigned long long* __restrict__ tkeys, const int32_t* __restrict__ tvals, unsigned int mask,
                                                             const double* __restrict__ vox, int with_h, double* __restrict__ part) {
  __shared__ double red[VG_RT / 32][VG_NS];
  const int i = blockIdx.x * VG_RT + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc[VG_NS];
#pragma unroll
  for (int e = 0; e < VG_NS; e++) acc[e] = 0.0;
  if (i < n) {
    const float4 pf = src[i];
    const double a[3] = {(double)pf.x, (double)pf.y, (double)pf.z};
    double p0[3], pi[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
      p0[r] = T0.R[3 * r] * a[0] + T0.R[3 * r + 1] * a[1] + T0.R[3 * r + 2] * a[2] + T0.t[r];
      pi[r] = Ti.R[3 * r] * a[0] + Ti.R[3 * r + 1] * a[1] + Ti.R[3 * r + 2] * a[2] + Ti.t[r];
    }
    long long c[3]; voxel_coord(p0[0], p0[1], p0[2], res, c);
    // R0 C_A R0^T (symmetric), computed once per source point
    const double* ca = scov + (size_t)6 * i;
    const double A[3][3] = {{ca[0], ca[1], ca[2]}, {ca[1], ca[3], ca[4]}, {ca[2], ca[4], ca[5]}};
    double RA[3][3], RAR[6];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int k = 0; k < 3; k++) RA[r][k] = T0.R[3 * r] * A[0][k] + T0.R[3 * r + 1] * A[1][k] + T0.R[3 * r + 2] * A[2][k];
    {
      int e = 0;
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int k = r; k < 3; k++) RAR[e++] = RA[r][0] * T0.R[3 * k] + RA[r][1] * T0.R[3 * k + 1] + RA[r][2] * T0.R[3 * k + 2];
    }
    for (int o = 0; o < n_off; o++) {
      int ox = 0, oy = 0, oz = 0;
      if (n_off == 7) { const int t7[7][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}}; ox = t7[o][0]; oy = t7[o][1]; oz = t7[o][2]; }
      else if (n_off == 27) { ox = o / 9 - 1; oy = (o / 3) % 3 - 1; oz = o % 3 - 1; }
      const int v = voxel_lookup(tkeys, tvals, mask, pack_key(c[0] + ox, c[1] + oy, c[2] + oz));
      if (v < 0) continue;
      const double* vb = vox + (size_t)10 * v;
      double rcr[6];
#pragma unroll
      for (int e = 0; e < 6; e++) rcr[e] = vb[3 + e] + RAR[e];
      double M[3][3]; inv3_sym(rcr, M);
      const double e3[3] = {vb[0] - pi[0], vb[1] - pi[1], vb[2] - pi[2]};
      const double w = sqrt(vb[9]);
      double Me[3];
#pragma unroll
      for (int r = 0; r < 3; r++) Me[r] = M[r][0] * e3[0] + M[r][1] * e3[1] + M[r][2] * e3[2];
      acc[27] += w * (e3[0] * Me[0] + e3[1] * Me[1] + e3[2] * Me[2]);
      acc[28] += 1.0;
      if (with_h) {
        // J = [skew(pi), -I] (3 x 6)
        const double J[3][6] = {{0.0, -pi[2], pi[1], -1.0, 0.0, 0.0}, {pi[2], 0.0, -pi[0], 0.0, -1.0, 0.0}, {-pi[1], pi[0], 0.0, 0.0, 0.0, -1.0}};
        double MJ[3][6];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
          for (int k = 0; k < 6; k++) MJ[r][k] = M[r][0] * J[0][k] + M[r][1] * J[1][k] + M[r][2] * J[2][k];
        int e = 0;
#pragma unroll
        for (int r = 0; r < 6; r++)
#pragma unroll
          for (int k = 0; k <= r; k++) acc[e++] += w * (J[0][r] * MJ[0][k] + J[1][r] * MJ[1][k] + J[2][r] * MJ[2][k]);
#pragma unroll
        for (int r = 0; r < 6; r++) acc[21 + r] += w * (J[0][r] * Me[0] + J[1][r] * Me[1] + J[2][r] * Me[2]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < VG_NS; e++) {
    double v = acc[e];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_down_sync(0xffffffffu, v, s);
    if (lane == 0) red[warp][e] = v;
  }
  __syncthreads();
  if (threadIdx.x < VG_NS) {
    double v = 0.0;
#pragma unroll
    for (int wv = 0; wv < VG_RT / 32; wv++) v += red[wv][threadIdx.x];
    part[(size_t)blockIdx.x * VG_NS + threadIdx.x] = v;
  }
}

// fixed-order sum of the block partials: one CTA, thread e owns column e
__global__ void vg_reduce_kernel(const double* __restrict__ part, int nblk, int ncol, double* __restrict__ out) {
  const int e = threadIdx.x;
  if (e >= ncol) return;
  double v = 0.0;
  for (int b = 0; b < nblk; b++) v += part[(size_t)b * ncol + e];
  out[e] = v;
}

// pcl::Registration::getFitnessScore: the source moved by the FLOAT final transformation, squared 1-NN distance in the target
__global__ void __launch_bounds__(VG_T) vg_fitness_kernel(const float4* __restrict__ src, int n, const float4* __restrict__ tgt, int m, const float* __restrict__ Tf /* 3x4 row-major */,
                                                          double* __restrict__ part) {
  __shared__ float4 tile[VG_TILE];
  __shared__ double red[VG_T / 32];
  const int i = blockIdx.x * VG_T + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float4 p = src[min(i, n - 1)];
  float4 q;
  // pcl::transformPointCloud: x * col0 + y * col1 + z * col2 + col3, left to right, float
  q.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(Tf[0], p.x), __fmul_rn(Tf[1], p.y)), __fmul_rn(Tf[2], p.z)), Tf[3]);
  q.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(Tf[4], p.x), __fmul_rn(Tf[5], p.y)), __fmul_rn(Tf[6], p.z)), Tf[7]);
  q.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(Tf[8], p.x), __fmul_rn(Tf[9], p.y)), __fmul_rn(Tf[10], p.z)), Tf[11]);
  q.w = 0.0f;
  float best = FLT_MAX;
  for (int base = 0; base < m; base += VG_TILE) {
    const int cnt = min(VG_TILE, m - base);
    __syncthreads();
    for (int k = threadIdx.x; k < cnt; k += VG_T) tile[k] = tgt[base + k];
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < cnt; j++) best = fminf(best, sqdist(q, tile[j]));
  }
  double v = i < n ? (double)best : 0.0;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_down_sync(0xffffffffu, v, s);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0.0; for (int wv = 0; wv < VG_T / 32; wv++) t += red[wv]; part[blockIdx.x] = t; }
}

// ---- host side ------------------------------------------------------------------------------------------------------------------------
struct Ctx {
  int n_src = 0, n_tgt = 0, n_off = 1, nblk = 0; unsigned int mask = 0; double res = 0.5;
  float4 *src = nullptr, *tgt = nullptr; double *scov = nullptr, *tcov = nullptr, *vox = nullptr, *part = nullptr, *out = nullptr, *fpart = nullptr;
  unsigned long long *keys = nullptr, *tkeys = nullptr; int32_t *tvals = nullptr, *nvox = nullptr; float* Tf = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  ~Ctx() {
    cudaFree(src); cudaFree(tgt); cudaFree(scov); cudaFree(tcov); cudaFree(vox); cudaFree(part); cudaFree(out); cudaFree(fpart); cudaFree(keys); cudaFree(tkeys);
    cudaFree(tvals); cudaFree(nvox); cudaFree(Tf);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
  }
};

#define VG_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

// uploads both clouds, computes the covariances and the target voxel map
cudaError_t vg_setup(Ctx& c, const float* src, int n_src, const float* tgt, int n_tgt, double res, int n_off) {
  c.n_src = n_src; c.n_tgt = n_tgt; c.res = res; c.n_off = n_off;
  unsigned int cap = 1024; while (cap < 2u * (unsigned int)n_tgt) cap <<= 1;
  c.mask = cap - 1; c.nblk = (n_src + VG_RT - 1) / VG_RT;
  const int fblk = (n_src + VG_T - 1) / VG_T;
  VG_TRY(cudaMalloc(&c.src, sizeof(float4) * (size_t)n_src)); VG_TRY(cudaMalloc(&c.tgt, sizeof(float4) * (size_t)n_tgt));
  VG_TRY(cudaMalloc(&c.scov, sizeof(double) * 6 * (size_t)n_src)); VG_TRY(cudaMalloc(&c.tcov, sizeof(double) * 6 * (size_t)n_tgt));
  VG_TRY(cudaMalloc(&c.vox, sizeof(double) * 10 * (size_t)n_tgt)); VG_TRY(cudaMalloc(&c.part, sizeof(double) * VG_NS * (size_t)c.nblk));
  VG_TRY(cudaMalloc(&c.out, sizeof(double) * 32)); VG_TRY(cudaMalloc(&c.fpart, sizeof(double) * (size_t)fblk));
  VG_TRY(cudaMalloc(&c.keys, sizeof(unsigned long long) * (size_t)n_tgt)); VG_TRY(cudaMalloc(&c.tkeys, sizeof(unsigned long long) * (size_t)cap));
  VG_TRY(cudaMalloc(&c.tvals, sizeof(int32_t) * (size_t)cap)); VG_TRY(cudaMalloc(&c.nvox, sizeof(int32_t))); VG_TRY(cudaMalloc(&c.Tf, sizeof(float) * 12));
  VG_TRY(cudaEventCreate(&c.e0)); VG_TRY(cudaEventCreate(&c.e1));
  VG_TRY(cudaMemcpy(c.src, src, sizeof(float4) * (size_t)n_src, cudaMemcpyHostToDevice));
  VG_TRY(cudaMemcpy(c.tgt, tgt, sizeof(float4) * (size_t)n_tgt, cudaMemcpyHostToDevice));
  cudaEventRecord(c.e0);
  vg_cov_kernel<<<(n_src + VG_T - 1) / VG_T, VG_T>>>(c.src, n_src, c.scov, nullptr);
  vg_cov_kernel<<<(n_tgt + VG_T - 1) / VG_T, VG_T>>>(c.tgt, n_tgt, c.tcov, nullptr);
  cudaMemsetAsync(c.tkeys, 0xff, sizeof(unsigned long long) * (size_t)cap);
  cudaMemsetAsync(c.nvox, 0, sizeof(int32_t));
  cudaMemsetAsync(c.vox, 0, sizeof(double) * 10 * (size_t)n_tgt);   // rows that are not a voxel's first point read as zero
  vg_key_kernel<<<(n_tgt + 255) / 256, 256>>>(c.tgt, n_tgt, res, c.keys);
  vg_voxel_kernel<<<(n_tgt + VG_T - 1) / VG_T, VG_T>>>(c.tgt, c.tcov, c.keys, n_tgt, c.tkeys, c.tvals, c.mask, c.vox, c.nvox);
  return cudaGetLastError();
}

// sums[0..20] H lower triangle (row-major), [21..26] b, [27] error, [28] correspondences
cudaError_t vg_linearize(Ctx& c, const Pose& T0, const Pose& Ti, bool with_h, double sums[VG_NS]) {
  vg_linearize_kernel<<<c.nblk, VG_RT>>>(c.src, c.scov, c.n_src, T0, Ti, c.res, c.n_off, c.tkeys, c.tvals, c.mask, c.vox, with_h ? 1 : 0, c.part);
  vg_reduce_kernel<<<1, 32>>>(c.part, c.nblk, VG_NS, c.out);
  return cudaMemcpy(sums, c.out, sizeof(double) * VG_NS, cudaMemcpyDeviceToHost);
}

Pose pose_from(const double T[16]) {
  Pose p;
  for (int r = 0; r < 3; r++) { for (int k = 0; k < 3; k++) p.R[3 * r + k] = T[4 * r + k]; p.t[r] = T[4 * r + 3]; }
  return p;
}
void unpack_h(const double s[VG_NS], double H[36], double b[6]) {
  int e = 0;
  for (int r = 0; r < 6; r++) for (int k = 0; k <= r; k++) { H[6 * r + k] = s[e]; H[6 * k + r] = s[e]; e++; }
  for (int r = 0; r < 6; r++) b[r] = s[21 + r];
}
// (H + lambda I) d = -b, symmetric 6 x 6, Gaussian elimination with partial pivoting (the reference: Eigen::LDLT)
bool solve6(const double H[36], double lambda, const double b[6], double d[6]) {
  double A[6][7];
  for (int r = 0; r < 6; r++) { for (int k = 0; k < 6; k++) A[r][k] = H[6 * r + k] + (r == k ? lambda : 0.0); A[r][6] = -b[r]; }
  for (int k = 0; k < 6; k++) {
    int p = k; for (int r = k + 1; r < 6; r++) if (std::fabs(A[r][k]) > std::fabs(A[p][k])) p = r;
    if (A[p][k] == 0.0 || !std::isfinite(A[p][k])) return false;
    if (p != k) for (int q = 0; q < 7; q++) std::swap(A[p][q], A[k][q]);
    for (int r = k + 1; r < 6; r++) { const double f = A[r][k] / A[k][k]; for (int q = k; q < 7; q++) A[r][q] -= f * A[k][q]; }
  }
  for (int k = 5; k >= 0; k--) { double s = A[k][6]; for (int q = k + 1; q < 6; q++) s -= A[k][q] * d[q]; d[k] = s / A[k][k]; }
  return true;
}
// delta = [so3_exp(d.head<3>()).toRotationMatrix(), d.tail<3>()] (so3.hpp:56-76 + Eigen's quaternion -> matrix)
Pose delta_pose(const double d[6]) {
  const double th2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  double im, re;
  if (th2 < 1e-10) { const double q4 = th2 * th2; im = 0.5 - 1.0 / 48.0 * th2 + 1.0 / 3840.0 * q4; re = 1.0 - 1.0 / 8.0 * th2 + 1.0 / 384.0 * q4; }
  else { const double th = std::sqrt(th2), h = 0.5 * th; im = std::sin(h) / th; re = std::cos(h); }
  const double w = re, x = im * d[0], y = im * d[1], z = im * d[2];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z, twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  Pose p;
  p.R[0] = 1 - (tyy + tzz); p.R[1] = txy - twz; p.R[2] = txz + twy;
  p.R[3] = txy + twz; p.R[4] = 1 - (txx + tzz); p.R[5] = tyz - twx;
  p.R[6] = txz - twy; p.R[7] = tyz + twx; p.R[8] = 1 - (txx + tyy);
  p.t[0] = d[3]; p.t[1] = d[4]; p.t[2] = d[5];
  return p;
}
Pose compose(const Pose& a, const Pose& b) {   // a * b
  Pose o;
  for (int r = 0; r < 3; r++) {
    for (int k = 0; k < 3; k++) o.R[3 * r + k] = a.R[3 * r] * b.R[k] + a.R[3 * r + 1] * b.R[3 + k] + a.R[3 * r + 2] * b.R[6 + k];
    o.t[r] = a.R[3 * r] * b.t[0] + a.R[3 * r + 1] * b.t[1] + a.R[3 * r + 2] * b.t[2] + a.t[r];
  }
  return o;
}
bool is_converged(const Pose& d, double rot_eps, double trans_eps) {   // lsq_registration_impl.hpp:76-86
  double m = 0.0;
  for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) m = std::max(m, std::fabs(d.R[3 * r + k] - (r == k ? 1.0 : 0.0)) / rot_eps);
  for (int r = 0; r < 3; r++) m = std::max(m, std::fabs(d.t[r]) / trans_eps);
  return m < 1.0;
}

int check_args(const char* who, const float* src, int n_src, const float* tgt, int n_tgt, const vils_vgicp_opts* o) {
  if (!src || !tgt || !o || n_src < 1 || n_tgt < VG_K || n_src < VG_K) return vils::fail(VILS_ERR_BAD_ARG, std::string(who) + ": null cloud or fewer points than k_correspondences (20)");
  if (o->k_correspondences != VG_K) return vils::fail(VILS_ERR_BAD_ARG, std::string(who) + ": k_correspondences must be 20 (FastGICP's default, the only value the reference uses)");
  if (o->neighbor_search != 1 && o->neighbor_search != 7 && o->neighbor_search != 27) return vils::fail(VILS_ERR_BAD_ARG, std::string(who) + ": neighbor_search must be 1, 7 or 27");
  if (!(o->resolution > 0.0) || o->max_iterations < 0 || o->lm_max_iterations < 0) return vils::fail(VILS_ERR_BAD_ARG, std::string(who) + ": bad option");
  return VILS_OK;
}

}  // namespace

extern "C" {

void vils_vgicp_default_opts(vils_vgicp_opts* o) {
  if (!o) return;
  o->resolution = 1.0;                 // FastVGICP() (fast_vgicp_impl.hpp:23); the estimator sets 0.5 (estimator.cpp:270)
  o->rotation_epsilon = 2e-3; o->transformation_epsilon = 5e-4; o->lm_init_lambda_factor = 1e-9;   // lsq_registration_impl.hpp:11-18
  o->k_correspondences = VG_K; o->neighbor_search = 1; o->max_iterations = 64; o->lm_max_iterations = 10; o->compute_fitness = 1; o->reserved = 0;
}

// calculate_covariances alone (parity entry point): cov6 = n x (xx xy xz yy yz zz), nn_idx (may be NULL) = n x 20 neighbour indices, nearest first
int vils_vgicp_covariances(const float* xyzi, int32_t n, double* cov6, int32_t* nn_idx, int32_t device) {
  if (!xyzi || !cov6 || n < VG_K) return vils::fail(VILS_ERR_BAD_ARG, "vils_vgicp_covariances: null pointer or fewer than 20 points");
  int st = vils::require_device(device); if (st) return st;
  float4* d_p = nullptr; double* d_c = nullptr; int32_t* d_n = nullptr;
  cudaError_t e = cudaMalloc(&d_p, sizeof(float4) * (size_t)n);
  if (e == cudaSuccess) e = cudaMalloc(&d_c, sizeof(double) * 6 * (size_t)n);
  if (e == cudaSuccess && nn_idx) e = cudaMalloc(&d_n, sizeof(int32_t) * VG_K * (size_t)n);
  if (e == cudaSuccess) e = cudaMemcpy(d_p, xyzi, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    vg_cov_kernel<<<(n + VG_T - 1) / VG_T, VG_T>>>(d_p, n, d_c, d_n);
    e = cudaMemcpy(cov6, d_c, sizeof(double) * 6 * (size_t)n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && nn_idx) e = cudaMemcpy(nn_idx, d_n, sizeof(int32_t) * VG_K * (size_t)n, cudaMemcpyDeviceToHost);
  }
  cudaFree(d_p); cudaFree(d_c); cudaFree(d_n);
  return e == cudaSuccess ? VILS_OK : vils::fail_cuda(e, "vils_vgicp_covariances");
}

// One FastVGICP::linearize at pose T (row-major 4 x 4): H (6 x 6 row-major), b (6), returns the error sum and the number of
// correspondences; voxels (may be NULL, capacity n_tgt x 10): mean(3) cov6(6) num_points per voxel, at the index of the voxel's first
// point, zero rows elsewhere (parity entry point for the voxel map).
int vils_vgicp_linearize(const float* src_xyzi, int32_t n_src, const float* tgt_xyzi, int32_t n_tgt, const double T[16], const vils_vgicp_opts* opts,
                         double* H, double* b, double* error, int32_t* n_corr, int32_t* n_voxels, double* voxels, int32_t device) {
  int st = check_args("vils_vgicp_linearize", src_xyzi, n_src, tgt_xyzi, n_tgt, opts); if (st) return st;
  if (!T || !H || !b || !error) return vils::fail(VILS_ERR_BAD_ARG, "vils_vgicp_linearize: null output");
  st = vils::require_device(device); if (st) return st;
  Ctx c;
  cudaError_t e = vg_setup(c, src_xyzi, n_src, tgt_xyzi, n_tgt, opts->resolution, opts->neighbor_search);
  double s[VG_NS];
  const Pose P = pose_from(T);
  if (e == cudaSuccess) e = vg_linearize(c, P, P, true, s);
  if (e == cudaSuccess && n_voxels) e = cudaMemcpy(n_voxels, c.nvox, sizeof(int32_t), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && voxels) e = cudaMemcpy(voxels, c.vox, sizeof(double) * 10 * (size_t)n_tgt, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return vils::fail_cuda(e, "vils_vgicp_linearize");
  unpack_h(s, H, b); *error = s[27]; if (n_corr) *n_corr = (int32_t)s[28];
  return VILS_OK;
}

// FastVGICP::align (estimator.cpp:269-297): LsqRegistration::computeTransformation with the LM stepper.  guess: row-major 4 x 4 (NULL =
// identity; the reference passes a float matrix, so callers round it to float first).  res->T is the final Isometry3d in double;
// final_transformation_ of the reference is its float cast, which is also what the fitness score uses.
int vils_vgicp_align(const float* src_xyzi, int32_t n_src, const float* tgt_xyzi, int32_t n_tgt, const double guess[16], const vils_vgicp_opts* opts,
                     vils_vgicp_result* res, int32_t device) {
  int st = check_args("vils_vgicp_align", src_xyzi, n_src, tgt_xyzi, n_tgt, opts); if (st) return st;
  if (!res) return vils::fail(VILS_ERR_BAD_ARG, "vils_vgicp_align: null result");
  st = vils::require_device(device); if (st) return st;
  Ctx c;
  cudaError_t e = vg_setup(c, src_xyzi, n_src, tgt_xyzi, n_tgt, opts->resolution, opts->neighbor_search);
  if (e != cudaSuccess) return vils::fail_cuda(e, "vils_vgicp_align (setup)");
  const double I4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  Pose x0 = pose_from(guess ? guess : I4);
  double lambda = -1.0, Hfin[36]; for (int k = 0; k < 36; k++) Hfin[k] = (k % 7 == 0) ? 1.0 : 0.0;   // final_hessian_.setIdentity()
  bool converged = false; int iters = 0, n_lin = 0; double last_err = 0.0, n_corr = 0.0;
  for (int it = 0; it < opts->max_iterations && !converged; it++) {
    iters = it;                                                    // nr_iterations_ = i
    // step_lm (lsq_registration_impl.hpp:123-166)
    double s[VG_NS], H[36], b[6];
    e = vg_linearize(c, x0, x0, true, s); n_lin++;
    if (e != cudaSuccess) return vils::fail_cuda(e, "vils_vgicp_align (linearize)");
    unpack_h(s, H, b); const double y0 = s[27]; last_err = y0; n_corr = s[28];
    if (!std::isfinite(y0)) return vils::fail(VILS_ERR_NOT_FINITE, "vils_vgicp_align: non-finite error sum");
    if (lambda < 0.0) { double m = 0.0; for (int k = 0; k < 6; k++) m = std::max(m, std::fabs(H[7 * k])); lambda = opts->lm_init_lambda_factor * m; }
    double nu = 2.0; bool stepped = false; Pose delta;
    for (int k = 0; k < opts->lm_max_iterations; k++) {
      double d[6];
      if (!solve6(H, lambda, b, d)) { lambda = nu * lambda; nu = 2 * nu; continue; }
      delta = delta_pose(d);
      const Pose xi = compose(delta, x0);
      double se[VG_NS];
      e = vg_linearize(c, x0, xi, false, se);
      if (e != cudaSuccess) return vils::fail_cuda(e, "vils_vgicp_align (compute_error)");
      const double yi = se[27];
      double den = 0.0; for (int q = 0; q < 6; q++) den += d[q] * (lambda * d[q] - b[q]);
      const double rho = (y0 - yi) / den;
      if (rho < 0) {
        if (is_converged(delta, opts->rotation_epsilon, opts->transformation_epsilon)) { stepped = true; break; }
        lambda = nu * lambda; nu = 2 * nu; continue;
      }
      x0 = xi; last_err = yi;
      lambda = lambda * std::max(1.0 / 3.0, 1 - std::pow(2 * rho - 1, 3));
      memcpy(Hfin, H, sizeof(Hfin));
      stepped = true; break;
    }
    if (!stepped) break;                                           // "lm not converged!!"
    converged = is_converged(delta, opts->rotation_epsilon, opts->transformation_epsilon);
  }
  memset(res, 0, sizeof(*res));
  for (int r = 0; r < 3; r++) { for (int k = 0; k < 3; k++) res->T[4 * r + k] = x0.R[3 * r + k]; res->T[4 * r + 3] = x0.t[r]; }
  res->T[15] = 1.0;
  memcpy(res->H, Hfin, sizeof(Hfin));
  res->error = last_err; res->iterations = iters; res->converged = converged ? 1 : 0; res->n_corr = (int32_t)n_corr; res->n_linearize = n_lin;
  e = cudaMemcpy(&res->n_voxels, c.nvox, sizeof(int32_t), cudaMemcpyDeviceToHost);
  res->fitness = -1.0;
  if (e == cudaSuccess && opts->compute_fitness) {
    float Tf[12];
    for (int r = 0; r < 3; r++) for (int k = 0; k < 4; k++) Tf[4 * r + k] = (float)res->T[4 * r + k];
    const int fblk = (n_src + VG_T - 1) / VG_T;
    e = cudaMemcpy(c.Tf, Tf, sizeof(Tf), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
      vg_fitness_kernel<<<fblk, VG_T>>>(c.src, n_src, c.tgt, n_tgt, c.Tf, c.fpart);
      vg_reduce_kernel<<<1, 32>>>(c.fpart, fblk, 1, c.out);
      double sum = 0.0;
      e = cudaMemcpy(&sum, c.out, sizeof(double), cudaMemcpyDeviceToHost);
      res->fitness = sum / n_src;
    }
  }
  if (e == cudaSuccess) { cudaEventRecord(c.e1); cudaEventSynchronize(c.e1); cudaEventElapsedTime(&res->elapsed_ms, c.e0, c.e1); }
  return e == cudaSuccess ? VILS_OK : vils::fail_cuda(e, "vils_vgicp_align");
}

}  // extern "C"
