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
// B200 layout: every cloud is binned into a dense grid whose cells ARE the reference's voxels (cell = floor(x / res - 0.5) - cmin):
// integer histogram -> single-CTA scan -> scatter -> rank inside the cell, which leaves the points cell-major (z fastest) and in ascending
// original order inside a cell.  That one structure serves (i) the 20-NN search (four lanes per query in cell order, so a warp walks the
// same cells; shells of cells are visited until the 20th distance is inside the searched block; a column of cells along z is one
// contiguous point range), (ii) the voxel map (one thread per cell sums its members in the reference's insertion order: bit-reproducible,
// no floating-point atomics), (iii) the voxel lookup of linearize (a bounds check and two loads) and (iv) the 1-NN of the fitness
// score.  linearize is one thread per source point, a fixed-shape block reduction of the 28 sums and a fixed-order pass over the block
// partials: deterministic as well.  The 6 x 6 LM step runs on the host between two launches.  (The first version searched
// exhaustively: 13.0 ms for a 28.8 k x 28.8 k pair, 10.6 ms of it in the two 20-NN passes.)
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/vils_cabi.h"
#include "common.h"
#include "small_eig.cuh"

namespace {

constexpr int VG_K = 20;          // FastGICP::k_correspondences_ (fast_gicp_impl.hpp:24); the reference never changes it
constexpr int VG_T = 128;         // threads per CTA of the search kernels
constexpr int VG_RT = 256;        // threads per CTA of the linearize kernel
constexpr int VG_NS = 29;         // 21 (H lower triangle) + 6 (b) + 1 (error) + 1 (correspondence count)
constexpr int VG_SCAN_T = 1024;
constexpr long long VG_MAX_CELLS = 16ll << 20;

struct Pose { double R[9]; double t[3]; };   // row-major rotation + translation of an Eigen::Isometry3d

// dense voxel grid over the bounding box of a cloud; cell (x, y, z) -> (x * dy + y) * dz + z
struct Grid {
  long long cx0, cy0, cz0;      // voxel coordinate of cell (0, 0, 0)
  int dx, dy, dz, ncell;
  double res;
  const int* start;             // ncell + 1: first sorted position of every cell
  const int* order;             // n: original point index at every sorted position (ascending inside a cell)
  const float4* spts;           // n: the points in sorted order, w = original index (bit pattern)
};

// GaussianVoxelMap::voxel_coord (fast_vgicp_voxel.hpp:150-152): floor(x / res - 0.5) per axis
__host__ __device__ inline void voxel_coord(double x, double y, double z, double res, long long c[3]) {
  c[0] = (long long)floor(x / res - 0.5); c[1] = (long long)floor(y / res - 0.5); c[2] = (long long)floor(z / res - 0.5);
}

__device__ __forceinline__ float sqdist(const float4& a, const float4& b) {
  // FLANN L2_Simple: ((dx^2 + dy^2) + dz^2) in float, no contraction
  const float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y), dz = __fsub_rn(a.z, b.z);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// ---- grid construction ------------------------------------------------------------------------------------------------------------------
__global__ void vg_cell_kernel(const float4* __restrict__ pts, int n, Grid G, int* __restrict__ cell_id, int* __restrict__ cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  long long c[3]; voxel_coord((double)p.x, (double)p.y, (double)p.z, G.res, c);
  const int cell = (int)(((c[0] - G.cx0) * G.dy + (c[1] - G.cy0)) * G.dz + (c[2] - G.cz0));
  cell_id[i] = cell;
  atomicAdd(cnt + cell, 1);
}

// exclusive scan of cnt[0..n) into start[0..n], one CTA: per-thread chunk sums, shared-memory scan of the 1024 sums, per-chunk prefix
__global__ void __launch_bounds__(VG_SCAN_T) vg_scan_kernel(const int* __restrict__ cnt, int n, int* __restrict__ start, int* __restrict__ cursor) {
  __shared__ int sums[VG_SCAN_T];
  const int t = threadIdx.x, chunk = (n + VG_SCAN_T - 1) / VG_SCAN_T, b = min(t * chunk, n), e = min(b + chunk, n);
  int s = 0;
  for (int k = b; k < e; k++) s += cnt[k];
  sums[t] = s;
  __syncthreads();
  for (int off = 1; off < VG_SCAN_T; off <<= 1) {
    const int v = t >= off ? sums[t - off] : 0;
    __syncthreads();
    sums[t] += v;
    __syncthreads();
  }
  int run = sums[t] - s;
  for (int k = b; k < e; k++) { start[k] = run; cursor[k] = run; run += cnt[k]; }
  if (t == VG_SCAN_T - 1) start[n] = sums[t];
}

__global__ void vg_scatter_kernel(const int* __restrict__ cell_id, int n, int* __restrict__ cursor, int* __restrict__ order) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  order[atomicAdd(cursor + cell_id[i], 1)] = i;
}

// The scatter leaves every cell's members in arrival order.  One thread per member: its rank among the members of its cell (a scan of
// the cell's short list) is its final slot, which puts every cell into ascending original order without a sort; the sorted point copy
// is written in the same pass.
__global__ void vg_rank_kernel(const int* __restrict__ start, const int* __restrict__ cell_id, const int* __restrict__ arrival, int n, const float4* __restrict__ pts,
                               int* __restrict__ order, float4* __restrict__ spts) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int v = arrival[k], c = cell_id[v], b = start[c], e = start[c + 1];
  int rank = 0;
  for (int j = b; j < e; j++) rank += arrival[j] < v;
  float4 p = pts[v]; p.w = __int_as_float(v);
  order[b + rank] = v; spts[b + rank] = p;
}

// ---- exact k-NN over the grid: shells of cells around the query's cell until the K-th best lies inside the searched block ---------------
// A list entry is one 64-bit key: the float bits of the (non-negative) squared distance above the point index, so that unsigned order
// is (distance, index) order and a compare or a select is two instructions instead of four.
typedef unsigned long long knn_key;
constexpr knn_key KNN_SENTINEL = (0x7f7fffffull << 32) | 0x7fffffffull;      // (FLT_MAX, INT_MAX)
__device__ __forceinline__ knn_key make_key(float dist, int idx) { return ((knn_key)__float_as_uint(dist) << 32) | (unsigned int)idx; }
__device__ __forceinline__ float key_dist(knn_key k) { return __uint_as_float((unsigned int)(k >> 32)); }
__device__ __forceinline__ int key_idx(knn_key k) { return (int)(unsigned int)(k & 0xffffffffull); }

// Branch-free insertion into the ascending list: K independent comparisons, then every slot takes its predecessor, the new entry or
// itself.  (A bubble of K - 1 dependent compare-and-swap steps, which the compiler turned into K - 1 branches, was 61 % of the k-NN
// kernel's samples.)  Precondition: v sorts before the last entry.
template <int K>
__device__ __forceinline__ void sorted_insert(knn_key (&key)[K], knn_key v) {
  bool g[K];
#pragma unroll
  for (int s = 0; s < K; s++) g[s] = key[s] > v;
#pragma unroll
  for (int s = K - 1; s > 0; s--) key[s] = g[s - 1] ? key[s - 1] : (g[s] ? v : key[s]);
  key[0] = g[0] ? v : key[0];
}
// L lanes share one query (L = 4 for the 20-NN pass, 1 for the fitness 1-NN): lane `sub` takes every L-th point of a range and keeps
// its own sorted K-list; the lists are merged at the end.
template <int K, int L>
__device__ __forceinline__ void knn_visit(const float4& q, const float4* __restrict__ spts, int b, int e, int sub, knn_key (&key)[K]) {
  for (int k0 = b + sub; k0 < e; k0 += 4 * L) {
    float4 pb[4];                                          // four loads in flight per lane
#pragma unroll
    for (int u = 0; u < 4; u++) pb[u] = spts[min(k0 + L * u, e - 1)];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (k0 + L * u >= e) break;
      const float4 p = pb[u];
      const knn_key v = make_key(sqdist(q, p), __float_as_int(p.w));
      if (v < key[K - 1]) sorted_insert<K>(key, v);
    }
  }
}
// (cx, cy, cz): the query's cell relative to the grid origin; it may lie outside the grid (fitness queries).  gmask: the L lanes of the
// group (their control flow is identical: the loops depend on the query only).
template <int K, int L>
__device__ void knn_grid(const Grid& G, const float4& q, long long cx, long long cy, long long cz, int sub, unsigned gmask, knn_key (&key)[K]) {
#pragma unroll
  for (int k = 0; k < K; k++) key[k] = KNN_SENTINEL;
  // first shell that can touch the grid
  long long r0 = 0;
  r0 = max(r0, max(-cx, cx - (G.dx - 1))); r0 = max(r0, max(-cy, cy - (G.dy - 1))); r0 = max(r0, max(-cz, cz - (G.dz - 1)));
  for (long long r = r0;; r++) {
    const long long xa = max(cx - r, 0ll), xb = min(cx + r, (long long)G.dx - 1), ya = max(cy - r, 0ll), yb = min(cy + r, (long long)G.dy - 1);
    const long long za = max(cz - r, 0ll), zb = min(cz + r, (long long)G.dz - 1);
    for (long long x = xa; x <= xb; x++)
      for (long long y = ya; y <= yb; y++) {
        const long long col = (x * G.dy + y) * G.dz;
        const bool edge = (x - cx == r) || (cx - x == r) || (y - cy == r) || (cy - y == r);
        int b0 = 0, e0 = 0, b1 = 0, e1 = 0;
        if (edge) {                                     // the whole column of the block: one contiguous point range
          if (za <= zb) { b0 = G.start[col + za]; e0 = G.start[col + zb + 1]; }
        } else {                                        // interior column: only the two caps of the shell
          const long long z1 = cz - r, z2 = cz + r;
          if (z1 >= 0 && z1 < G.dz) { b0 = G.start[col + z1]; e0 = G.start[col + z1 + 1]; }
          if (r > 0 && z2 >= 0 && z2 < G.dz) { b1 = G.start[col + z2]; e1 = G.start[col + z2 + 1]; }
        }
#pragma unroll 1
        for (int u = 0; u < 2; u++) knn_visit<K, L>(q, G.spts, u ? b1 : b0, u ? e1 : e0, sub, key);   // a single call site keeps the K-wide lists in registers
      }
    // every point outside the block [c - r, c + r]^3 is at least r * res away from a query inside cell c: done once K candidates are closer
    const float lim = (float)((double)r * G.res);
    const knn_key limkey = make_key(lim * lim * (1.0f - 1e-5f), 0);          // key < limkey  <=>  distance < the limit
    bool stop;
    if (L == 1) stop = key[K - 1] < limkey;
    else {
      int cnt = 0;
#pragma unroll
      for (int k = 0; k < K; k++) cnt += key[k] < limkey;
#pragma unroll
      for (int s = 1; s < L; s <<= 1) cnt += __shfl_xor_sync(gmask, cnt, s);
      stop = cnt >= K;
    }
    if (stop) break;
    if (cx - r <= 0 && cx + r >= G.dx - 1 && cy - r <= 0 && cy + r >= G.dy - 1 && cz - r <= 0 && cz + r >= G.dz - 1) break;   // whole grid searched
  }
}

// calculate_covariances (fast_gicp_impl.hpp:240-298), RegularizationMethod::PLANE.  cov6 = xx xy xz yy yz zz of the regularised
// covariance U diag(1, 1, 1e-3) V^T; for the symmetric PSD scatter U = V, i.e. I - (1 - 1e-3) u3 u3^T with u3 the direction of least spread.
// Four lanes per query, queries in sorted (cell-major) order; results are stored at the query's original index.
constexpr int VG_L = 4;
__global__ void __launch_bounds__(VG_T) vg_cov_kernel(Grid G, const float4* __restrict__ pts, int n, double* __restrict__ cov6, int32_t* __restrict__ nn_out) {
  const int t = blockIdx.x * VG_T + threadIdx.x, s = t / VG_L, sub = t % VG_L;
  if (s >= n) return;                                        // whole groups leave together
  const unsigned gmask = ((1u << VG_L) - 1u) << ((threadIdx.x & 31) & ~(VG_L - 1));
  const float4 q = G.spts[s];
  const int i = __float_as_int(q.w);
  long long c[3]; voxel_coord((double)q.x, (double)q.y, (double)q.z, G.res, c);
  knn_key key[VG_K];
  knn_grid<VG_K, VG_L>(G, q, c[0] - G.cx0, c[1] - G.cy0, c[2] - G.cz0, sub, gmask, key);
  // merge: VG_K rounds, the group's smallest head wins and its lane pops; neighbour k lands in lane k % VG_L
  int mine[VG_K / VG_L];
#pragma unroll
  for (int k = 0; k < VG_K; k++) {
    knn_key best = key[0];
#pragma unroll
    for (int x = 1; x < VG_L; x <<= 1) { const knn_key o = __shfl_xor_sync(gmask, best, x); best = o < best ? o : best; }
    if (key[0] == best) {
#pragma unroll
      for (int j = 0; j < VG_K - 1; j++) key[j] = key[j + 1];
      key[VG_K - 1] = KNN_SENTINEL;
    }
    if (k % VG_L == sub) mine[k / VG_L] = key_idx(best);
  }
  double m[3] = {0, 0, 0};
#pragma unroll
  for (int k = 0; k < VG_K / VG_L; k++) { const float4 p = pts[mine[k]]; m[0] += (double)p.x; m[1] += (double)p.y; m[2] += (double)p.z; }
#pragma unroll
  for (int a = 0; a < 3; a++) {
#pragma unroll
    for (int x = 1; x < VG_L; x <<= 1) m[a] += __shfl_xor_sync(gmask, m[a], x);
    m[a] /= VG_K;
  }
  double cs[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < VG_K / VG_L; k++) {
    const float4 p = pts[mine[k]];
    const double x = (double)p.x - m[0], y = (double)p.y - m[1], z = (double)p.z - m[2];
    cs[0] += x * x; cs[1] += x * y; cs[2] += x * z; cs[3] += y * y; cs[4] += y * z; cs[5] += z * z;
  }
#pragma unroll
  for (int a = 0; a < 6; a++) {
#pragma unroll
    for (int x = 1; x < VG_L; x <<= 1) cs[a] += __shfl_xor_sync(gmask, cs[a], x);
    cs[a] /= VG_K;
  }
  if (nn_out) {
#pragma unroll
    for (int k = 0; k < VG_K / VG_L; k++) nn_out[(size_t)VG_K * i + VG_L * k + sub] = mine[k];
  }
  if (sub != 0) return;
  double C[3][3] = {{cs[0], cs[1], cs[2]}, {cs[1], cs[3], cs[4]}, {cs[2], cs[4], cs[5]}};
  double w[3], V[3][3];
  vils_eig::eig3(C, w, V);                                   // ascending: column 0 = least spread
  const double ux = V[0][0], uy = V[1][0], uz = V[2][0], sc = 1.0 - 1e-3;
  double* o = cov6 + (size_t)6 * i;
  o[0] = 1.0 - sc * ux * ux; o[1] = -sc * ux * uy; o[2] = -sc * ux * uz; o[3] = 1.0 - sc * uy * uy; o[4] = -sc * uy * uz; o[5] = 1.0 - sc * uz * uz;
}

// create_voxelmap (fast_vgicp_voxel.hpp:113-148), ADDITIVE: one thread per cell appends its members in ascending point order (the
// order of the reference's insertion loop) and finalises.  vox: 10 doubles per voxel at the index of its first point: mean(3) cov6(6) num.
__global__ void vg_voxel_kernel(Grid G, const double* __restrict__ cov6, double* __restrict__ vox, int32_t* __restrict__ n_vox) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= G.ncell) return;
  const int b = G.start[c], e = G.start[c + 1];
  if (b == e) return;
  double m[3] = {0, 0, 0}, cv[6] = {0, 0, 0, 0, 0, 0};
  for (int k = b; k < e; k++) {
    const float4 p = G.spts[k]; const double* cg = cov6 + (size_t)6 * __float_as_int(p.w);
    m[0] += (double)p.x; m[1] += (double)p.y; m[2] += (double)p.z;
#pragma unroll
    for (int q = 0; q < 6; q++) cv[q] += cg[q];
  }
  const int cnt = e - b;
  double* o = vox + (size_t)10 * G.order[b];
  o[0] = m[0] / cnt; o[1] = m[1] / cnt; o[2] = m[2] / cnt;
#pragma unroll
  for (int q = 0; q < 6; q++) o[3 + q] = cv[q] / cnt;
  o[9] = (double)cnt;
  atomicAdd(n_vox, 1);
}

// lookup_voxel: row of `vox` for voxel coordinate (vx, vy, vz), -1 if the voxel is empty
__device__ __forceinline__ int voxel_lookup(const Grid& G, long long vx, long long vy, long long vz) {
  const long long x = vx - G.cx0, y = vy - G.cy0, z = vz - G.cz0;
  if (x < 0 || y < 0 || z < 0 || x >= G.dx || y >= G.dy || z >= G.dz) return -1;
  const long long cell = (x * G.dy + y) * G.dz + z;
  const int b = G.start[cell];
  return G.start[cell + 1] == b ? -1 : G.order[b];
}

__device__ __forceinline__ void inv3_sym(const double a[6], double o[3][3]) {
  // 3x3 inverse by cofactors of the symmetric matrix xx xy xz yy yz zz
  const double xx = a[0], xy = a[1], xz = a[2], yy = a[3], yz = a[4], zz = a[5];
  const double c00 = yy * zz - yz * yz, c01 = xz * yz - xy * zz, c02 = xy * yz - xz * yy;
  const double det = xx * c00 + xy * c01 + xz * c02, r = 1.0 / det;
  o[0][0] = c00 * r; o[0][1] = o[1][0] = c01 * r; o[0][2] = o[2][0] = c02 * r;
  o[1][1] = (xx * zz - xz * xz) * r; o[1][2] = o[2][1] = (xy * xz - xx * yz) * r; o[2][2] = (xx * yy - xy * xy) * r;
}

// update_correspondences + linearize / compute_error (fast_vgicp_impl.hpp:76-176, 178-203).  T0: the pose the correspondences and the
// fused covariances were computed at (the last linearize), Ti: the pose the error is evaluated at (= T0 for linearize).
// n_off = 1 / 7 / 27 (NeighborSearchMethod).  part: gridDim.x x VG_NS block partials.
__global__ void __launch_bounds__(VG_RT) vg_linearize_kernel(const float4* __restrict__ src, const double* __restrict__ scov, int n, Pose T0, Pose Ti, int n_off, Grid G,
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
    long long c[3]; voxel_coord(p0[0], p0[1], p0[2], G.res, c);
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
      const int v = voxel_lookup(G, c[0] + ox, c[1] + oy, c[2] + oz);
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

// fixed-order sum of the block partials: one CTA of 32 x 8 threads, thread (e, y) sums rows y, y + 8, ... of column e, then the 8 slices in order
__global__ void vg_reduce_kernel(const double* __restrict__ part, int nblk, int ncol, double* __restrict__ out) {
  __shared__ double sl[8][32];
  const int e = threadIdx.x, y = threadIdx.y;
  double v = 0.0;
  if (e < ncol) for (int b = y; b < nblk; b += 8) v += part[(size_t)b * ncol + e];
  sl[y][e] = v;
  __syncthreads();
  if (y == 0 && e < ncol) { double t = 0.0; for (int k = 0; k < 8; k++) t += sl[k][e]; out[e] = t; }
}

// pcl::Registration::getFitnessScore: the source moved by the FLOAT final transformation, squared 1-NN distance in the target
__global__ void __launch_bounds__(VG_T) vg_fitness_kernel(const float4* __restrict__ src, int n, Grid G, const float* __restrict__ Tf /* 3x4 row-major */, double* __restrict__ part) {
  __shared__ double red[VG_T / 32];
  const int t = blockIdx.x * VG_T + threadIdx.x, i = t / VG_L, sub = t % VG_L, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;   // four lanes per query
  double v = 0.0;
  if (i < n) {
    const unsigned gmask = ((1u << VG_L) - 1u) << (lane & ~(VG_L - 1));
    const float4 p = src[i];
    float4 q;
    // pcl::transformPointCloud: x * col0 + y * col1 + z * col2 + col3, left to right, float
    q.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(Tf[0], p.x), __fmul_rn(Tf[1], p.y)), __fmul_rn(Tf[2], p.z)), Tf[3]);
    q.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(Tf[4], p.x), __fmul_rn(Tf[5], p.y)), __fmul_rn(Tf[6], p.z)), Tf[7]);
    q.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(Tf[8], p.x), __fmul_rn(Tf[9], p.y)), __fmul_rn(Tf[10], p.z)), Tf[11]);
    q.w = 0.0f;
    long long c[3]; voxel_coord((double)q.x, (double)q.y, (double)q.z, G.res, c);
    knn_key key[1];
    knn_grid<1, VG_L>(G, q, c[0] - G.cx0, c[1] - G.cy0, c[2] - G.cz0, sub, gmask, key);
    knn_key best = key[0];
#pragma unroll
    for (int x = 1; x < VG_L; x <<= 1) { const knn_key o = __shfl_xor_sync(gmask, best, x); best = o < best ? o : best; }
    v = sub == 0 ? (double)key_dist(best) : 0.0;
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_down_sync(0xffffffffu, v, s);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0.0; for (int wv = 0; wv < VG_T / 32; wv++) t += red[wv]; part[blockIdx.x] = t; }
}

// ---- host side ------------------------------------------------------------------------------------------------------------------------
#define VG_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

// One device arena per host thread, grown on demand and kept between calls (a scan match allocates ~20 buffers; separate cudaMalloc /
// cudaFree pairs cost 7 ms per call against 1.1 ms of kernels), plus the two timing events.
struct Arena {
  char* base = nullptr; size_t cap = 0, off = 0; int device = -1;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaError_t reserve(size_t bytes, int dev) {
    if (dev != device) { if (base) cudaFree(base); base = nullptr; cap = 0; if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); e0 = e1 = nullptr; device = dev; }
    if (!e0) { VG_TRY(cudaEventCreate(&e0)); VG_TRY(cudaEventCreate(&e1)); }
    if (bytes > cap) {
      if (base) { cudaFree(base); base = nullptr; cap = 0; }
      const size_t want = bytes + bytes / 4;
      VG_TRY(cudaMalloc(&base, want));
      cap = want;
    }
    off = 0;
    return cudaSuccess;
  }
  template <class T> T* take(size_t n) { off = (off + 255) & ~(size_t)255; T* p = reinterpret_cast<T*>(base + off); off += n * sizeof(T); return p; }
  template <class T> static size_t need(size_t n) { return n * sizeof(T) + 256; }
};
thread_local Arena g_arena;

// a cloud on the device with its voxel grid
struct Cloud {
  int n = 0; Grid G{};
  float4 *pts = nullptr, *spts = nullptr; int *cell_id = nullptr, *cnt = nullptr, *start = nullptr, *cursor = nullptr, *order = nullptr, *arrival = nullptr; double* cov = nullptr;
  // bounding box in voxel coordinates on the host (the same IEEE arithmetic as the kernels)
  int prepare(const float* xyzi, int n_, double res) {
    n = n_;
    // floor(x / res - 0.5) is monotonic in x: the voxel box is the voxel of the coordinate-wise minimum and maximum (one pass of float
    // min / max instead of three double divisions per point, which alone cost 2 ms for two 28.8 k-point clouds)
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    bool bad = false;
    for (int i = 0; i < n; i++) {
      const float* p = xyzi + (size_t)4 * i;
      for (int a = 0; a < 3; a++) { bad |= !(std::fabs(p[a]) <= FLT_MAX); mn[a] = std::min(mn[a], p[a]); mx[a] = std::max(mx[a], p[a]); }
    }
    if (bad) return vils::fail(VILS_ERR_NOT_FINITE, "vils_vgicp: non-finite point (remove NaNs first, as estimator.cpp:236 does)");
    long long lo[3], hi[3];
    voxel_coord((double)mn[0], (double)mn[1], (double)mn[2], res, lo); voxel_coord((double)mx[0], (double)mx[1], (double)mx[2], res, hi);
    const long long dx = hi[0] - lo[0] + 1, dy = hi[1] - lo[1] + 1, dz = hi[2] - lo[2] + 1;
    if (dx > VG_MAX_CELLS || dy > VG_MAX_CELLS || dz > VG_MAX_CELLS || dx * dy > VG_MAX_CELLS || dx * dy * dz > VG_MAX_CELLS)
      return vils::fail(VILS_ERR_CAPACITY, "vils_vgicp: the voxel grid over the cloud's bounding box exceeds 16 M cells (resolution too fine for the extent)");
    G.cx0 = lo[0]; G.cy0 = lo[1]; G.cz0 = lo[2]; G.dx = (int)dx; G.dy = (int)dy; G.dz = (int)dz; G.ncell = (int)(dx * dy * dz); G.res = res;
    return VILS_OK;
  }
  size_t bytes() const {
    return 2 * Arena::need<float4>(n) + 3 * Arena::need<int>(n) + 2 * Arena::need<int>(G.ncell) + Arena::need<int>((size_t)G.ncell + 1) + Arena::need<double>(6 * (size_t)n);
  }
  cudaError_t build(Arena& A, const float* xyzi) {
    pts = A.take<float4>(n); spts = A.take<float4>(n); cell_id = A.take<int>(n); order = A.take<int>(n); arrival = A.take<int>(n);
    cnt = A.take<int>(G.ncell); start = A.take<int>((size_t)G.ncell + 1); cursor = A.take<int>(G.ncell); cov = A.take<double>(6 * (size_t)n);
    G.start = start; G.order = order; G.spts = spts;
    return cudaMemcpyAsync(pts, xyzi, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice);
  }
  void launch_grid_and_cov(int32_t* nn_out) {
    cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)G.ncell);
    vg_cell_kernel<<<(n + 255) / 256, 256>>>(pts, n, G, cell_id, cnt);
    vg_scan_kernel<<<1, VG_SCAN_T>>>(cnt, G.ncell, start, cursor);
    vg_scatter_kernel<<<(n + 255) / 256, 256>>>(cell_id, n, cursor, arrival);
    vg_rank_kernel<<<(n + 255) / 256, 256>>>(start, cell_id, arrival, n, pts, order, spts);
    vg_cov_kernel<<<(int)(((size_t)n * VG_L + VG_T - 1) / VG_T), VG_T>>>(G, pts, n, cov, nn_out);
  }
};

struct Ctx {
  int n_off = 1, nblk = 0;
  Cloud S, T;
  double *vox = nullptr, *part = nullptr, *out = nullptr, *fpart = nullptr; int32_t* nvox = nullptr; float* Tf = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
};

// uploads both clouds, builds their grids, computes the covariances and the target voxel map
cudaError_t vg_setup(Ctx& c, const float* src, const float* tgt, int n_off, int device) {
  c.n_off = n_off; c.nblk = (c.S.n + VG_RT - 1) / VG_RT;
  const int fblk = (int)(((size_t)c.S.n * VG_L + VG_T - 1) / VG_T);
  Arena& A = g_arena;
  VG_TRY(A.reserve(c.S.bytes() + c.T.bytes() + Arena::need<double>(10 * (size_t)c.T.n) + Arena::need<double>(VG_NS * (size_t)c.nblk) + Arena::need<double>(32) +
                   Arena::need<double>(fblk) + Arena::need<int32_t>(1) + Arena::need<float>(12), device));
  c.e0 = A.e0; c.e1 = A.e1;
  VG_TRY(c.S.build(A, src)); VG_TRY(c.T.build(A, tgt));
  c.vox = A.take<double>(10 * (size_t)c.T.n); c.part = A.take<double>(VG_NS * (size_t)c.nblk); c.out = A.take<double>(32); c.fpart = A.take<double>(fblk);
  c.nvox = A.take<int32_t>(1); c.Tf = A.take<float>(12);
  cudaEventRecord(c.e0);
  c.S.launch_grid_and_cov(nullptr);
  c.T.launch_grid_and_cov(nullptr);
  cudaMemsetAsync(c.nvox, 0, sizeof(int32_t));
  cudaMemsetAsync(c.vox, 0, sizeof(double) * 10 * (size_t)c.T.n);   // rows that are not a voxel's first point read as zero
  vg_voxel_kernel<<<(c.T.G.ncell + 255) / 256, 256>>>(c.T.G, c.T.cov, c.vox, c.nvox);
  return cudaGetLastError();
}

// sums[0..20] H lower triangle (row-major), [21..26] b, [27] error, [28] correspondences
cudaError_t vg_linearize(Ctx& c, const Pose& T0, const Pose& Ti, bool with_h, double sums[VG_NS]) {
  vg_linearize_kernel<<<c.nblk, VG_RT>>>(c.S.pts, c.S.cov, c.S.n, T0, Ti, c.n_off, c.T.G, c.vox, with_h ? 1 : 0, c.part);
  vg_reduce_kernel<<<1, dim3(32, 8)>>>(c.part, c.nblk, VG_NS, c.out);
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
  Cloud c; int32_t* d_n = nullptr;
  st = c.prepare(xyzi, n, 0.5); if (st) return st;                      // the search grid; any cell size gives the same neighbours
  Arena& A = g_arena;
  cudaError_t e = A.reserve(c.bytes() + Arena::need<int32_t>(VG_K * (size_t)n), device);
  if (e == cudaSuccess) e = c.build(A, xyzi);
  if (e == cudaSuccess) {
    if (nn_idx) d_n = A.take<int32_t>(VG_K * (size_t)n);
    c.launch_grid_and_cov(d_n);
    e = cudaMemcpy(cov6, c.cov, sizeof(double) * 6 * (size_t)n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && nn_idx) e = cudaMemcpy(nn_idx, d_n, sizeof(int32_t) * VG_K * (size_t)n, cudaMemcpyDeviceToHost);
  }
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
  st = c.S.prepare(src_xyzi, n_src, opts->resolution); if (st) return st;
  st = c.T.prepare(tgt_xyzi, n_tgt, opts->resolution); if (st) return st;
  cudaError_t e = vg_setup(c, src_xyzi, tgt_xyzi, opts->neighbor_search, device);
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
  st = c.S.prepare(src_xyzi, n_src, opts->resolution); if (st) return st;
  st = c.T.prepare(tgt_xyzi, n_tgt, opts->resolution); if (st) return st;
  cudaError_t e = vg_setup(c, src_xyzi, tgt_xyzi, opts->neighbor_search, device);
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
    const int fblk = (int)(((size_t)n_src * VG_L + VG_T - 1) / VG_T);
    e = cudaMemcpy(c.Tf, Tf, sizeof(Tf), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
      vg_fitness_kernel<<<fblk, VG_T>>>(c.S.spts, n_src, c.T.G, c.Tf, c.fpart);   // the source in its own cell order: a warp's queries walk the same target cells
      vg_reduce_kernel<<<1, dim3(32, 8)>>>(c.fpart, fblk, 1, c.out);
      double sum = 0.0;
      e = cudaMemcpy(&sum, c.out, sizeof(double), cudaMemcpyDeviceToHost);
      res->fitness = sum / n_src;
    }
  }
  if (e == cudaSuccess) { cudaEventRecord(c.e1); cudaEventSynchronize(c.e1); cudaEventElapsedTime(&res->elapsed_ms, c.e0, c.e1); }
  return e == cudaSuccess ? VILS_OK : vils::fail_cuda(e, "vils_vgicp_align");
}

}  // extern "C"
