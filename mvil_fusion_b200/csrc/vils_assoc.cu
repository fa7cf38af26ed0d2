// vils_assoc.cu — scan-to-map association: the producer of the LiDAR edge / plane factor arrays (SURVEY.md §8f-3 i).
// Reference: lidar_mapping/src/localMapping.cpp:611-766 — for every corner point: pointAssociateToMap, 5-NN in the corner map
// (pcl::KdTreeFLANN), if the 5th neighbour is within 1 m: mean + 3x3 scatter, SelfAdjointEigenSolver, line if l2 > 3 l1, end points
// a, b = centre +- 0.1 direction (:626-667); for every surface point: 10-NN, the 5 with the closest intensity, colPivHouseholderQr plane
// fit A n = -1, plane valid if every neighbour is within 0.2 m of it (:675-747).  The outputs are exactly the arrays
// vils_window.edge_{p,a,b} / plane_{p,n,d} take.
// The K nearest map points of a query come from a uniform 1 m hash grid over the map, built on the device per call (count / scan / scatter):
// one warp per query walks the 27 cells around it, every lane keeps its K best in registers, then a K-round shuffle merge.  A map point
// closer than 1 m to the query is always inside those 27 cells, so the result is EXACT whenever the K-th distance found is below 1 m — the
// only case in which the reference uses the neighbours at all (5th neighbour within 1 m, :626 / :679).  Queries whose K-th distance is not
// (sparse map regions) are re-run by the exhaustive kernel (the map streams through shared memory in 2048-point tiles shared by the
// CTA's 8 warps, threshold tightened after every tile), so every output, the reported neighbour indices of invalid points included, is
// what an exact k-d tree returns (ties broken by the smaller map index).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "../../include/vils_cabi.h"
#include "common.h"
#include "small_eig.cuh"

namespace {

constexpr int KMAX = 10;
constexpr int ASSOC_TILE = 2048;   // map points per shared-memory tile (32 KB)

struct Knn { float d[KMAX]; int i[KMAX]; };

template <int K>
__device__ __forceinline__ void knn_insert(Knn& h, float d, int idx) {
  if (!(d < h.d[K - 1] || (d == h.d[K - 1] && idx < h.i[K - 1]))) return;
  h.d[K - 1] = d; h.i[K - 1] = idx;
#pragma unroll
  for (int k = K - 1; k > 0; k--) {
    const bool sw = h.d[k] < h.d[k - 1] || (h.d[k] == h.d[k - 1] && h.i[k] < h.i[k - 1]);
    if (sw) { const float td = h.d[k]; h.d[k] = h.d[k - 1]; h.d[k - 1] = td; const int ti = h.i[k]; h.i[k] = h.i[k - 1]; h.i[k - 1] = ti; }
  }
}

// least squares A n = b (5 x 3) by Householder QR, FP64
__device__ void lstsq53(double A[5][3], double b[5], double n[3]) {
  #pragma unroll
  for (int k = 0; k < 3; k++) {
    double nr = 0; for (int i = k; i < 5; i++) nr += A[i][k] * A[i][k];
    nr = sqrt(nr);
    if (nr == 0) continue;
    const double alpha = A[k][k] > 0 ? -nr : nr;
    double v[5]; double vn = 0;
    #pragma unroll
    for (int i = k; i < 5; i++) { v[i] = A[i][k]; if (i == k) v[i] -= alpha; vn += v[i] * v[i]; }
    if (vn == 0) continue;
    #pragma unroll
    for (int j = k; j < 3; j++) { double s = 0; for (int i = k; i < 5; i++) s += v[i] * A[i][j]; s *= 2.0 / vn; for (int i = k; i < 5; i++) A[i][j] -= s * v[i]; }
    { double s = 0; for (int i = k; i < 5; i++) s += v[i] * b[i]; s *= 2.0 / vn; for (int i = k; i < 5; i++) b[i] -= s * v[i]; }
  }
  #pragma unroll
  for (int k = 2; k >= 0; k--) { double s = b[k]; for (int j = k + 1; j < 3; j++) s -= A[k][j] * n[j]; n[k] = s / A[k][k]; }
}

// pointAssociateToMap (:170-179): FP64 rotation, result stored in a float point
__device__ __forceinline__ void to_map(const float4 po, const double* __restrict__ qt, float& sx, float& sy, float& sz) {
  const double qx = qt[0], qy = qt[1], qz = qt[2], qw = qt[3];
  const double px = po.x, py = po.y, pz = po.z;
  const double tx2 = 2.0 * (qy * pz - qz * py), ty2 = 2.0 * (qz * px - qx * pz), tz2 = 2.0 * (qx * py - qy * px);
  sx = (float)(px + qw * tx2 + (qy * tz2 - qz * ty2) + qt[4]); sy = (float)(py + qw * ty2 + (qz * tx2 - qx * tz2) + qt[5]);
  sz = (float)(pz + qw * tz2 + (qx * ty2 - qy * tx2) + qt[6]);
}

// ---- uniform hash grid over the map: 1 m cells (the reference's validity radius), cell -> bucket by a multiplicative hash, points of a
// bucket contiguous after a counting sort; every sorted point carries its cell key, so a bucket shared by several cells is filtered exactly.
constexpr float GRID_CELL = 1.0f;
__device__ __forceinline__ int cell_of(float v) { const float c = floorf(v * (1.0f / GRID_CELL)); return (int)fminf(fmaxf(c, -1048575.0f), 1048575.0f); }
__device__ __forceinline__ unsigned long long cell_key(int ix, int iy, int iz) {
  return ((unsigned long long)(unsigned)(ix + 1048576) << 42) | ((unsigned long long)(unsigned)(iy + 1048576) << 21) | (unsigned long long)(unsigned)(iz + 1048576);
}
__device__ __forceinline__ unsigned cell_bucket(int ix, int iy, int iz, unsigned mask) {
  unsigned h = (unsigned)ix * 73856093u ^ (unsigned)iy * 19349663u ^ (unsigned)iz * 83492791u;
  h ^= h >> 15; h *= 2654435761u; h ^= h >> 13;
  return h & mask;
}

__global__ void grid_count_kernel(const float4* __restrict__ map, int n, unsigned mask, int* __restrict__ count, unsigned* __restrict__ bucket_of) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 m = map[i];
  const unsigned b = cell_bucket(cell_of(m.x), cell_of(m.y), cell_of(m.z), mask);
  bucket_of[i] = b;
  atomicAdd(count + b, 1);
}

// exclusive scan of `count` (nb = power of two >= 1024) by ONE block of 1024 threads, 1024 consecutive buckets per step (coalesced) with a
// running carry: start[b], start[nb] = n; `cursor` = copy of start
__global__ void __launch_bounds__(1024) grid_scan_kernel(const int* __restrict__ count, int nb, int* __restrict__ start, int* __restrict__ cursor) {
  __shared__ int wsum[32];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  int carry = 0;
  for (int base = 0; base < nb; base += 1024) {
    const int c = count[base + t];
    int v = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
    __syncthreads();                               // wsum of the previous step has been read
    if (lane == 31) wsum[warp] = v;
    __syncthreads();
    int w = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += u; }
    const int before = __shfl_sync(0xffffffffu, w, warp ? warp - 1 : 0), total = __shfl_sync(0xffffffffu, w, 31);
    const int excl = carry + (warp ? before : 0) + v - c;
    start[base + t] = excl; cursor[base + t] = excl;
    carry += total;
  }
  if (t == 0) start[nb] = carry;
}

// sorted copy of the map: x y z and, in w, the ORIGINAL index of the point (the intensity is read from the map itself by the fit)
__global__ void grid_scatter_kernel(const float4* __restrict__ map, int n, const unsigned* __restrict__ bucket_of, int* __restrict__ cursor, float4* __restrict__ spts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 m = map[i];
  spts[atomicAdd(cursor + bucket_of[i], 1)] = make_float4(m.x, m.y, m.z, __int_as_float(i));
}

// K-round merge of the lanes' sorted lists: every round the lane whose head is the global minimum pops it.  All lanes get the result.
template <int K>
__device__ __forceinline__ void knn_merge(Knn& h, float nd[K], int ni[K]) {
#pragma unroll
  for (int r = 0; r < K; r++) {
    float bd = h.d[0]; int bi = h.i[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, bd, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
    }
    nd[r] = bd; ni[r] = bi;
    if (h.i[0] == bi && h.d[0] == bd) {
#pragma unroll
      for (int k = 0; k < K - 1; k++) { h.d[k] = h.d[k + 1]; h.i[k] = h.i[k + 1]; }
      h.d[K - 1] = 3.0e38f; h.i[K - 1] = 0x7fffffff;
    }
  }
}

// line / plane fit of one query from its K nearest neighbours (one thread).  out: 10 doubles per query
//   corner: p(3) a(3) b(3) -      surface: p(3) n(3) d - - -
template <int K>
__device__ void finish_query(const float4* __restrict__ map, int n_map, const float4 po, int qi, int mode, const float nd[K], const int ni[K],
                             double* __restrict__ out, uint8_t* __restrict__ valid, int32_t* __restrict__ nn_out) {
  const double px = po.x, py = po.y, pz = po.z;
  double* o = out + (size_t)qi * 10;
  o[0] = px; o[1] = py; o[2] = pz;
  for (int k = 3; k < 10; k++) o[k] = 0.0;
  uint8_t ok = 0;
  int sel[5];
  if (n_map >= K && nd[4] < 1.0f) {
    if (mode == 0) {
      for (int j = 0; j < 5; j++) sel[j] = ni[j];
      double c[3] = {0, 0, 0}, P5[5][3];
      for (int j = 0; j < 5; j++) { const float4 m = map[sel[j]]; P5[j][0] = m.x; P5[j][1] = m.y; P5[j][2] = m.z; c[0] += m.x; c[1] += m.y; c[2] += m.z; }
      c[0] /= 5.0; c[1] /= 5.0; c[2] /= 5.0;
      double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      for (int j = 0; j < 5; j++) { const double d0 = P5[j][0] - c[0], d1 = P5[j][1] - c[1], d2 = P5[j][2] - c[2]; const double d[3] = {d0, d1, d2}; for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) C[a][b] += d[a] * d[b]; }
      double w[3], V[3][3]; vils_eig::eig3(C, w, V);
      if (w[2] > 3 * w[1]) {
        ok = 1;
        for (int a = 0; a < 3; a++) { o[3 + a] = 0.1 * V[a][2] + c[a]; o[6 + a] = -0.1 * V[a][2] + c[a]; }
      }
    } else {
      // the 5 of the 10 neighbours whose intensity is closest to the point's (std::sort of pair<float diff, index>, :680-692)
      float df[K]; int id[K];
      for (int j = 0; j < K; j++) { df[j] = fabsf(map[ni[j]].w - po.w); id[j] = ni[j]; }
      for (int a = 1; a < K; a++) { const float d = df[a]; const int ii = id[a]; int b = a - 1; while (b >= 0 && (df[b] > d || (df[b] == d && id[b] > ii))) { df[b + 1] = df[b]; id[b + 1] = id[b]; b--; } df[b + 1] = d; id[b + 1] = ii; }
      for (int j = 0; j < 5; j++) sel[j] = id[j];
      double A[5][3], A0[5][3], b[5], n[3] = {0, 0, 0};
      for (int j = 0; j < 5; j++) { const float4 m = map[sel[j]]; A[j][0] = A0[j][0] = m.x; A[j][1] = A0[j][1] = m.y; A[j][2] = A0[j][2] = m.z; b[j] = -1.0; }
      lstsq53(A, b, n);
      const double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
      const double d = 1.0 / nn; n[0] /= nn; n[1] /= nn; n[2] /= nn;
      ok = 1;
      for (int j = 0; j < 5; j++) if (fabs(n[0] * A0[j][0] + n[1] * A0[j][1] + n[2] * A0[j][2] + d) > 0.2) { ok = 0; break; }
      if (!(nn > 0) || !isfinite(d)) ok = 0;
      if (ok) { o[3] = n[0]; o[4] = n[1]; o[5] = n[2]; o[6] = d; }
    }
  } else {
    for (int j = 0; j < 5; j++) sel[j] = K > j ? ni[j] : -1;
  }
  valid[qi] = ok;
  if (nn_out) for (int j = 0; j < 5; j++) nn_out[(size_t)qi * 5 + j] = sel[j];
}

// Exhaustive search: of the queries qlist[0 .. *nq) (the fallback of the grid search), or of every query when qlist is null.  Writes the K
// nearest (squared distance, map index) of each query to knn_d / knn_i.
template <int K>
__global__ void __launch_bounds__(256) associate_kernel(const float4* __restrict__ map, int n_map, const float4* __restrict__ scan, int n_scan, const double* __restrict__ qt,
                                                        const int* __restrict__ qlist, const int* __restrict__ nq, float* __restrict__ knn_d, int* __restrict__ knn_i) {
  __shared__ float4 tile[ASSOC_TILE];
  const int lane = threadIdx.x & 31, gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int n_q = qlist ? *nq : n_scan;
  if (blockIdx.x * (blockDim.x >> 5) >= n_q) return;          // whole CTA past the list (block-uniform)
  const bool active = gw < n_q;
  const int qi = active ? (qlist ? qlist[gw] : gw) : (qlist ? qlist[n_q - 1] : n_scan - 1);
  const float4 po = scan[qi];
  float sx, sy, sz; to_map(po, qt, sx, sy, sz);
  Knn h;
#pragma unroll
  for (int k = 0; k < KMAX; k++) { h.d[k] = 3.0e38f; h.i[k] = 0x7fffffff; }
  float T = 3.0e38f;
  // the map streams through shared memory in tiles shared by the CTA's warps (8 queries per tile load: L2 traffic / 8)
  for (int base = 0; base < n_map; base += ASSOC_TILE) {
    const int cnt = min(ASSOC_TILE, n_map - base);
    __syncthreads();
    for (int j = threadIdx.x; j < cnt; j += blockDim.x) tile[j] = map[base + j];
    __syncthreads();
    if (active) {
      for (int j = lane; j < cnt; j += 32) {
        const float4 m = tile[j];
        const float dx = m.x - sx, dy = m.y - sy, dz = m.z - sz;
        const float d = dx * dx + dy * dy + dz * dz;
        if (d <= T) knn_insert<K>(h, d, base + j);           // T: the warp's current K-th best; almost every point fails this one compare
      }
      // tighten T to the K-th smallest over all lanes' lists (non-destructive K-round merge on a copy): per-lane lists alone give a
      // threshold 32x looser in quantile, and any lane that inserts makes the whole warp pay the insertion
      Knn hc = h;
      float kth = 3.0e38f;
#pragma unroll
      for (int r = 0; r < K; r++) {
        float bd = hc.d[0]; int bi = hc.i[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float od = __shfl_xor_sync(0xffffffffu, bd, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        kth = bd;
        if (hc.i[0] == bi && hc.d[0] == bd) {
#pragma unroll
          for (int k = 0; k < K - 1; k++) { hc.d[k] = hc.d[k + 1]; hc.i[k] = hc.i[k + 1]; }
          hc.d[K - 1] = 3.0e38f; hc.i[K - 1] = 0x7fffffff;
        }
      }
      T = kth;
    }
  }
  if (!active) return;
  float nd[K]; int ni[K];
  knn_merge<K>(h, nd, ni);
  if (lane < K) { float d = nd[0]; int i = ni[0];
#pragma unroll
    for (int k = 1; k < K; k++) if (lane == k) { d = nd[k]; i = ni[k]; }
    knn_d[(size_t)qi * K + lane] = d; knn_i[(size_t)qi * K + lane] = i; }
}

// Grid search: one warp per query.  Every map point closer than 1 m (2 m) to the query lies in the 3 x 3 x 3 (5 x 5 x 5) cells around it, so:
//   - K-th distance found in the 27 cells below 1 m: the K nearest are exact;
//   - 5th distance found not below 1 m: fewer than 5 map points within 1 m, the reference adds no residual (:626 / :679) and the neighbours are
//     never used: the query is settled as invalid with what was found;
//   - surface points (K = 10) with the 5th below 1 m but not the 10th: the shell out to 5 x 5 x 5 is searched too, exact if the 10th is below 2 m;
//   - what is left (5 to 9 neighbours within 2 m: the rim of the map) goes to qlist for the exhaustive kernel.
template <int K>
__device__ __forceinline__ void grid_walk(Knn& h, const int cx, const int cy, const int cz, const int R, const bool shell_only, const float sx, const float sy, const float sz,
                                          const unsigned mask, const int* __restrict__ start, const float4* __restrict__ spts, const int lane) {
  const int side = 2 * R + 1, ncell = side * side * side;
  for (int c0 = 0; c0 < ncell; c0 += 32) {
    // the lanes look up 32 cells' bucket ranges at once, then the warp walks them one after the other
    int b0 = 0, b1 = 0; unsigned long long key = 0;
    const int c = c0 + lane;
    if (c < ncell) {
      const int ox = c % side - R, oy = (c / side) % side - R, oz = c / (side * side) - R;
      if (!(shell_only && abs(ox) <= 1 && abs(oy) <= 1 && abs(oz) <= 1)) {
        const unsigned b = cell_bucket(cx + ox, cy + oy, cz + oz, mask);
        b0 = start[b]; b1 = start[b + 1]; key = cell_key(cx + ox, cy + oy, cz + oz);
      }
    }
    unsigned nonempty = __ballot_sync(0xffffffffu, b1 > b0);
    while (nonempty) {
      const int src = __ffs(nonempty) - 1; nonempty &= nonempty - 1;
      const int s0 = __shfl_sync(0xffffffffu, b0, src), s1 = __shfl_sync(0xffffffffu, b1, src);
      const unsigned long long kc = __shfl_sync(0xffffffffu, key, src);
      for (int j = s0 + lane; j < s1; j += 128) {           // four candidates per lane in flight
        float4 m[4];
#pragma unroll
        for (int u = 0; u < 4; u++) m[u] = j + 32 * u < s1 ? spts[j + 32 * u] : make_float4(3.0e18f, 3.0e18f, 3.0e18f, 0.0f);
#pragma unroll
        for (int u = 0; u < 4; u++) {
          if (cell_key(cell_of(m[u].x), cell_of(m[u].y), cell_of(m[u].z)) != kc) continue;   // another cell of the same bucket (or padding)
          const float dx = m[u].x - sx, dy = m[u].y - sy, dz = m[u].z - sz;
          const float d = dx * dx + dy * dy + dz * dz;
          if (d <= h.d[K - 1]) knn_insert<K>(h, d, __float_as_int(m[u].w));
        }
      }
    }
  }
}

template <int K>
__global__ void __launch_bounds__(256) associate_grid_kernel(int n_map, const float4* __restrict__ scan, int n_scan, const double* __restrict__ qt,
                                                             unsigned mask, const int* __restrict__ start, const float4* __restrict__ spts,
                                                             float* __restrict__ knn_d, int* __restrict__ knn_i, int* __restrict__ qlist, int* __restrict__ nq) {
  const int lane = threadIdx.x & 31, qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (qi >= n_scan) return;
  const float4 po = scan[qi];
  float sx, sy, sz; to_map(po, qt, sx, sy, sz);
  const int cx = cell_of(sx), cy = cell_of(sy), cz = cell_of(sz);
  Knn h;
#pragma unroll
  for (int k = 0; k < KMAX; k++) { h.d[k] = 3.0e38f; h.i[k] = 0x7fffffff; }
  grid_walk<K>(h, cx, cy, cz, 1, false, sx, sy, sz, mask, start, spts, lane);
  float nd[K]; int ni[K];
  { Knn hc = h; knn_merge<K>(hc, nd, ni); }
  const float r1 = GRID_CELL * GRID_CELL;
  bool settled = n_map < K || nd[K - 1] < r1 || !(nd[4] < r1);
  if (!settled) {                                   // K = 10 only: 5 to 9 points within 1 m
    grid_walk<K>(h, cx, cy, cz, 2, true, sx, sy, sz, mask, start, spts, lane);
    knn_merge<K>(h, nd, ni);
    settled = nd[K - 1] < 4.0f * r1;
  }
  if (!settled) { if (lane == 0) qlist[atomicAdd(nq, 1)] = qi; return; }
  if (lane < K) { float d = nd[0]; int i = ni[0];
#pragma unroll
    for (int k = 1; k < K; k++) if (lane == k) { d = nd[k]; i = ni[k]; }
    knn_d[(size_t)qi * K + lane] = d; knn_i[(size_t)qi * K + lane] = i; }
}

// Line / plane fit, one THREAD per query (as the tail of the search kernels it ran on one lane of every warp with 31 idle)
template <int K>
__global__ void __launch_bounds__(128) associate_fit_kernel(const float4* __restrict__ map, int n_map, const float4* __restrict__ scan, int n_scan, int mode,
                                                            const float* __restrict__ knn_d, const int* __restrict__ knn_i,
                                                            double* __restrict__ out, uint8_t* __restrict__ valid, int32_t* __restrict__ nn_out) {
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= n_scan) return;
  float nd[K]; int ni[K];
#pragma unroll
  for (int k = 0; k < K; k++) { nd[k] = knn_d[(size_t)qi * K + k]; ni[k] = knn_i[(size_t)qi * K + k]; }
  finish_query<K>(map, n_map, scan[qi], qi, mode, nd, ni, out, valid, nn_out);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// DepthRegister::get_depth (feature_tracker_/src/feature_tracker.h:98-343): LiDAR depth for the tracked features — the producer of the
// 8th feature channel (depth > 0 => lidar_depth_flag => constant inverse depth in the window solve, estimator.cpp:1217-1221).
//   A  depth_bin_kernel   : cloud -> camera-aligned LiDAR frame (two float affine transforms, as pcl::transformPointCloud applies them),
//                           view filter, 0.5 degree range image; the CLOSEST point of every bin wins (:143-168), here by a 64-bit atomicMin on
//                           (range bits, point index) — same winner as the sequential `dist < rangeImage` scan, ties to the first point.
//   B  depth_compact_kernel: occupied bins -> dense list of unit-sphere points with their range (:241-250)
//   C  depth_feature_kernel: per feature (one warp) the 3 nearest unit-sphere points (:258-262, exhaustive instead of a kd-tree), distance
//                           gate, range spread <= 2 m, depth = mean range * feature.x, kept if > 3 m (:263-337)
// ---------------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 affine(const float* T, float4 p) {   // row-major 3x4
  return make_float4(T[0] * p.x + T[1] * p.y + T[2] * p.z + T[3], T[4] * p.x + T[5] * p.y + T[6] * p.z + T[7], T[8] * p.x + T[9] * p.y + T[10] * p.z + T[11], p.w);
}
__global__ void depth_bin_kernel(const float4* __restrict__ cloud, int n, const float* __restrict__ T /* two 3x4 */, int nb, unsigned long long* __restrict__ bins,
                                 float4* __restrict__ local) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = affine(T + 12, affine(T, cloud[i]));
  local[i] = p;
  if (p.x < 0 || fabsf(p.y / p.x) > 10 || fabsf(p.z / p.x) > 10) return;
  const float bin_res = 180.0f / (float)nb;
  const float row_angle = (float)((double)atan2f(p.z, sqrtf(p.x * p.x + p.y * p.y)) * 180.0 / M_PI + 90.0);
  const int row_id = (int)roundf(row_angle / bin_res);
  const float col_angle = (float)((double)atan2f(p.x, p.y) * 180.0 / M_PI);
  const int col_id = (int)roundf(col_angle / bin_res);
  if (row_id < 0 || row_id >= nb || col_id < 0 || col_id >= nb) return;
  const float dist = sqrtf(p.x * p.x + p.y * p.y + p.z * p.z);
  atomicMin(&bins[(size_t)row_id * nb + col_id], ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned int)i);
}
__global__ void depth_compact_kernel(const unsigned long long* __restrict__ bins, int nbins, const float4* __restrict__ local, float4* __restrict__ sphere, int* __restrict__ count) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbins) return;
  const unsigned long long v = bins[b];
  if (v == ~0ull) return;
  float4 p = local[(unsigned int)(v & 0xffffffffu)];
  const float range = sqrtf(p.x * p.x + p.y * p.y + p.z * p.z);
  p.x /= range; p.y /= range; p.z /= range; p.w = range;
  sphere[atomicAdd(count, 1)] = p;
}
__global__ void __launch_bounds__(256) depth_feature_kernel(const float4* __restrict__ sphere, const int* __restrict__ count, const float* __restrict__ feat, int m, int nb,
                                                            float* __restrict__ depth) {
  __shared__ float4 tile[ASSOC_TILE];
  const int lane = threadIdx.x & 31, fi0 = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const bool active = fi0 < m;
  const int fi = active ? fi0 : m - 1, n = *count;
  // feature on the unit sphere, camera -> LiDAR axis convention (:226-238)
  float fx = feat[3 * fi], fy = feat[3 * fi + 1], fz = feat[3 * fi + 2];
  const float nn = sqrtf(fx * fx + fy * fy + fz * fz); fx /= nn; fy /= nn; fz /= nn;
  const float px = fz, py = -fx, pz = -fy;
  Knn h;
#pragma unroll
  for (int k = 0; k < KMAX; k++) { h.d[k] = 3.0e38f; h.i[k] = 0x7fffffff; }
  for (int base = 0; base < n; base += ASSOC_TILE) {
    const int cnt = min(ASSOC_TILE, n - base);
    __syncthreads();
    for (int j = threadIdx.x; j < cnt; j += blockDim.x) tile[j] = sphere[base + j];
    __syncthreads();
    if (active)
      for (int j = lane; j < cnt; j += 32) {
        const float4 q = tile[j];
        const float dx = q.x - px, dy = q.y - py, dz = q.z - pz;
        knn_insert<3>(h, dx * dx + dy * dy + dz * dz, base + j);
      }
  }
  if (!active) return;
  float nd[3]; int ni[3];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    float bd = h.d[0]; int bi = h.i[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, bd, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
    }
    nd[r] = bd; ni[r] = bi;
    if (h.i[0] == bi && h.d[0] == bd) { h.d[0] = h.d[1]; h.i[0] = h.i[1]; h.d[1] = h.d[2]; h.i[1] = h.i[2]; h.d[2] = 3.0e38f; h.i[2] = 0x7fffffff; }
  }
  if (lane != 0) return;
  float out = -1.0f;
  const float bin_res = 180.0f / (float)nb;
  const float thr = (float)pow(sin((double)bin_res / 180.0 * M_PI) * 5.0, 2);
  if (n >= 10 && ni[2] != 0x7fffffff && nd[2] < thr) {
    const float r1 = sphere[ni[0]].w, r2 = sphere[ni[1]].w, r3 = sphere[ni[2]].w;
    const float mn = fminf(r1, fminf(r2, r3)), mx = fmaxf(r1, fmaxf(r2, r3));
    if (!(mx - mn > 2)) {
      const float s_ = (r1 + r2 + r3) / 3;
      const float d = px * s_;                       // depth for the z-normalised feature (LiDAR x = camera z)
      if (d > 3.0f) out = d;
    }
  }
  depth[fi] = out;
}

}  // namespace

extern "C" {

// DepthRegister::get_depth (feature_tracker_/src/feature_tracker.h:98-343).  cloud: n points x y z intensity (4 packed floats) of the
// stacked depth cloud in the world frame; T1, T2: the two 3x4 row-major float affine transforms the reference applies one after the
// other (transNow.inverse(), then Tlc_ * TransFormLC.inverse());  feat: m undistorted features (x, y, 1);  depth[i] = LiDAR depth of
// feature i along the camera z axis, or -1.  num_bins = 360 in the reference.
int vils_depth_register(const float* cloud_xyzi, int32_t n, const float T1[12], const float T2[12], int32_t num_bins, const float* feat_xyz, int32_t m,
                        float* depth, float* ms, int32_t device) {
  if (n < 0 || m < 0 || (n && !cloud_xyzi) || !T1 || !T2 || num_bins <= 0 || num_bins > 4096 || (m && (!feat_xyz || !depth))) return vils::fail(VILS_ERR_BAD_ARG, "vils_depth_register: bad argument");
  for (int i = 0; i < m; i++) depth[i] = -1.0f;
  if (m == 0 || n == 0) return VILS_OK;
  int st = vils::require_device(device); if (st) return st;
  const size_t nbins = (size_t)num_bins * num_bins;
  float4* d_cloud = nullptr; float4* d_local = nullptr; float4* d_sphere = nullptr; unsigned long long* d_bins = nullptr; float* d_T = nullptr; float* d_feat = nullptr; float* d_depth = nullptr; int* d_cnt = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaError_t e = cudaMalloc(&d_cloud, sizeof(float4) * (size_t)n);
  if (e == cudaSuccess) e = cudaMalloc(&d_local, sizeof(float4) * (size_t)n);
  if (e == cudaSuccess) e = cudaMalloc(&d_sphere, sizeof(float4) * std::min(nbins, (size_t)n));
  if (e == cudaSuccess) e = cudaMalloc(&d_bins, sizeof(unsigned long long) * nbins);
  if (e == cudaSuccess) e = cudaMalloc(&d_T, sizeof(float) * 24);
  if (e == cudaSuccess) e = cudaMalloc(&d_feat, sizeof(float) * 3 * (size_t)m);
  if (e == cudaSuccess) e = cudaMalloc(&d_depth, sizeof(float) * (size_t)m);
  if (e == cudaSuccess) e = cudaMalloc(&d_cnt, sizeof(int));
  if (e == cudaSuccess) e = cudaEventCreate(&e0);
  if (e == cudaSuccess) e = cudaEventCreate(&e1);
  if (e == cudaSuccess) {
    float T[24]; memcpy(T, T1, sizeof(float) * 12); memcpy(T + 12, T2, sizeof(float) * 12);
    cudaMemcpy(d_cloud, cloud_xyzi, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice);
    cudaMemcpy(d_T, T, sizeof(T), cudaMemcpyHostToDevice);
    cudaMemcpy(d_feat, feat_xyz, sizeof(float) * 3 * (size_t)m, cudaMemcpyHostToDevice);
    cudaEventRecord(e0);
    cudaMemsetAsync(d_bins, 0xff, sizeof(unsigned long long) * nbins);
    cudaMemsetAsync(d_cnt, 0, sizeof(int));
    depth_bin_kernel<<<(n + 255) / 256, 256>>>(d_cloud, n, d_T, num_bins, d_bins, d_local);
    depth_compact_kernel<<<(int)((nbins + 255) / 256), 256>>>(d_bins, (int)nbins, d_local, d_sphere, d_cnt);
    depth_feature_kernel<<<(m + 7) / 8, 256>>>(d_sphere, d_cnt, d_feat, m, num_bins, d_depth);
    cudaEventRecord(e1);
    e = cudaMemcpy(depth, d_depth, sizeof(float) * (size_t)m, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && ms) cudaEventElapsedTime(ms, e0, e1);
  }
  cudaFree(d_cloud); cudaFree(d_local); cudaFree(d_sphere); cudaFree(d_bins); cudaFree(d_T); cudaFree(d_feat); cudaFree(d_depth); cudaFree(d_cnt);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  return e == cudaSuccess ? VILS_OK : vils::fail_cuda(e, "vils_depth_register");
}


// Scan-to-map association (localMapping.cpp:611-766).  map / scan: PCL-style x y z intensity as 4 packed floats per point (HOST pointers);
// q_w_curr (x y z w), t_w_curr: the current scan-to-map pose.  mode 0: corner points -> LidarEdgeFactor inputs, out[i] = p(3) a(3) b(3) -;
// mode 1: surface points -> LidarPlaneNormFactor inputs, out[i] = p(3) n(3) d - - -.  valid[i] = 1 where the reference adds a residual
// block.  nn_idx (may be NULL): the 5 map indices used per point.  ms (may be NULL): device time of the kernel.
int vils_lidar_associate(const float* map_xyzi, int32_t n_map, const float* scan_xyzi, int32_t n_scan, const double q_w_curr[4], const double t_w_curr[3],
                         int32_t mode, double* out, uint8_t* valid, int32_t* nn_idx, float* ms, int32_t device) {
  if (n_map < 0 || n_scan < 0 || (n_map && !map_xyzi) || (n_scan && (!scan_xyzi || !out || !valid)) || !q_w_curr || !t_w_curr || (mode != 0 && mode != 1))
    return vils::fail(VILS_ERR_BAD_ARG, "vils_lidar_associate: bad argument");
  if (n_scan == 0) return VILS_OK;
  int st = vils::require_device(device); if (st) return st;
  float4* d_map = nullptr; float4* d_scan = nullptr; double* d_qt = nullptr; double* d_out = nullptr; uint8_t* d_valid = nullptr; int32_t* d_nn = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaError_t e = cudaMalloc(&d_map, sizeof(float4) * (size_t)(n_map > 0 ? n_map : 1));
  if (e == cudaSuccess) e = cudaMalloc(&d_scan, sizeof(float4) * (size_t)n_scan);
  if (e == cudaSuccess) e = cudaMalloc(&d_qt, sizeof(double) * 8);
  if (e == cudaSuccess) e = cudaMalloc(&d_out, sizeof(double) * 10 * (size_t)n_scan);
  if (e == cudaSuccess) e = cudaMalloc(&d_valid, (size_t)n_scan);
  if (e == cudaSuccess) e = cudaMalloc(&d_nn, sizeof(int32_t) * 5 * (size_t)n_scan);
  float* d_knn_d = nullptr; int* d_knn_i = nullptr;
  if (e == cudaSuccess) e = cudaMalloc(&d_knn_d, sizeof(float) * KMAX * (size_t)n_scan);
  if (e == cudaSuccess) e = cudaMalloc(&d_knn_i, sizeof(int) * KMAX * (size_t)n_scan);
  // hash grid over the map: nb buckets (power of two, about one per 16 points — a 1 m cell of a LiDAR map holds tens of points —, at least
  // 1024), all arrays in one allocation
  int nb = 1024; while (nb < n_map / 16) nb <<= 1;
  uint8_t* d_grid = nullptr; int* g_count = nullptr; int* g_start = nullptr; int* g_cursor = nullptr; unsigned* g_bucket = nullptr;
  float4* g_pts = nullptr; int* g_list = nullptr; int* g_nq = nullptr;
  if (e == cudaSuccess && n_map >= 4096) {
    const size_t nm = (size_t)n_map;
    size_t off = 0; auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_pts = take(16 * nm), o_bkt = take(4 * nm), o_cnt = take(4 * (size_t)nb),
                 o_start = take(4 * ((size_t)nb + 1)), o_cur = take(4 * (size_t)nb), o_list = take(4 * (size_t)n_scan), o_nq = take(4);
    e = cudaMalloc(&d_grid, off);
    if (e == cudaSuccess) {
      g_pts = reinterpret_cast<float4*>(d_grid + o_pts);
      g_bucket = reinterpret_cast<unsigned*>(d_grid + o_bkt); g_count = reinterpret_cast<int*>(d_grid + o_cnt); g_start = reinterpret_cast<int*>(d_grid + o_start);
      g_cursor = reinterpret_cast<int*>(d_grid + o_cur); g_list = reinterpret_cast<int*>(d_grid + o_list); g_nq = reinterpret_cast<int*>(d_grid + o_nq);
    }
  }
  if (e == cudaSuccess) e = cudaEventCreate(&e0);
  if (e == cudaSuccess) e = cudaEventCreate(&e1);
  if (e == cudaSuccess) {
    double qt[8] = {q_w_curr[0], q_w_curr[1], q_w_curr[2], q_w_curr[3], t_w_curr[0], t_w_curr[1], t_w_curr[2], 0.0};
    if (n_map) cudaMemcpy(d_map, map_xyzi, sizeof(float4) * (size_t)n_map, cudaMemcpyHostToDevice);
    cudaMemcpy(d_scan, scan_xyzi, sizeof(float4) * (size_t)n_scan, cudaMemcpyHostToDevice);
    cudaMemcpy(d_qt, qt, sizeof(qt), cudaMemcpyHostToDevice);
    cudaEventRecord(e0);
    const int warps = 8, grid = (n_scan + warps - 1) / warps;
    static const bool no_grid = getenv("VILS_ASSOC_GRID") && atoi(getenv("VILS_ASSOC_GRID")) == 0;   // experiments: exhaustive search only
    const bool use_grid = !no_grid && n_map >= 4096 && e == cudaSuccess && d_grid;
    if (use_grid) {
      cudaMemsetAsync(g_count, 0, sizeof(int) * (size_t)nb);
      cudaMemsetAsync(g_nq, 0, sizeof(int));
      grid_count_kernel<<<(n_map + 255) / 256, 256>>>(d_map, n_map, (unsigned)(nb - 1), g_count, g_bucket);
      grid_scan_kernel<<<1, 1024>>>(g_count, nb, g_start, g_cursor);
      grid_scatter_kernel<<<(n_map + 255) / 256, 256>>>(d_map, n_map, g_bucket, g_cursor, g_pts);
      if (mode == 0) associate_grid_kernel<5><<<grid, 32 * warps>>>(n_map, d_scan, n_scan, d_qt, (unsigned)(nb - 1), g_start, g_pts, d_knn_d, d_knn_i, g_list, g_nq);
      else associate_grid_kernel<10><<<grid, 32 * warps>>>(n_map, d_scan, n_scan, d_qt, (unsigned)(nb - 1), g_start, g_pts, d_knn_d, d_knn_i, g_list, g_nq);
    }
    // exhaustive search of the queries the grid could not settle (CTAs past the list exit at once), or of all queries without a grid
    const int* qlist = use_grid ? g_list : nullptr;
    if (mode == 0) {
      associate_kernel<5><<<grid, 32 * warps>>>(d_map, n_map, d_scan, n_scan, d_qt, qlist, g_nq, d_knn_d, d_knn_i);
      associate_fit_kernel<5><<<(n_scan + 127) / 128, 128>>>(d_map, n_map, d_scan, n_scan, 0, d_knn_d, d_knn_i, d_out, d_valid, d_nn);
    } else {
      associate_kernel<10><<<grid, 32 * warps>>>(d_map, n_map, d_scan, n_scan, d_qt, qlist, g_nq, d_knn_d, d_knn_i);
      associate_fit_kernel<10><<<(n_scan + 127) / 128, 128>>>(d_map, n_map, d_scan, n_scan, 1, d_knn_d, d_knn_i, d_out, d_valid, d_nn);
    }
    cudaEventRecord(e1);
    e = cudaMemcpy(out, d_out, sizeof(double) * 10 * (size_t)n_scan, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(valid, d_valid, (size_t)n_scan, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && nn_idx) e = cudaMemcpy(nn_idx, d_nn, sizeof(int32_t) * 5 * (size_t)n_scan, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && ms) cudaEventElapsedTime(ms, e0, e1);
  }
  cudaFree(d_map); cudaFree(d_scan); cudaFree(d_qt); cudaFree(d_out); cudaFree(d_valid); cudaFree(d_nn); cudaFree(d_grid); cudaFree(d_knn_d); cudaFree(d_knn_i);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  return e == cudaSuccess ? VILS_OK : vils::fail_cuda(e, "vils_lidar_associate");
}

}  // extern "C"
