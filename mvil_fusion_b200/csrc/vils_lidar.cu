// vils_lidar.cu — LiDAR per-point path of libvils_b200.so:
//   * stamp : PointProcessor::PointToRing (lidar_compensator/src/PointProcessor.cc:127-341) — ring id from elevation,
//             relative time from azimuth, intensity <- int(I) + rel_time;
//   * deskew: TransformToEnd (vils_estimator/src/lidar_frontend.cpp:1001-1041) — FP32 motion compensation to sweep end.
// One thread per point, 16-byte vector loads/stores on the PCL PointXYZI layout (x y z pad | intensity pad pad pad).
// Compiled with --fmad=false: the reference is plain IEEE FP32 arithmetic, and keeping every multiply/add separately
// rounded makes the GPU result bit-identical to it wherever the transcendental calls agree (they are evaluated in FP64
// and rounded once, i.e. correctly rounded FP32).
#include <cmath>
#include <cstdio>

#include <vector>

#include <algorithm>
#include <cstdlib>

#include "common.h"

#ifndef VILS_DESKEW_TMA_DEFAULT
#define VILS_DESKEW_TMA_DEFAULT 0
#endif
namespace {

struct DeskewParams { float qx, qy, qz, qw, tx, ty, tz, time_factor; double min_r, max_r; };

__device__ __forceinline__ float sin_cr(float x) { return (float)sin((double)x); }
__device__ __forceinline__ float acos_cr(float x) { return (float)acos((double)x); }

// Eigen QuaternionBase::_transformVector: uv = 2 (q.vec x v); v + w uv + q.vec x uv
__device__ __forceinline__ void rot(float qw, float qx, float qy, float qz, float& x, float& y, float& z) {
  float ux = qy * z - qz * y, uy = qz * x - qx * z, uz = qx * y - qy * x;
  ux += ux; uy += uy; uz += uz;
  const float cx = qy * uz - qz * uy, cy = qz * ux - qx * uz, cz = qx * uy - qy * ux;
  x = x + qw * ux + cx; y = y + qw * uy + cy; z = z + qw * uz + cz;
}

__device__ __forceinline__ void deskew_point(const DeskewParams& P, float& x, float& y, float& z, float& I) {
  const double distance = (double)sqrtf(x * x + y * y);                  // :1008
  const float s = P.time_factor * (I - (float)(int)I);                    // :1009
  if (s < 0 || (double)s > 1.001 || distance < P.min_r || distance > P.max_r) {   // :1011
    x = y = z = nanf(""); return;
  }
  x -= s * P.tx; y -= s * P.ty; z -= s * P.tz;                            // :1021-1023
  // q_s = identity.slerp(s, q_e)  [Eigen 3.3 slerp]
  const float d = P.qw, absD = fabsf(d);
  float s0, s1;
  if (absD >= 1.0f - 1.1920929e-07f) { s0 = 1.0f - s; s1 = s; }
  else { const float th = acos_cr(absD), st = sin_cr(th); s0 = sin_cr((1.0f - s) * th) / st; s1 = sin_cr(s * th) / st; }
  if (d < 0) s1 = -s1;
  float w = s0 + s1 * P.qw, qx = s1 * P.qx, qy = s1 * P.qy, qz = s1 * P.qz;   // s0*(1,0,0,0) + s1*q_e
  // conjugate().normalized()
  qx = -qx; qy = -qy; qz = -qz;
  const float n = sqrtf(w * w + qx * qx + qy * qy + qz * qz);
  w /= n; qx /= n; qy /= n; qz /= n;
  rot(w, qx, qy, qz, x, y, z);                                            // :1031
  rot(P.qw, P.qx, P.qy, P.qz, x, y, z);                                   // :1034
  x += P.tx; y += P.ty; z += P.tz;                                        // :1035-1037
  I = (float)(int)I;                                                       // :1038
}

__global__ void deskew_kernel8(float4* __restrict__ pts, int n, DeskewParams P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 a = pts[2 * i], b = pts[2 * i + 1];
  deskew_point(P, a.x, a.y, a.z, b.x);
  pts[2 * i] = a; pts[2 * i + 1] = b;
}
// ---- TMA variant (sm_90+ bulk async copies, mbarrier completion): persistent CTAs stream 8 KB tiles (256 PCL points) HBM -> shared memory with
// cp.async.bulk, de-skew them in place in shared memory and hand them back with a bulk shared -> global store.  DK_STAGES tiles are in flight per CTA,
// so the copy engine always has loads queued while the threads compute; no thread issues a global load or store itself.
constexpr int DK_TILE = 256, DK_STAGES = 4;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_1d(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__global__ void __launch_bounds__(DK_TILE) deskew_kernel8_tma(float4* __restrict__ pts, int n, DeskewParams P) {
  __shared__ __align__(128) float4 tile[DK_STAGES][2 * DK_TILE];
  __shared__ __align__(8) uint64_t full[DK_STAGES];
  const int tid = threadIdx.x, G = gridDim.x;
  const int ntiles = (n + DK_TILE - 1) / DK_TILE;
  if (tid == 0) { for (int s = 0; s < DK_STAGES; s++) mbar_init(&full[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  auto issue = [&](int it) {   // thread 0: load of this CTA's it-th tile into stage it % DK_STAGES
    const int t = blockIdx.x + it * G;
    if (t >= ntiles) return;
    const int cnt = min(DK_TILE, n - t * DK_TILE); const uint32_t bytes = (uint32_t)cnt * 32u;
    mbar_expect_tx(&full[it % DK_STAGES], bytes);
    tma_load_1d(tile[it % DK_STAGES], pts + (size_t)2 * t * DK_TILE, bytes, &full[it % DK_STAGES]);
  };
  if (tid == 0) for (int it = 0; it < DK_STAGES - 1; it++) issue(it);
  for (int it = 0;; it++) {
    const int t = blockIdx.x + it * G;
    if (t >= ntiles) break;
    const int s = it % DK_STAGES;
    if (tid == 0) {
      // stage (it - 1) % DK_STAGES was handed to the store engine in the previous iteration: it may be refilled once that store has READ it
      if (it >= 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      issue(it + DK_STAGES - 1);
    }
    mbar_wait(&full[s], (uint32_t)((it / DK_STAGES) & 1));
    const int cnt = min(DK_TILE, n - t * DK_TILE);
    if (tid < cnt) {
      float4 a = tile[s][2 * tid], b = tile[s][2 * tid + 1];
      deskew_point(P, a.x, a.y, a.z, b.x);
      tile[s][2 * tid] = a; tile[s][2 * tid + 1] = b;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the bulk store
    __syncthreads();
    if (tid == 0) tma_store_1d(pts + (size_t)2 * t * DK_TILE, tile[s], (uint32_t)cnt * 32u);
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__global__ void deskew_kernel_generic(float* __restrict__ pts, int n, int stride, DeskewParams P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float* p = pts + (size_t)i * stride;
  const int io = stride >= 8 ? 4 : 3;
  float x = p[0], y = p[1], z = p[2], I = p[io];
  deskew_point(P, x, y, z, I);
  p[0] = x; p[1] = y; p[2] = z; p[io] = I;
}

// ---- stamp ----
struct StampParams { float lower, factor, scan_period; int n_rings; };
__global__ void stamp_pass1(const float* __restrict__ pts, int n, int stride, StampParams P, int* __restrict__ ring, float* __restrict__ azi, int* first_valid) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = pts + (size_t)i * stride;
  const float x = p[0], y = p[1], z = p[2];
  int id = -1; float az = 0;
  if (isfinite(x) && isfinite(y) && isfinite(z)) {                        // PointProcessor.cc:161-166
    const float dis = sqrtf(x * x + y * y);                               // :168
    const float ele = (float)atan2((double)z, (double)dis);               // :169 (FP32 atan2, correctly rounded)
    az = (float)(2 * M_PI - (double)(float)atan2((double)y, (double)x)); // :170
    if ((double)az >= 2 * M_PI) az = (float)((double)az - 2 * M_PI);      // :174-177
    const float deg = (float)((double)ele * 180.0 / M_PI);                // RadToDeg
    id = (int)((deg - P.lower) * P.factor + 0.5);                          // PointProcessor.h:77-81 (double 0.5 promotes)
    if (id >= P.n_rings || id < 0) id = -1;                                // :181-184
  }
  ring[i] = id; azi[i] = az;
  if (id >= 0) atomicMin(first_valid, i);                                  // :186-190 start_ori = first kept point
}
__global__ void stamp_pass2(float* __restrict__ pts, int n, int stride, StampParams P, const int* __restrict__ ring, const float* __restrict__ azi, const int* first_valid) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || ring[i] < 0) return;
  const float start = azi[*first_valid];
  float rel = azi[i] - start;                                              // :318
  if (rel < 0) rel = (float)((double)rel + 2 * M_PI);                      // :319-322
  const float rel_time = (float)((double)(P.scan_period * rel) / (2 * M_PI));   // :324
  float* I = pts + (size_t)i * stride + (stride >= 8 ? 4 : 3);
  *I = (float)(int)(*I) + rel_time;                                        // :331
}

struct LidarDev { float* d = nullptr; int n = 0, stride = 0, device = 0; cudaStream_t st = nullptr; cudaEvent_t e0 = nullptr, e1 = nullptr; };

int run_deskew(float* d, int n, int stride, const DeskewParams& P, cudaStream_t st) {
  if (n == 0) return VILS_OK;
  const int T = 256, B = (n + T - 1) / T;
  // VILS_DESKEW_TMA: 1 = bulk-async (TMA) tile pipeline, 0 = one point per thread with 16-byte loads / stores.  Measured on B200 (round 2,
  // profiles/r2_deskew_tma.txt) — the default is the faster of the two at LiDAR-scan sizes.
  static const int use_tma = getenv("VILS_DESKEW_TMA") ? atoi(getenv("VILS_DESKEW_TMA")) : VILS_DESKEW_TMA_DEFAULT;
  if (stride == 8 && (reinterpret_cast<uintptr_t>(d) & 15) == 0 && use_tma && n >= 4 * DK_TILE) {
    static int n_sm = 0; if (!n_sm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
    static const int gm = getenv("VILS_DESKEW_TMA_GRID") ? atoi(getenv("VILS_DESKEW_TMA_GRID")) : 4;
    const int tiles = (n + DK_TILE - 1) / DK_TILE, grid = std::min(tiles, n_sm * gm);
    deskew_kernel8_tma<<<grid, DK_TILE, 0, st>>>(reinterpret_cast<float4*>(d), n, P);
  } else if (stride == 8 && (reinterpret_cast<uintptr_t>(d) & 15) == 0) deskew_kernel8<<<B, T, 0, st>>>(reinterpret_cast<float4*>(d), n, P);
  else deskew_kernel_generic<<<B, T, 0, st>>>(d, n, stride, P);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? VILS_OK : vils::fail_cuda(e, "deskew launch");
}

DeskewParams mk(const float q[4], const float t[3], float tf, float mn, float mx) {
  DeskewParams P; P.qx = q[0]; P.qy = q[1]; P.qz = q[2]; P.qw = q[3]; P.tx = t[0]; P.ty = t[1]; P.tz = t[2]; P.time_factor = tf;
  P.min_r = (double)mn; P.max_r = (double)mx; return P;
}

}  // namespace

extern "C" {

int vils_deskew(float* xyzi, int32_t n, int32_t stride, const float q[4], const float t[3], float time_factor, float min_r, float max_r, int32_t device) {
  if (!xyzi || n < 0 || stride < 4 || !q || !t) return vils::fail(VILS_ERR_BAD_ARG, "vils_deskew: bad argument");
  int st = vils::require_device(device); if (st) return st;
  if (n == 0) return VILS_OK;
  float* d = nullptr; const size_t bytes = (size_t)n * stride * sizeof(float);
  cudaError_t e = cudaMalloc(&d, bytes);
  if (e != cudaSuccess) return vils::fail_cuda(e, "vils_deskew alloc");
  e = cudaMemcpy(d, xyzi, bytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) { st = run_deskew(d, n, stride, mk(q, t, time_factor, min_r, max_r), 0); if (st) { cudaFree(d); return st; } e = cudaMemcpy(xyzi, d, bytes, cudaMemcpyDeviceToHost); }
  cudaFree(d);
  return e == cudaSuccess ? VILS_OK : vils::fail_cuda(e, "vils_deskew copy");
}

}  // extern "C" (reopened below)

// PointProcessor::PointToRing() (lidar_compensator/src/PointProcessor.cc:106-125): cloud_in_rings_ = ring 0 points, then ring 1, ... each in
// scan order.  One CTA per ring walks the cloud in order; a block-wide exclusive scan of "ring == r" gives every kept point its stable
// position inside the ring, the ring's base comes from a histogram.
__global__ void ring_hist_kernel(const int* __restrict__ ring, int n, int n_rings, int* __restrict__ count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && ring[i] >= 0 && ring[i] < n_rings) atomicAdd(&count[ring[i]], 1);
}
__global__ void __launch_bounds__(1024) ring_scatter_kernel(const float* __restrict__ pts, const int* __restrict__ ring, int n, int stride, const int* __restrict__ start,
                                                           float* __restrict__ out) {
  __shared__ int wsum[32];
  __shared__ int base_s;
  const int r = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) base_s = start[r];
  __syncthreads();
  for (int c0 = 0; c0 < n; c0 += 1024) {
    const int i = c0 + t;
    const bool mine = i < n && ring[i] == r;
    const unsigned int bal = __ballot_sync(0xffffffffu, mine);
    const int within = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < 32; w++) { const int v = wsum[w]; if (w < warp) before += v; total += v; }
    if (mine) {
      const size_t dst = (size_t)(base_s + before + within) * stride, src = (size_t)i * stride;
      for (int k = 0; k < stride; k++) out[dst + k] = pts[src + k];
    }
    __syncthreads();
    if (t == 0) base_s += total;
    __syncthreads();
  }
}

extern "C" {

// PointProcessor::PointToRing (PointProcessor.cc:106-341) complete: stamp (ring id + relative time in the intensity) and the ring-major
// re-ordering.  out: the kept points, ring 0 first (n_kept x stride floats, capacity n); ring_start[n_rings + 1]: first point of every
// ring in `out`.  The input cloud is not modified.
int vils_point_to_ring(const float* xyzi, int32_t n, int32_t stride, float lower_deg, float upper_deg, int32_t n_rings, float scan_period, float* out,
                       int32_t* ring_start, int32_t device) {
  if (!xyzi || !out || !ring_start || n < 0 || stride < 4 || n_rings < 2 || n_rings > 1024) return vils::fail(VILS_ERR_BAD_ARG, "vils_point_to_ring: bad argument");
  int st = vils::require_device(device); if (st) return st;
  for (int r = 0; r <= n_rings; r++) ring_start[r] = 0;
  if (n == 0) return VILS_OK;
  float* d = nullptr; float* dout = nullptr; int* ring = nullptr; float* azi = nullptr; int* first = nullptr; int* cnt = nullptr;
  const size_t bytes = (size_t)n * stride * sizeof(float);
  cudaError_t e = cudaMalloc(&d, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&dout, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&ring, sizeof(int) * n);
  if (e == cudaSuccess) e = cudaMalloc(&azi, sizeof(float) * n);
  if (e == cudaSuccess) e = cudaMalloc(&first, sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&cnt, sizeof(int) * (n_rings + 1));
  const int big = n;
  if (e == cudaSuccess) e = cudaMemcpy(first, &big, sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d, xyzi, bytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(cnt, 0, sizeof(int) * (n_rings + 1));
  std::vector<int> hc(n_rings + 1, 0);
  if (e == cudaSuccess) {
    StampParams P; P.lower = lower_deg; P.factor = (n_rings - 1) / (upper_deg - lower_deg); P.scan_period = scan_period; P.n_rings = n_rings;
    const int T = 256, B = (n + T - 1) / T;
    stamp_pass1<<<B, T>>>(d, n, stride, P, ring, azi, first);
    stamp_pass2<<<B, T>>>(d, n, stride, P, ring, azi, first);
    ring_hist_kernel<<<B, T>>>(ring, n, n_rings, cnt);
    e = cudaMemcpy(hc.data(), cnt, sizeof(int) * n_rings, cudaMemcpyDeviceToHost);
  }
  if (e == cudaSuccess) {
    int acc = 0;
    for (int r = 0; r < n_rings; r++) { ring_start[r] = acc; acc += hc[r]; }
    ring_start[n_rings] = acc;
    e = cudaMemcpy(cnt, ring_start, sizeof(int) * (n_rings + 1), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) { ring_scatter_kernel<<<n_rings, 1024>>>(d, ring, n, stride, cnt, dout); e = cudaGetLastError(); }
    if (e == cudaSuccess && acc > 0) e = cudaMemcpy(out, dout, (size_t)acc * stride * sizeof(float), cudaMemcpyDeviceToHost);
  }
  cudaFree(d); cudaFree(dout); cudaFree(ring); cudaFree(azi); cudaFree(first); cudaFree(cnt);
  return e == cudaSuccess ? VILS_OK : vils::fail_cuda(e, "vils_point_to_ring");
}

int vils_stamp_rings(float* xyzi, int32_t n, int32_t stride, float lower_deg, float upper_deg, int32_t n_rings, float scan_period, int32_t* ring_out, int32_t device) {
  if (!xyzi || !ring_out || n < 0 || stride < 4 || n_rings < 2) return vils::fail(VILS_ERR_BAD_ARG, "vils_stamp_rings: bad argument");
  int st = vils::require_device(device); if (st) return st;
  if (n == 0) return VILS_OK;
  float* d = nullptr; int* ring = nullptr; float* azi = nullptr; int* first = nullptr;
  const size_t bytes = (size_t)n * stride * sizeof(float);
  cudaError_t e = cudaMalloc(&d, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&ring, sizeof(int) * n);
  if (e == cudaSuccess) e = cudaMalloc(&azi, sizeof(float) * n);
  if (e == cudaSuccess) e = cudaMalloc(&first, sizeof(int));
  const int big = n;
  if (e == cudaSuccess) e = cudaMemcpy(first, &big, sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d, xyzi, bytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    StampParams P; P.lower = lower_deg; P.factor = (n_rings - 1) / (upper_deg - lower_deg); P.scan_period = scan_period; P.n_rings = n_rings;
    const int T = 256, B = (n + T - 1) / T;
    stamp_pass1<<<B, T>>>(d, n, stride, P, ring, azi, first);
    stamp_pass2<<<B, T>>>(d, n, stride, P, ring, azi, first);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(xyzi, d, bytes, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(ring_out, ring, sizeof(int) * n, cudaMemcpyDeviceToHost);
  cudaFree(d); cudaFree(ring); cudaFree(azi); cudaFree(first);
  return e == cudaSuccess ? VILS_OK : vils::fail_cuda(e, "vils_stamp_rings");
}

int vils_lidar_dev_alloc(int32_t n, int32_t stride, int32_t device, void** handle) {
  if (!handle || n <= 0 || stride < 4) return vils::fail(VILS_ERR_BAD_ARG, "vils_lidar_dev_alloc: bad argument");
  int st = vils::require_device(device); if (st) return st;
  LidarDev* h = new LidarDev(); h->n = n; h->stride = stride; h->device = device;
  cudaError_t e = cudaMalloc(&h->d, (size_t)n * stride * sizeof(float));
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&h->e0);
  if (e == cudaSuccess) e = cudaEventCreate(&h->e1);
  if (e != cudaSuccess) { cudaFree(h->d); delete h; return vils::fail_cuda(e, "vils_lidar_dev_alloc"); }
  *handle = h; return VILS_OK;
}
int vils_lidar_dev_upload(void* handle, const float* xyzi) {
  LidarDev* h = static_cast<LidarDev*>(handle); if (!h || !xyzi) return vils::fail(VILS_ERR_BAD_ARG, "null");
  cudaSetDevice(h->device);
  cudaError_t e = cudaMemcpyAsync(h->d, xyzi, (size_t)h->n * h->stride * sizeof(float), cudaMemcpyHostToDevice, h->st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
  return e == cudaSuccess ? VILS_OK : vils::fail_cuda(e, "lidar upload");
}
int vils_lidar_dev_deskew(void* handle, const float q[4], const float t[3], float time_factor, float min_r, float max_r, float* ms) {
  LidarDev* h = static_cast<LidarDev*>(handle); if (!h || !q || !t) return vils::fail(VILS_ERR_BAD_ARG, "null");
  cudaSetDevice(h->device);
  cudaEventRecord(h->e0, h->st);
  int st = run_deskew(h->d, h->n, h->stride, mk(q, t, time_factor, min_r, max_r), h->st); if (st) return st;
  cudaEventRecord(h->e1, h->st);
  cudaError_t e = cudaStreamSynchronize(h->st);
  if (e != cudaSuccess) return vils::fail_cuda(e, "deskew");
  if (ms) cudaEventElapsedTime(ms, h->e0, h->e1);
  return VILS_OK;
}
int vils_lidar_dev_download(void* handle, float* xyzi) {
  LidarDev* h = static_cast<LidarDev*>(handle); if (!h || !xyzi) return vils::fail(VILS_ERR_BAD_ARG, "null");
  cudaSetDevice(h->device);
  cudaError_t e = cudaMemcpyAsync(xyzi, h->d, (size_t)h->n * h->stride * sizeof(float), cudaMemcpyDeviceToHost, h->st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
  return e == cudaSuccess ? VILS_OK : vils::fail_cuda(e, "lidar download");
}
void vils_lidar_dev_free(void* handle) {
  LidarDev* h = static_cast<LidarDev*>(handle); if (!h) return;
  cudaSetDevice(h->device); cudaFree(h->d); cudaEventDestroy(h->e0); cudaEventDestroy(h->e1); cudaStreamDestroy(h->st); delete h;
}

}  // extern "C"
