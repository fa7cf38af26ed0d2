// vils_klt.cu — pyramidal Lucas-Kanade tracker behind FeatureTracker::readImage()
// (feature_tracker_/src/feature_tracker.cpp:113: cv::calcOpticalFlowPyrLK(cur, forw, cur_pts, forw_pts, status, err, Size(21,21), 3)).
//
// The algorithm is OpenCV's (modules/video/src/lkpyramid.cpp, a third-party dependency that is NOT in the reference tree;
// restated from its published behaviour, checked against cv2 4.13.0 in tests/test_klt_gpu.py):
//   * pyramids: level 0 = image, level k = pyrDown 5x5 [1 4 6 4 1]^2, (sum + 128) >> 8, BORDER_REFLECT_101, size (w+1)/2;
//   * derivatives of the PREVIOUS pyramid: un-normalised 3x3 Scharr in int16, reflect-101 inside, constant 0 outside;
//   * per point and level (coarse to fine): bilinear patch with 14-bit integer weights (patch in Q5, derivatives as is),
//     A = sum [IxIx IxIy; IxIy IyIy] * 2^-20, reject when minEig(A)/(w*h) < 1e-4 or det < FLT_EPSILON (status only at level 0),
//     <= 30 iterations of d = A^-1 b, stop at |d|^2 <= 1e-4 or on oscillation (half step back), next level starts at 2x;
//   * err = mean |J - I| / 32 at level 0.
// Integer pipeline (pyramids, Scharr, interpolated patches, differences) is bit-exact; the sums A and b are accumulated
// exactly (integers held in FP64) and rounded once to FP32, where OpenCV rounds after every add — the only source of
// sub-1e-3 px differences.  Compiled with --fmad=false so the FP32 step arithmetic is IEEE like the CPU's.
//
// Mapping: pyrDown / Scharr are tiled per-pixel kernels (coalesced u8 / short2); tracking runs ONE CTA PER FEATURE for the
// whole level chain (448 threads = one per window pixel, the previous-image patch and its derivatives stay in registers,
// b is reduced with warp shuffles), so a 150-corner frame fills the 148 SMs once.
#include <cstdio>
#include <cstring>
#include <vector>

#include <algorithm>
#include <utility>

#include "common.h"

namespace {

constexpr int MAX_LEVELS = 8;
constexpr int TRACK_THREADS = 448;

struct Pyr { uint8_t* img[MAX_LEVELS]; short2* der[MAX_LEVELS]; int w[MAX_LEVELS], h[MAX_LEVELS]; };

__device__ __forceinline__ int refl101(int i, int n) { if (i < 0) i = -i; if (i >= n) i = 2 * n - 2 - i; return i; }

__global__ void pyrdown_kernel(const uint8_t* __restrict__ src, int sw, int sh, uint8_t* __restrict__ dst, int dw, int dh) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dw || y >= dh) return;
  const int k[5] = {1, 4, 6, 4, 1};
  int sum = 0;
#pragma unroll
  for (int dy = 0; dy < 5; dy++) {
    const int sy = refl101(2 * y + dy - 2, sh);
    int row = 0;
#pragma unroll
    for (int dx = 0; dx < 5; dx++) row += k[dx] * src[(size_t)sy * sw + refl101(2 * x + dx - 2, sw)];
    sum += k[dy] * row;
  }
  dst[(size_t)y * dw + x] = (uint8_t)((sum + 128) >> 8);
}

// calcScharrDeriv: dx = [-3 0 3; -10 0 10; -3 0 3], dy = transpose; reflect-101 at the image edge.
__global__ void scharr_kernel(const uint8_t* __restrict__ src, int w, int h, short2* __restrict__ dst) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const int ym = refl101(y - 1, h), yp = refl101(y + 1, h), xm = refl101(x - 1, w), xp = refl101(x + 1, w);
  auto P = [&](int yy, int xx) { return (int)src[(size_t)yy * w + xx]; };
  const int t0m = (P(ym, xm) + P(yp, xm)) * 3 + P(y, xm) * 10, t0p = (P(ym, xp) + P(yp, xp)) * 3 + P(y, xp) * 10;
  const int t1m = P(yp, xm) - P(ym, xm), t1c = P(yp, x) - P(ym, x), t1p = P(yp, xp) - P(ym, xp);
  dst[(size_t)y * w + x] = make_short2((short)(t0p - t0m), (short)((t1p + t1m) * 3 + t1c * 10));
}

__device__ __forceinline__ int sample_u8(const uint8_t* img, int w, int h, int x, int y) { return img[(size_t)refl101(y, h) * w + refl101(x, w)]; }
__device__ __forceinline__ short2 sample_der(const short2* d, int w, int h, int x, int y) {
  if (x < 0 || y < 0 || x >= w || y >= h) return make_short2(0, 0);   // copyMakeBorder(BORDER_CONSTANT) around derivI
  return d[(size_t)y * w + x];
}

__device__ __forceinline__ double block_sum3(double v, double* red, int slot) {   // exact: integers in FP64
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[slot * 16 + (threadIdx.x >> 5)] = v;
  return v;
}

struct TrackParams { Pyr prev, next; const float2* prev_pts; float2* next_pts; uint8_t* status; float* err; int n, levels, win; float min_eig; int max_iter; double eps2; };

__global__ void __launch_bounds__(TRACK_THREADS) track_kernel(TrackParams T) {
  const int p = blockIdx.x;
  if (p >= T.n) return;
  const int tid = threadIdx.x, win = T.win, npx = win * win;
  const bool active = tid < npx;
  const int wx = active ? tid % win : 0, wy = active ? tid / win : 0;
  __shared__ double red[3 * 16];
  __shared__ float sh_f[8];
  __shared__ int sh_i[4];
  const float halfWin = (win - 1) * 0.5f;
  const int nwarp = TRACK_THREADS / 32;
  float2 nextPtG = make_float2(0, 0);   // nextPts[ptidx] (full window coordinates), carried across levels
  bool st = true; float errv = 0.f;
  const float2 p0 = T.prev_pts[p];
  for (int level = T.levels - 1; level >= 0; level--) {
    const int w = T.prev.w[level], h = T.prev.h[level];
    const uint8_t* I = T.prev.img[level]; const short2* dI = T.prev.der[level]; const uint8_t* J = T.next.img[level];
    const float scale = 1.f / (float)(1 << level);
    float2 prevPt = make_float2(p0.x * scale, p0.y * scale);
    float2 nextPt = (level == T.levels - 1) ? prevPt : make_float2(nextPtG.x * 2.f, nextPtG.y * 2.f);
    nextPtG = nextPt;
    prevPt.x -= halfWin; prevPt.y -= halfWin;
    const int ipx = (int)floorf(prevPt.x), ipy = (int)floorf(prevPt.y);
    if (ipx < -win || ipx >= w || ipy < -win || ipy >= h) { if (level == 0) { st = false; errv = 0.f; } continue; }
    float a = prevPt.x - ipx, b = prevPt.y - ipy;
    int iw00 = __float2int_rn((1.f - a) * (1.f - b) * 16384.f), iw01 = __float2int_rn(a * (1.f - b) * 16384.f);
    int iw10 = __float2int_rn((1.f - a) * b * 16384.f), iw11 = 16384 - iw00 - iw01 - iw10;
    int Ival = 0, Ix = 0, Iy = 0;
    if (active) {
      const int x = ipx + wx, y = ipy + wy;
      Ival = (sample_u8(I, w, h, x, y) * iw00 + sample_u8(I, w, h, x + 1, y) * iw01 + sample_u8(I, w, h, x, y + 1) * iw10 + sample_u8(I, w, h, x + 1, y + 1) * iw11 + (1 << 8)) >> 9;
      const short2 d00 = sample_der(dI, w, h, x, y), d01 = sample_der(dI, w, h, x + 1, y), d10 = sample_der(dI, w, h, x, y + 1), d11 = sample_der(dI, w, h, x + 1, y + 1);
      Ix = (d00.x * iw00 + d01.x * iw01 + d10.x * iw10 + d11.x * iw11 + (1 << 13)) >> 14;
      Iy = (d00.y * iw00 + d01.y * iw01 + d10.y * iw10 + d11.y * iw11 + (1 << 13)) >> 14;
      Ival = (short)Ival; Ix = (short)Ix; Iy = (short)Iy;
    }
    __syncthreads();
    block_sum3((double)Ix * Ix, red, 0); block_sum3((double)Ix * Iy, red, 1); block_sum3((double)Iy * Iy, red, 2);
    __syncthreads();
    double s11 = 0, s12 = 0, s22 = 0;
    for (int k = 0; k < nwarp; k++) { s11 += red[k]; s12 += red[16 + k]; s22 += red[32 + k]; }
    const float FLT_SCALE = 1.f / (float)(1 << 20);
    const float A11 = (float)s11 * FLT_SCALE, A12 = (float)s12 * FLT_SCALE, A22 = (float)s22 * FLT_SCALE;
    float D = A11 * A22 - A12 * A12;
    const float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * npx);
    if (minEig < T.min_eig || D < 1.1920929e-07f) { if (level == 0) st = false; continue; }
    D = 1.f / D;
    nextPt.x -= halfWin; nextPt.y -= halfWin;
    float2 prevDelta = make_float2(0, 0);
    for (int j = 0; j < T.max_iter; j++) {
      const int inx = (int)floorf(nextPt.x), iny = (int)floorf(nextPt.y);
      if (inx < -win || inx >= w || iny < -win || iny >= h) { if (level == 0) st = false; break; }
      a = nextPt.x - inx; b = nextPt.y - iny;
      iw00 = __float2int_rn((1.f - a) * (1.f - b) * 16384.f); iw01 = __float2int_rn(a * (1.f - b) * 16384.f);
      iw10 = __float2int_rn((1.f - a) * b * 16384.f); iw11 = 16384 - iw00 - iw01 - iw10;
      double c1 = 0, c2 = 0;
      if (active) {
        const int x = inx + wx, y = iny + wy;
        const int diff = ((sample_u8(J, w, h, x, y) * iw00 + sample_u8(J, w, h, x + 1, y) * iw01 + sample_u8(J, w, h, x, y + 1) * iw10 + sample_u8(J, w, h, x + 1, y + 1) * iw11 + (1 << 8)) >> 9) - Ival;
        c1 = (double)(diff * Ix); c2 = (double)(diff * Iy);
      }
      __syncthreads();
      block_sum3(c1, red, 0); block_sum3(c2, red, 1);
      __syncthreads();
      double sb1 = 0, sb2 = 0;
      for (int k = 0; k < nwarp; k++) { sb1 += red[k]; sb2 += red[16 + k]; }
      const float b1 = (float)sb1 * FLT_SCALE, b2 = (float)sb2 * FLT_SCALE;
      const float2 delta = make_float2((A12 * b2 - A22 * b1) * D, (A12 * b1 - A11 * b2) * D);
      nextPt.x += delta.x; nextPt.y += delta.y;
      nextPtG = make_float2(nextPt.x + halfWin, nextPt.y + halfWin);
      if ((double)delta.x * delta.x + (double)delta.y * delta.y <= T.eps2) break;    // Point2f::ddot is FP64
      if (j > 0 && fabsf(delta.x + prevDelta.x) < 0.01f && fabsf(delta.y + prevDelta.y) < 0.01f) {
        nextPtG.x -= delta.x * 0.5f; nextPtG.y -= delta.y * 0.5f;
        break;
      }
      prevDelta = delta;
    }
    if (st && level == 0) {   // err = mean |J - I| / 32 over the window at the final position
      const float nx = nextPtG.x - halfWin, ny = nextPtG.y - halfWin;
      const int inx = (int)floorf(nx), iny = (int)floorf(ny);
      if (inx < -win || inx >= w || iny < -win || iny >= h) { st = false; }
      else {
        const float aa = nx - inx, bb = ny - iny;
        iw00 = __float2int_rn((1.f - aa) * (1.f - bb) * 16384.f); iw01 = __float2int_rn(aa * (1.f - bb) * 16384.f);
        iw10 = __float2int_rn((1.f - aa) * bb * 16384.f); iw11 = 16384 - iw00 - iw01 - iw10;
        double e = 0;
        if (active) {
          const int x = inx + wx, y = iny + wy;
          const int diff = ((sample_u8(J, w, h, x, y) * iw00 + sample_u8(J, w, h, x + 1, y) * iw01 + sample_u8(J, w, h, x, y + 1) * iw10 + sample_u8(J, w, h, x + 1, y + 1) * iw11 + (1 << 8)) >> 9) - Ival;
          e = (double)abs(diff);
        }
        __syncthreads();
        block_sum3(e, red, 0);
        __syncthreads();
        double se = 0; for (int k = 0; k < nwarp; k++) se += red[k];
        errv = (float)se * (1.f / (float)(32 * npx));
      }
    }
  }
  (void)sh_f; (void)sh_i;
  if (tid == 0) { T.next_pts[p] = nextPtG; T.status[p] = st ? 1 : 0; T.err[p] = st ? errv : (errv); }
}

}  // namespace

struct vils_klt {
  int rows = 0, cols = 0, max_pts = 0, win = 21, levels = 4, device = 0;
  Pyr prev{}, next{};
  float2* d_prev_pts = nullptr; float2* d_next_pts = nullptr; uint8_t* d_status = nullptr; float* d_err = nullptr;
  uint8_t* h_stage = nullptr;   // pinned staging for the two images
  cudaStream_t st = nullptr; cudaEvent_t e0 = nullptr, e1 = nullptr;
  float last_ms = 0; int n = 0;
  cudaGraphExec_t graph_exec = nullptr; bool graph_failed = false;
  // device-resident chain (vils_klt_advance): the `next` pyramid of one call is the `prev` pyramid of the following one
  bool has_next = false; int parity = 0; cudaGraphExec_t graph_adv[2] = {nullptr, nullptr}; bool graph_adv_failed = false;
};

static void klt_launch_pyramids(vils_klt* k) {
  dim3 B(32, 8);
  for (int img = 0; img < 2; img++) {
    Pyr& P = img ? k->next : k->prev;
    for (int l = 1; l < k->levels; l++) {
      dim3 G((P.w[l] + B.x - 1) / B.x, (P.h[l] + B.y - 1) / B.y);
      pyrdown_kernel<<<G, B, 0, k->st>>>(P.img[l - 1], P.w[l - 1], P.h[l - 1], P.img[l], P.w[l], P.h[l]);
    }
  }
  for (int l = 0; l < k->levels; l++) {
    dim3 G((k->prev.w[l] + B.x - 1) / B.x, (k->prev.h[l] + B.y - 1) / B.y);
    scharr_kernel<<<G, B, 0, k->st>>>(k->prev.img[l], k->prev.w[l], k->prev.h[l], k->prev.der[l]);
  }
}

// vils_klt_advance: only the NEW image's pyramid is built; the Scharr derivatives are taken on the pyramid inherited from the previous call
static void klt_launch_advance(vils_klt* k) {
  dim3 B(32, 8);
  for (int l = 1; l < k->levels; l++) {
    dim3 G((k->next.w[l] + B.x - 1) / B.x, (k->next.h[l] + B.y - 1) / B.y);
    pyrdown_kernel<<<G, B, 0, k->st>>>(k->next.img[l - 1], k->next.w[l - 1], k->next.h[l - 1], k->next.img[l], k->next.w[l], k->next.h[l]);
  }
  for (int l = 0; l < k->levels; l++) {
    dim3 G((k->prev.w[l] + B.x - 1) / B.x, (k->prev.h[l] + B.y - 1) / B.y);
    scharr_kernel<<<G, B, 0, k->st>>>(k->prev.img[l], k->prev.w[l], k->prev.h[l], k->prev.der[l]);
  }
}

static int klt_build_and_track(vils_klt* k) {
  // the 2 x (levels - 1) pyrDown and the `levels` Scharr launches have fixed arguments for the life of the handle: they are captured
  // once into a CUDA graph (10 launches -> 1 for 640x480, a launch-bound stretch at this image size); the tracking kernel, whose grid
  // is the number of points of the call, follows as a normal launch
  if (!k->graph_exec && !k->graph_failed) {
    cudaGraph_t g = nullptr;
    if (cudaStreamBeginCapture(k->st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      klt_launch_pyramids(k);
      if (cudaStreamEndCapture(k->st, &g) != cudaSuccess || !g || cudaGraphInstantiate(&k->graph_exec, g, 0) != cudaSuccess) { k->graph_exec = nullptr; k->graph_failed = true; }
      if (g) cudaGraphDestroy(g);
    } else k->graph_failed = true;
    cudaGetLastError();
  }
  if (k->graph_exec) cudaGraphLaunch(k->graph_exec, k->st);
  else klt_launch_pyramids(k);
  if (k->n > 0) {
    TrackParams T; T.prev = k->prev; T.next = k->next; T.prev_pts = k->d_prev_pts; T.next_pts = k->d_next_pts; T.status = k->d_status; T.err = k->d_err;
    T.n = k->n; T.levels = k->levels; T.win = k->win; T.min_eig = 1e-4f; T.max_iter = 30; T.eps2 = 0.01 * 0.01;   // criteria.epsilon *= criteria.epsilon (double)
    track_kernel<<<k->n, TRACK_THREADS, 0, k->st>>>(T);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? VILS_OK : vils::fail_cuda(e, "klt launch");
}

extern "C" {

int vils_klt_create(int32_t rows, int32_t cols, int32_t max_pts, int32_t win, int32_t max_level, int32_t device, vils_klt** out) {
  if (!out || rows <= 0 || cols <= 0 || max_pts <= 0 || win < 3 || win > 21 || (win & 1) == 0 || max_level < 0 || max_level >= MAX_LEVELS)
    return vils::fail(VILS_ERR_BAD_ARG, "vils_klt_create: bad argument (odd window <= 21 supported)");
  int st = vils::require_device(device); if (st) return st;
  vils_klt* k = new vils_klt(); k->rows = rows; k->cols = cols; k->max_pts = max_pts; k->win = win; k->device = device;
  // buildOpticalFlowPyramid stops when a level is not larger than the window
  int w = cols, h = rows, L = 1;
  k->prev.w[0] = k->next.w[0] = w; k->prev.h[0] = k->next.h[0] = h;
  for (int l = 1; l <= max_level; l++) {
    w = (w + 1) / 2; h = (h + 1) / 2;
    if (w <= win || h <= win) break;
    k->prev.w[l] = k->next.w[l] = w; k->prev.h[l] = k->next.h[l] = h; L = l + 1;
  }
  k->levels = L;
  cudaError_t e = cudaStreamCreateWithFlags(&k->st, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&k->e0);
  if (e == cudaSuccess) e = cudaEventCreate(&k->e1);
  for (int l = 0; l < L && e == cudaSuccess; l++) {
    const size_t px = (size_t)k->prev.w[l] * k->prev.h[l];
    e = cudaMalloc(&k->prev.img[l], px); if (e == cudaSuccess) e = cudaMalloc(&k->next.img[l], px);
    if (e == cudaSuccess) e = cudaMalloc(&k->prev.der[l], px * sizeof(short2));
  }
  if (e == cudaSuccess) e = cudaMalloc(&k->d_prev_pts, sizeof(float2) * max_pts);
  if (e == cudaSuccess) e = cudaMalloc(&k->d_next_pts, sizeof(float2) * max_pts);
  if (e == cudaSuccess) e = cudaMalloc(&k->d_status, max_pts);
  if (e == cudaSuccess) e = cudaMalloc(&k->d_err, sizeof(float) * max_pts);
  if (e == cudaSuccess) e = cudaMallocHost(&k->h_stage, (((size_t)2 * rows * cols + 15) & ~(size_t)15) + (size_t)max_pts * 16);
  if (e != cudaSuccess) { vils_klt_destroy(k); return vils::fail_cuda(e, "vils_klt_create"); }
  *out = k; return VILS_OK;
}

void vils_klt_destroy(vils_klt* k) {
  if (!k) return;
  cudaSetDevice(k->device);
  if (k->st) cudaStreamSynchronize(k->st);
  if (k->graph_exec) cudaGraphExecDestroy(k->graph_exec);
  for (int p = 0; p < 2; p++) if (k->graph_adv[p]) cudaGraphExecDestroy(k->graph_adv[p]);
  for (int l = 0; l < MAX_LEVELS; l++) { cudaFree(k->prev.img[l]); cudaFree(k->next.img[l]); cudaFree(k->prev.der[l]); }
  cudaFree(k->d_prev_pts); cudaFree(k->d_next_pts); cudaFree(k->d_status); cudaFree(k->d_err); cudaFreeHost(k->h_stage);
  if (k->e0) cudaEventDestroy(k->e0);
  if (k->e1) cudaEventDestroy(k->e1);
  if (k->st) cudaStreamDestroy(k->st);
  delete k;
}

int vils_klt_upload(vils_klt* k, const uint8_t* prev, const uint8_t* next, int32_t stride, const float* prev_xy, int32_t n) {
  if (!k || !prev || !next || stride < k->cols || n < 0 || n > k->max_pts || (n && !prev_xy)) return vils::fail(VILS_ERR_BAD_ARG, "vils_klt_upload: bad argument");
  cudaSetDevice(k->device);
  const size_t px = (size_t)k->rows * k->cols;
  for (int y = 0; y < k->rows; y++) {   // compact into pinned staging (drops the row padding)
    memcpy(k->h_stage + (size_t)y * k->cols, prev + (size_t)y * stride, k->cols);
    memcpy(k->h_stage + px + (size_t)y * k->cols, next + (size_t)y * stride, k->cols);
  }
  float* hp = reinterpret_cast<float*>(k->h_stage + ((2 * px + 15) & ~(size_t)15));
  if (n) memcpy(hp, prev_xy, sizeof(float) * 2 * n);
  cudaMemcpyAsync(k->prev.img[0], k->h_stage, px, cudaMemcpyHostToDevice, k->st);
  cudaMemcpyAsync(k->next.img[0], k->h_stage + px, px, cudaMemcpyHostToDevice, k->st);
  if (n) cudaMemcpyAsync(k->d_prev_pts, hp, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, k->st);
  k->n = n;
  cudaError_t e = cudaStreamSynchronize(k->st);
  return e == cudaSuccess ? VILS_OK : vils::fail_cuda(e, "vils_klt_upload");
}

int vils_klt_track_device(vils_klt* k) {
  if (!k) return vils::fail(VILS_ERR_BAD_ARG, "null");
  cudaSetDevice(k->device);
  cudaEventRecord(k->e0, k->st);
  int st = klt_build_and_track(k); if (st) return st;
  cudaEventRecord(k->e1, k->st);
  cudaError_t e = cudaStreamSynchronize(k->st);
  if (e != cudaSuccess) return vils::fail_cuda(e, "vils_klt_track_device");
  cudaEventElapsedTime(&k->last_ms, k->e0, k->e1);
  return VILS_OK;
}

int vils_klt_download(vils_klt* k, float* next_xy, uint8_t* status, float* err) {
  if (!k) return vils::fail(VILS_ERR_BAD_ARG, "null");
  cudaSetDevice(k->device);
  if (k->n == 0) return VILS_OK;
  if (next_xy) cudaMemcpyAsync(next_xy, k->d_next_pts, sizeof(float) * 2 * k->n, cudaMemcpyDeviceToHost, k->st);
  if (status) cudaMemcpyAsync(status, k->d_status, k->n, cudaMemcpyDeviceToHost, k->st);
  if (err) cudaMemcpyAsync(err, k->d_err, sizeof(float) * k->n, cudaMemcpyDeviceToHost, k->st);
  cudaError_t e = cudaStreamSynchronize(k->st);
  return e == cudaSuccess ? VILS_OK : vils::fail_cuda(e, "vils_klt_download");
}

// The call readImage makes: host buffers in, host buffers out.  Everything is queued on the handle's stream and closed by ONE
// synchronisation (copies in, pyramids, derivatives, tracking, copies out through pinned staging).
int vils_klt_track(vils_klt* k, const uint8_t* prev, const uint8_t* next, int32_t stride, const float* prev_xy, int32_t n, float* next_xy,
                   uint8_t* status, float* err) {
  if (!k || !prev || !next || stride < k->cols || n < 0 || n > k->max_pts || (n && !prev_xy)) return vils::fail(VILS_ERR_BAD_ARG, "vils_klt_track: bad argument");
  cudaSetDevice(k->device);
  const size_t px = (size_t)k->rows * k->cols;
  for (int y = 0; y < k->rows; y++) {   // compact into pinned staging (drops the row padding)
    memcpy(k->h_stage + (size_t)y * k->cols, prev + (size_t)y * stride, k->cols);
    memcpy(k->h_stage + px + (size_t)y * k->cols, next + (size_t)y * stride, k->cols);
  }
  float* hp = reinterpret_cast<float*>(k->h_stage + ((2 * px + 15) & ~(size_t)15));   // max_pts x 16 bytes: [xy in / xy out (8) | err (4) | status (1)]
  if (n) memcpy(hp, prev_xy, sizeof(float) * 2 * n);
  cudaMemcpyAsync(k->prev.img[0], k->h_stage, px, cudaMemcpyHostToDevice, k->st);
  cudaMemcpyAsync(k->next.img[0], k->h_stage + px, px, cudaMemcpyHostToDevice, k->st);
  if (n) cudaMemcpyAsync(k->d_prev_pts, hp, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, k->st);
  k->n = n;
  cudaEventRecord(k->e0, k->st);
  int st = klt_build_and_track(k); if (st) return st;
  cudaEventRecord(k->e1, k->st);
  float* h_err = hp + 2 * (size_t)k->max_pts; uint8_t* h_st = reinterpret_cast<uint8_t*>(h_err + k->max_pts);
  if (n) {
    cudaMemcpyAsync(hp, k->d_next_pts, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, k->st);
    cudaMemcpyAsync(h_err, k->d_err, sizeof(float) * n, cudaMemcpyDeviceToHost, k->st);
    cudaMemcpyAsync(h_st, k->d_status, n, cudaMemcpyDeviceToHost, k->st);
  }
  cudaError_t e = cudaStreamSynchronize(k->st);
  if (e != cudaSuccess) return vils::fail_cuda(e, "vils_klt_track");
  cudaEventElapsedTime(&k->last_ms, k->e0, k->e1);
  if (n) {
    if (next_xy) memcpy(next_xy, hp, sizeof(float) * 2 * n);
    if (err) memcpy(err, h_err, sizeof(float) * n);
    if (status) memcpy(status, h_st, n);
  }
  return VILS_OK;
}

// FeatureTracker::readImage keeps forw_img as next call's cur_img (feature_tracker.cpp:160-164): so does the device.  next_dev: the new
// (equalised) image ALREADY ON THE DEVICE (vils_frontend_current), pitch_bytes apart.  The pyramid built for it by this call is next call's
// `prev`; only one pyramid is built per frame and no image crosses PCIe here.  The first call only loads the image (n is ignored).
int vils_klt_advance(vils_klt* k, const uint8_t* next_dev, int32_t pitch_bytes, void* ready_event, const float* prev_xy, int32_t n, float* next_xy, uint8_t* status,
                     float* err) {
  if (!k || !next_dev || pitch_bytes < k->cols || n < 0 || n > k->max_pts || (n && !prev_xy)) return vils::fail(VILS_ERR_BAD_ARG, "vils_klt_advance: bad argument");
  cudaSetDevice(k->device);
  const bool track = k->has_next;
  if (track) { for (int l = 0; l < k->levels; l++) std::swap(k->prev.img[l], k->next.img[l]); k->parity ^= 1; }
  if (ready_event) cudaStreamWaitEvent(k->st, static_cast<cudaEvent_t>(ready_event), 0);   // the producer of next_dev (vils_frontend_load) is not synchronised
  cudaMemcpy2DAsync(k->next.img[0], k->cols, next_dev, pitch_bytes, k->cols, k->rows, cudaMemcpyDeviceToDevice, k->st);
  float* hp = reinterpret_cast<float*>(k->h_stage + ((2 * (size_t)k->rows * k->cols + 15) & ~(size_t)15));
  const int nt = track ? n : 0;
  if (nt) { memcpy(hp, prev_xy, sizeof(float) * 2 * nt); cudaMemcpyAsync(k->d_prev_pts, hp, sizeof(float) * 2 * nt, cudaMemcpyHostToDevice, k->st); }
  k->n = nt;
  cudaEventRecord(k->e0, k->st);
  // the launch arguments alternate between two pointer sets (the pyramids swap roles every call): one captured graph per parity
  cudaGraphExec_t& ge = k->graph_adv[k->parity];
  if (!ge && !k->graph_adv_failed) {
    cudaGraph_t g = nullptr;
    if (cudaStreamBeginCapture(k->st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      klt_launch_advance(k);
      if (cudaStreamEndCapture(k->st, &g) != cudaSuccess || !g || cudaGraphInstantiate(&ge, g, 0) != cudaSuccess) { ge = nullptr; k->graph_adv_failed = true; }
      if (g) cudaGraphDestroy(g);
    } else k->graph_adv_failed = true;
    cudaGetLastError();
  }
  if (ge) cudaGraphLaunch(ge, k->st); else klt_launch_advance(k);
  if (nt > 0) {
    TrackParams T; T.prev = k->prev; T.next = k->next; T.prev_pts = k->d_prev_pts; T.next_pts = k->d_next_pts; T.status = k->d_status; T.err = k->d_err;
    T.n = nt; T.levels = k->levels; T.win = k->win; T.min_eig = 1e-4f; T.max_iter = 30; T.eps2 = 0.01 * 0.01;
    track_kernel<<<nt, TRACK_THREADS, 0, k->st>>>(T);
  }
  const cudaError_t le = cudaGetLastError();
  cudaEventRecord(k->e1, k->st);
  float* h_err = hp + 2 * (size_t)k->max_pts; uint8_t* h_st = reinterpret_cast<uint8_t*>(h_err + k->max_pts);
  if (nt) {
    cudaMemcpyAsync(hp, k->d_next_pts, sizeof(float) * 2 * nt, cudaMemcpyDeviceToHost, k->st);
    cudaMemcpyAsync(h_err, k->d_err, sizeof(float) * nt, cudaMemcpyDeviceToHost, k->st);
    cudaMemcpyAsync(h_st, k->d_status, nt, cudaMemcpyDeviceToHost, k->st);
  }
  const cudaError_t e = cudaStreamSynchronize(k->st);
  if (le != cudaSuccess) return vils::fail_cuda(le, "vils_klt_advance launch");
  if (e != cudaSuccess) return vils::fail_cuda(e, "vils_klt_advance");
  cudaEventElapsedTime(&k->last_ms, k->e0, k->e1);
  k->has_next = true;
  if (nt) {
    if (next_xy) memcpy(next_xy, hp, sizeof(float) * 2 * nt);
    if (err) memcpy(err, h_err, sizeof(float) * nt);
    if (status) memcpy(status, h_st, nt);
  } else if (n && status) memset(status, 0, n);              // nothing to track against yet
  return VILS_OK;
}

int vils_klt_last_device_ms(vils_klt* k, float* ms) { if (!k || !ms) return VILS_ERR_BAD_ARG; *ms = k->last_ms; return VILS_OK; }

}  // extern "C"
