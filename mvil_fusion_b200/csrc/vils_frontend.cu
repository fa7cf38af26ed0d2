// vils_frontend.cu — the rest of FeatureTracker::readImage around the LK call (feature_tracker_/src/feature_tracker.cpp:81-167):
//   CLAHE(3.0, 8x8)            :87-93   -> vils_clahe            (u8, bit-exact integer histograms / LUTs, FP32 bilinear blend)
//   setMask()                   :36-69   -> vils_set_mask         (track-count ordered greedy pick on the host, disc raster on the device)
//   goodFeaturesToTrack(...)    :149     -> vils_good_features    (min-eigenvalue map, masked max, 3x3 non-max suppression, sort on
//                                                                 the device; the inherently sequential min-distance pick on the host)
//   undistortedPoints()         :258-306 -> vils_lift_projective  (PinholeCamera::liftProjective, camera_model/.../PinholeCamera.cc:450-510)
// Third-party algorithms restated from OpenCV 4.x (imgproc/src/clahe.cpp, featureselect.cpp, corner.cpp, drawing.cpp); the executable
// oracle in tests/ is cv2 4.13 itself.  Compiled with --fmad=false: the FP32 blend of CLAHE follows OpenCV's operation order.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/vils_cabi.h"
#include "common.h"

namespace {

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
  return i;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// CLAHE (OpenCV clahe.cpp: CLAHE_CalcLut_Body + CLAHE_Interpolation_Body), 8-bit
// ---------------------------------------------------------------------------------------------------------------------------------
// One CTA per tile: 256-bin histogram in shared memory (integer atomics: exact), clip + redistribute, cumulative LUT.
__global__ void __launch_bounds__(256) clahe_lut_kernel(const uint8_t* __restrict__ src, int rows, int cols, int stride, int tiles_x, int tile_w, int tile_h,
                                                        int clip_limit, float lut_scale, uint8_t* __restrict__ lut) {
  __shared__ int hist[256];
  __shared__ int red[256];
  const int t = threadIdx.x, tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  hist[t] = 0;
  __syncthreads();
  // the image is virtually extended by BORDER_REFLECT_101 to a multiple of the tile grid (clahe.cpp: copyMakeBorder)
  for (int p = t; p < tile_w * tile_h; p += 256) {
    const int y = reflect101(ty * tile_h + p / tile_w, rows), x = reflect101(tx * tile_w + p % tile_w, cols);
    atomicAdd(&hist[src[(size_t)y * stride + x]], 1);
  }
  __syncthreads();
  if (clip_limit > 0) {
    int h = hist[t];
    red[t] = h > clip_limit ? h - clip_limit : 0;
    if (h > clip_limit) h = clip_limit;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (t < o) red[t] += red[t + o]; __syncthreads(); }
    const int clipped = red[0];
    const int batch = clipped / 256; int residual = clipped - batch * 256;
    h += batch;
    if (residual != 0) {
      const int step = max(256 / residual, 1);
      // for (i = 0; i < 256 && residual > 0; i += step, residual--) hist[i]++
      if (t % step == 0 && t / step < residual) h++;
    }
    __syncthreads();
    hist[t] = h;
  }
  __syncthreads();
  // inclusive prefix sum (Hillis-Steele on 256 ints)
  red[t] = hist[t];
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) { const int v = t >= o ? red[t - o] : 0; __syncthreads(); red[t] += v; __syncthreads(); }
  const float v = (float)red[t] * lut_scale;
  int r = __float2int_rn(v);                                  // saturate_cast<uchar>(float) = cvRound, saturated
  lut[(size_t)blockIdx.x * 256 + t] = (uint8_t)min(max(r, 0), 255);
}

__global__ void clahe_apply_kernel(const uint8_t* __restrict__ src, int rows, int cols, int stride, int tiles_x, int tiles_y, float inv_tw, float inv_th,
                                   const uint8_t* __restrict__ lut, uint8_t* __restrict__ dst, int dst_stride) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= cols || y >= rows) return;
  const float tyf = y * inv_th - 0.5f, txf = x * inv_tw - 0.5f;
  int ty1 = (int)floorf(tyf), tx1 = (int)floorf(txf);
  int ty2 = ty1 + 1, tx2 = tx1 + 1;
  const float ya = tyf - ty1, ya1 = 1.0f - ya, xa = txf - tx1, xa1 = 1.0f - xa;
  ty1 = max(ty1, 0); ty2 = min(ty2, tiles_y - 1); tx1 = max(tx1, 0); tx2 = min(tx2, tiles_x - 1);
  const int v = src[(size_t)y * stride + x];
  const uint8_t* l1 = lut + (size_t)ty1 * tiles_x * 256; const uint8_t* l2 = lut + (size_t)ty2 * tiles_x * 256;
  const float res = (l1[tx1 * 256 + v] * xa1 + l1[tx2 * 256 + v] * xa) * ya1 + (l2[tx1 * 256 + v] * xa1 + l2[tx2 * 256 + v] * xa) * ya;
  dst[(size_t)y * dst_stride + x] = (uint8_t)min(max(__float2int_rn(res), 0), 255);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// goodFeaturesToTrack (featureselect.cpp) on cornerMinEigenVal (corner.cpp), blockSize 3, Sobel 3
// ---------------------------------------------------------------------------------------------------------------------------------
constexpr int ET_X = 32, ET_Y = 16;   // output tile; the 8-bit input tile with a 2-pixel halo is staged in shared memory
__global__ void __launch_bounds__(ET_X * ET_Y) min_eig_kernel(const uint8_t* __restrict__ src, int rows, int cols, int stride, float* __restrict__ eig) {
  __shared__ uint8_t tile[ET_Y + 4][ET_X + 4];
  __shared__ float cxx[ET_Y + 2][ET_X + 2], cxy[ET_Y + 2][ET_X + 2], cyy[ET_Y + 2][ET_X + 2];
  const int tx = threadIdx.x, ty = threadIdx.y, t = ty * ET_X + tx;
  const int x0 = blockIdx.x * ET_X, y0 = blockIdx.y * ET_Y;
  for (int p = t; p < (ET_Y + 4) * (ET_X + 4); p += ET_X * ET_Y) {
    const int ly = p / (ET_X + 4), lx = p % (ET_X + 4);
    // derivative taps reflect at the IMAGE border (BORDER_REFLECT_101 of Sobel); the covariance box filter reflects again below
    tile[ly][lx] = src[(size_t)reflect101(y0 + ly - 2, rows) * stride + reflect101(x0 + lx - 2, cols)];
  }
  __syncthreads();
  const float scale = (float)(1.0 / (4.0 * 3.0) * (1.0 / 255.0));   // corner.cpp: 1 / (2^(ksize-1) * block_size), * 1/255 for 8-bit input
  for (int p = t; p < (ET_Y + 2) * (ET_X + 2); p += ET_X * ET_Y) {
    const int ly = p / (ET_X + 2), lx = p % (ET_X + 2);
    // covariance sample at image position (y0 + ly - 1, x0 + lx - 1); outside the image the box filter's REFLECT_101 applies to the
    // covariance image itself, i.e. the sample is the one of the mirrored position
    const int gy = reflect101(y0 + ly - 1, rows), gx = reflect101(x0 + lx - 1, cols);
    const int cy = gy - y0 + 2, cx = gx - x0 + 2;
    float dx, dy;
    if (cy >= 1 && cy < ET_Y + 3 && cx >= 1 && cx < ET_X + 3) {
      const int a00 = tile[cy - 1][cx - 1], a01 = tile[cy - 1][cx], a02 = tile[cy - 1][cx + 1], a10 = tile[cy][cx - 1], a12 = tile[cy][cx + 1],
                a20 = tile[cy + 1][cx - 1], a21 = tile[cy + 1][cx], a22 = tile[cy + 1][cx + 1];
      dx = (float)((a02 - a00) + 2 * (a12 - a10) + (a22 - a20)) * scale;
      dy = (float)((a20 - a00) + 2 * (a21 - a01) + (a22 - a02)) * scale;
    } else {   // mirrored position fell outside the staged tile (only for images narrower than the halo): read from global memory
      int a[3][3];
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) a[i][j] = src[(size_t)reflect101(gy + i - 1, rows) * stride + reflect101(gx + j - 1, cols)];
      dx = (float)((a[0][2] - a[0][0]) + 2 * (a[1][2] - a[1][0]) + (a[2][2] - a[2][0])) * scale;
      dy = (float)((a[2][0] - a[0][0]) + 2 * (a[2][1] - a[0][1]) + (a[2][2] - a[0][2])) * scale;
    }
    cxx[ly][lx] = dx * dx; cxy[ly][lx] = dx * dy; cyy[ly][lx] = dy * dy;
  }
  __syncthreads();
  const int x = x0 + tx, y = y0 + ty;
  if (x >= cols || y >= rows) return;
  float s[3];
  {
    float (*c[3])[ET_X + 2] = {cxx, cxy, cyy};
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const float r0 = (c[k][ty][tx] + c[k][ty][tx + 1]) + c[k][ty][tx + 2];
      const float r1 = (c[k][ty + 1][tx] + c[k][ty + 1][tx + 1]) + c[k][ty + 1][tx + 2];
      const float r2 = (c[k][ty + 2][tx] + c[k][ty + 2][tx + 1]) + c[k][ty + 2][tx + 2];
      s[k] = (r0 + r1) + r2;
    }
  }
  const float a = s[0] * 0.5f, b = s[1], c2 = s[2] * 0.5f;
  eig[(size_t)y * cols + x] = (a + c2) - sqrtf((a - c2) * (a - c2) + b * b);   // calcMinEigenVal
}

// minMaxLoc(eig, 0, &maxVal, 0, 0, mask): the minimum eigenvalue of the best corner is positive, so the IEEE bit pattern orders
__global__ void masked_max_kernel(const float* __restrict__ eig, const uint8_t* __restrict__ mask, int n, unsigned int* out) {
  unsigned int m = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float v = eig[i];
    if (v > 0.0f && (!mask || mask[i])) m = max(m, __float_as_uint(v));
  }
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// threshold(TOZERO) + dilate 3x3 + "val != 0 && val == dilated && mask" over the interior (y, x in 1..n-2)
__global__ void candidates_kernel(const float* __restrict__ eig, const uint8_t* __restrict__ mask, int rows, int cols, const unsigned int* maxbits, float quality,
                                  unsigned long long* cand, unsigned int* count, unsigned int cap, unsigned int* hist) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x < 1 || y < 1 || x >= cols - 1 || y >= rows - 1) return;
  const float thr = (float)((double)__uint_as_float(*maxbits) * (double)quality);   // threshold(eig, eig, maxVal*qualityLevel, ...): double product
  const float v = eig[(size_t)y * cols + x];
  if (!(v > thr) || (mask && !mask[(size_t)y * cols + x])) return;
  float m = v;
#pragma unroll
  for (int dy = -1; dy <= 1; dy++)
#pragma unroll
    for (int dx = -1; dx <= 1; dx++) m = fmaxf(m, eig[(size_t)(y + dy) * cols + x + dx]);
  if (v != m) return;
  const unsigned int k = atomicAdd(count, 1u);
  // sort key, descending: value first, then the address (greaterThanPtr: equal values -> higher address first)
  if (k < cap) { cand[k] = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned int)(y * cols + x); atomicAdd(&hist[__float_as_uint(v) >> 20], 1u); }
}

// Top-K selection for the greedy pick (it consumes the candidates in descending order and stops after max_corners accepts, i.e. long
// before the end of the list): one CTA finds, from the 4096-bin histogram of the keys' top 12 bits (sign, exponent, 3 mantissa bits), the
// bin down to which at least `want` candidates lie, compacts those candidates into shared memory, sorts them there (bitonic, <= 8192
// keys) and writes them out.  info[2] = number of sorted keys written (0 = does not fit: caller falls back to the full sort).
constexpr unsigned int TOPK_SMEM = 8192;
__global__ void __launch_bounds__(1024) topk_sort_kernel(const unsigned long long* __restrict__ cand, const unsigned int* __restrict__ hist, unsigned int want,
                                                         unsigned int cap, unsigned long long* __restrict__ top, unsigned int* __restrict__ info) {
  extern __shared__ unsigned long long sk[];
  __shared__ unsigned int cut_bin, n_sel;
  const unsigned int n = min(info[0], cap), t = threadIdx.x;
  __shared__ unsigned int part[1024];
  {
    // suffix sums over the 4096 bins (4 per thread): the cut is the highest bin b with  sum_{bins >= b} >= want
    unsigned int hb[4]; unsigned int mine = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) { hb[b] = hist[4 * t + b]; mine += hb[b]; }
    part[t] = mine;
    if (t == 0) { cut_bin = 0; n_sel = 0; }
    __syncthreads();
    for (unsigned int o = 1; o < 1024; o <<= 1) { const unsigned int v = t + o < 1024 ? part[t + o] : 0; __syncthreads(); part[t] += v; __syncthreads(); }
    const unsigned int all = part[0];
    unsigned int acc = part[t] - mine;                      // candidates in the bins above this thread's four
#pragma unroll
    for (int b = 3; b >= 0; b--) {
      const unsigned int before = acc; acc += hb[b];
      if (before < want && acc >= want) { cut_bin = 4 * t + b; info[3] = acc; }
    }
    if (t == 0 && all < want) info[3] = all;                // fewer candidates than wanted: take them all (cut_bin = 0)
  }
  __syncthreads();
  const unsigned int total_sel = info[3];
  if (total_sel > TOPK_SMEM) { if (t == 0) info[2] = 0; return; }
  unsigned int np2 = 1; while (np2 < total_sel) np2 <<= 1;
  for (unsigned int i = t; i < np2; i += blockDim.x) sk[i] = 0ull;
  __syncthreads();
  for (unsigned int i = t; i < n; i += blockDim.x) {
    const unsigned long long k = cand[i];
    if ((unsigned int)(k >> 52) >= cut_bin) sk[atomicAdd(&n_sel, 1u)] = k;
  }
  __syncthreads();
  for (unsigned int k = 2; k <= np2; k <<= 1)
    for (unsigned int j = k >> 1; j > 0; j >>= 1) {
      for (unsigned int i = t; i < np2; i += blockDim.x) {
        const unsigned int l = i ^ j;
        if (l > i) {
          const unsigned long long a = sk[i], b = sk[l];
          const bool desc = (i & k) == 0;
          if (desc ? a < b : a > b) { sk[i] = b; sk[l] = a; }
        }
      }
      __syncthreads();
    }
  for (unsigned int i = t; i < total_sel; i += blockDim.x) top[i] = sk[i];
  if (t == 0) info[2] = total_sel;
}

// Single-CTA bitonic sort (descending) of n_pad = 2^k 64-bit keys in global memory; candidates are a few thousand at most.
__global__ void __launch_bounds__(1024) sort_desc_kernel(unsigned long long* keys, const unsigned int* count, unsigned int cap) {
  unsigned int n = min(*count, cap), np2 = 1;
  while (np2 < n) np2 <<= 1;
  for (unsigned int i = n + threadIdx.x; i < np2; i += blockDim.x) keys[i] = 0ull;
  __syncthreads();
  for (unsigned int k = 2; k <= np2; k <<= 1)
    for (unsigned int j = k >> 1; j > 0; j >>= 1) {
      for (unsigned int i = threadIdx.x; i < np2; i += blockDim.x) {
        const unsigned int l = i ^ j;
        if (l > i) {
          const unsigned long long a = keys[i], b = keys[l];
          const bool desc = (i & k) == 0;
          if (desc ? a < b : a > b) { keys[i] = b; keys[l] = a; }
        }
      }
      __syncthreads();
    }
}

// setMask(): filled cv::circle(mask, p, radius, 0, -1) for every kept point. hw[d] = half-width of the rasterised disc at row offset d.
__global__ void mask_discs_kernel(uint8_t* mask, int rows, int cols, const int* __restrict__ centers, int n, int radius, const int* __restrict__ hw) {
  const int k = blockIdx.x;
  if (k >= n) return;
  const int cx = centers[2 * k], cy = centers[2 * k + 1], side = 2 * radius + 1;
  for (int p = threadIdx.x; p < side * side; p += blockDim.x) {
    const int dy = p / side - radius, dx = p % side - radius;
    const int y = cy + dy, x = cx + dx;
    if (y < 0 || y >= rows || x < 0 || x >= cols) continue;
    if (abs(dx) <= hw[abs(dy)]) mask[(size_t)y * cols + x] = 0;
  }
}

// PinholeCamera::liftProjective (PinholeCamera.cc:450-510) with the recursive distortion model (n = 8), ::distortion (:646-662)
__global__ void lift_kernel(const float* __restrict__ uv, int n, double fx, double fy, double cx, double cy, double k1, double k2, double p1, double p2,
                            int no_distortion, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double inv_K11 = 1.0 / fx, inv_K13 = -cx / fx, inv_K22 = 1.0 / fy, inv_K23 = -cy / fy;
  const double mx_d = inv_K11 * (double)uv[2 * i] + inv_K13, my_d = inv_K22 * (double)uv[2 * i + 1] + inv_K23;
  double mx_u = mx_d, my_u = my_d;
  if (!no_distortion) {
    double ux = mx_d, uy = my_d;
    for (int it = 0; it < 8; it++) {
      const double mx2 = ux * ux, my2 = uy * uy, mxy = ux * uy, rho2 = mx2 + my2, rad = k1 * rho2 + k2 * rho2 * rho2;
      const double dux = ux * rad + 2.0 * p1 * mxy + p2 * (rho2 + 2.0 * mx2), duy = uy * rad + 2.0 * p2 * mxy + p1 * (rho2 + 2.0 * my2);
      ux = mx_d - dux; uy = my_d - duy;
    }
    mx_u = ux; my_u = uy;
  }
  out[3 * i] = mx_u; out[3 * i + 1] = my_u; out[3 * i + 2] = 1.0;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// rejectWithF (feature_tracker.cpp:169-202): cv::findFundamentalMat(un_cur, un_forw, FM_RANSAC, F_THRESHOLD, 0.99, status).
// RANSAC with ONE THREAD PER HYPOTHESIS (1024 = OpenCV's maxIters rounded up, no early exit so the result is deterministic):
// 8 distinct correspondences -> normalised 8-point system (Hartley normalisation from all points) -> null vector by full-pivot
// Gauss-Jordan -> F in pixel units -> inlier count with OpenCV's symmetric epipolar distance (fundam.cpp computeError:
// max(d(x2, F x1)^2, d(x1, F^T x2)^2) <= threshold^2).  The best hypothesis classifies the points.  OpenCV draws 7-point samples from
// its own RNG, so the masks agree on clear inliers / outliers, not bit for bit (see tests).
// ---------------------------------------------------------------------------------------------------------------------------------
constexpr int RANSAC_H = 1024;
__device__ __forceinline__ unsigned int rng_next(unsigned int& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }
__device__ __forceinline__ double epi_err(const double* F, double x1, double y1, double x2, double y2) {
  const double a = F[0] * x1 + F[1] * y1 + F[2], b = F[3] * x1 + F[4] * y1 + F[5], c = F[6] * x1 + F[7] * y1 + F[8];
  const double s2 = 1.0 / (a * a + b * b), d2 = x2 * a + y2 * b + c;
  const double a2 = F[0] * x2 + F[3] * y2 + F[6], b2 = F[1] * x2 + F[4] * y2 + F[7], c2 = F[2] * x2 + F[5] * y2 + F[8];
  const double s1 = 1.0 / (a2 * a2 + b2 * b2), d1 = x1 * a2 + y1 * b2 + c2;
  return fmax(d1 * d1 * s1, d2 * d2 * s2);
}
__global__ void __launch_bounds__(RANSAC_H) ransac_f_kernel(const float* __restrict__ p1, const float* __restrict__ p2, int n, double thresh2, unsigned int seed,
                                                            uint8_t* __restrict__ status, int* __restrict__ info, double* __restrict__ F_out) {
  extern __shared__ double sm[];                 // x1[n] y1[n] x2[n] y2[n]
  __shared__ double nrm[6];                      // cx1 cy1 s1 cx2 cy2 s2
  __shared__ int cnt[RANSAC_H];
  __shared__ double Fbest[9];
  double* X1 = sm; double* Y1 = sm + n; double* X2 = sm + 2 * n; double* Y2 = sm + 3 * n;
  const int t = threadIdx.x;
  for (int i = t; i < n; i += RANSAC_H) { X1[i] = p1[2 * i]; Y1[i] = p1[2 * i + 1]; X2[i] = p2[2 * i]; Y2[i] = p2[2 * i + 1]; }
  __syncthreads();
  if (t < 2) {                                   // Hartley normalisation: centroid, mean distance sqrt(2)
    const double* X = t ? X2 : X1; const double* Y = t ? Y2 : Y1;
    double cx = 0, cy = 0;
    for (int i = 0; i < n; i++) { cx += X[i]; cy += Y[i]; }
    cx /= n; cy /= n;
    double d = 0;
    for (int i = 0; i < n; i++) d += sqrt((X[i] - cx) * (X[i] - cx) + (Y[i] - cy) * (Y[i] - cy));
    nrm[3 * t] = cx; nrm[3 * t + 1] = cy; nrm[3 * t + 2] = d > 0 ? sqrt(2.0) * n / d : 1.0;
  }
  __syncthreads();
  // ---- hypothesis t
  int my = 0; double F[9];
  {
    unsigned int s = seed * 2654435761u + 0x9e3779b9u * (t + 1); rng_next(s); rng_next(s);
    int idx[8];
    for (int k = 0; k < 8; k++) {
      int c; bool dup;
      do { c = (int)(rng_next(s) % (unsigned int)n); dup = false; for (int j = 0; j < k; j++) dup |= idx[j] == c; } while (dup);
      idx[k] = c;
    }
    double A[8][9]; int perm[9];
    const double c1x = nrm[0], c1y = nrm[1], s1 = nrm[2], c2x = nrm[3], c2y = nrm[4], s2 = nrm[5];
    for (int k = 0; k < 8; k++) {
      const double x1 = (X1[idx[k]] - c1x) * s1, y1 = (Y1[idx[k]] - c1y) * s1, x2 = (X2[idx[k]] - c2x) * s2, y2 = (Y2[idx[k]] - c2y) * s2;
      A[k][0] = x2 * x1; A[k][1] = x2 * y1; A[k][2] = x2; A[k][3] = y2 * x1; A[k][4] = y2 * y1; A[k][5] = y2; A[k][6] = x1; A[k][7] = y1; A[k][8] = 1.0;
    }
    for (int j = 0; j < 9; j++) perm[j] = j;
    bool ok = true;
    for (int k = 0; k < 8 && ok; k++) {          // full-pivot Gauss-Jordan: A -> [I | c] in permuted columns
      int pi = k, pj = k; double best = 0;
      for (int i = k; i < 8; i++) for (int j = k; j < 9; j++) if (fabs(A[i][j]) > best) { best = fabs(A[i][j]); pi = i; pj = j; }
      if (best < 1e-10) { ok = false; break; }
      if (pi != k) for (int j = 0; j < 9; j++) { const double v = A[k][j]; A[k][j] = A[pi][j]; A[pi][j] = v; }
      if (pj != k) { for (int i = 0; i < 8; i++) { const double v = A[i][k]; A[i][k] = A[i][pj]; A[i][pj] = v; } const int v = perm[k]; perm[k] = perm[pj]; perm[pj] = v; }
      const double inv = 1.0 / A[k][k];
      for (int j = 0; j < 9; j++) A[k][j] *= inv;
      for (int i = 0; i < 8; i++) if (i != k) { const double f = A[i][k]; if (f != 0.0) for (int j = 0; j < 9; j++) A[i][j] -= f * A[k][j]; }
    }
    if (ok) {
      double f[9];
      f[perm[8]] = 1.0;
      for (int k = 0; k < 8; k++) f[perm[k]] = -A[k][8];
      // F = T2^T Fn T1 with T = [s 0 -s cx; 0 s -s cy; 0 0 1]
      double G[9];                               // Fn T1
      for (int r = 0; r < 3; r++) { G[3 * r] = f[3 * r] * s1; G[3 * r + 1] = f[3 * r + 1] * s1; G[3 * r + 2] = f[3 * r + 2] - s1 * (f[3 * r] * c1x + f[3 * r + 1] * c1y); }
      for (int c = 0; c < 3; c++) { F[c] = s2 * G[c]; F[3 + c] = s2 * G[3 + c]; F[6 + c] = G[6 + c] - s2 * (c2x * G[c] + c2y * G[3 + c]); }
      double nf = 0; for (int k = 0; k < 9; k++) nf += F[k] * F[k];
      nf = 1.0 / sqrt(nf); for (int k = 0; k < 9; k++) F[k] *= nf;
      for (int i = 0; i < n; i++) my += epi_err(F, X1[i], Y1[i], X2[i], Y2[i]) <= thresh2;
    }
  }
  cnt[t] = my;
  __syncthreads();
  __shared__ int best_t;
  if (t == 0) { int b = 0; for (int h = 1; h < RANSAC_H; h++) if (cnt[h] > cnt[b]) b = h; best_t = b; info[0] = cnt[b]; info[1] = b; }
  __syncthreads();
  if (t == best_t) for (int k = 0; k < 9; k++) { Fbest[k] = F[k]; F_out[k] = F[k]; }
  __syncthreads();
  const bool none = cnt[best_t] < 8;
  for (int i = t; i < n; i += RANSAC_H) status[i] = none ? 0 : (epi_err(Fbest, X1[i], Y1[i], X2[i], Y2[i]) <= thresh2);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// FeatureManager::triangulate (vils_estimator/src/feature_manager.cpp:214-268): per feature, the 2L x 4 DLT matrix of its L observations
// relative to the anchor camera, right singular vector of the smallest singular value (Eigen::JacobiSVD in the reference), depth =
// V[2] / V[3].  One thread per feature: one-sided (Hestenes) Jacobi SVD — the same rotation family as Eigen's two-sided Jacobi and, like
// it, accurate to the conditioning of A (no A^T A squaring).
// ---------------------------------------------------------------------------------------------------------------------------------
constexpr int TRI_MAX_OBS = 32;
__global__ void triangulate_kernel(int n_feat, const int32_t* __restrict__ start, const int32_t* __restrict__ off, const double* __restrict__ pts,
                                   int n_kf, const double* __restrict__ Ps, const double* __restrict__ Rs, const double* __restrict__ ex /* tic(3) ric(9) */,
                                   double init_depth, double* __restrict__ depth) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_feat) return;
  const int L = min(off[f + 1] - off[f], TRI_MAX_OBS), i0 = start[f];
  double A[2 * TRI_MAX_OBS][4], V[4][4];
  const double* tic = ex; const double* ric = ex + 3;
  auto cam = [&](int k, double R[9], double t[3]) {          // R = Rs[k] ric, t = Ps[k] + Rs[k] tic
    const double* Rk = Rs + 9 * k;
    for (int a = 0; a < 3; a++) {
      t[a] = Ps[3 * k + a] + Rk[3 * a] * tic[0] + Rk[3 * a + 1] * tic[1] + Rk[3 * a + 2] * tic[2];
      for (int b = 0; b < 3; b++) R[3 * a + b] = Rk[3 * a] * ric[b] + Rk[3 * a + 1] * ric[3 + b] + Rk[3 * a + 2] * ric[6 + b];
    }
  };
  double R0[9], t0[3]; cam(min(i0, n_kf - 1), R0, t0);
  for (int j = 0; j < L; j++) {
    double R1[9], t1[3]; cam(min(i0 + j, n_kf - 1), R1, t1);
    // R = R0^T R1, t = R0^T (t1 - t0);  P = [R^T | -R^T t]
    double R[9], t[3], P[3][4];
    for (int a = 0; a < 3; a++) {
      t[a] = R0[a] * (t1[0] - t0[0]) + R0[3 + a] * (t1[1] - t0[1]) + R0[6 + a] * (t1[2] - t0[2]);
      for (int b = 0; b < 3; b++) R[3 * a + b] = R0[a] * R1[b] + R0[3 + a] * R1[3 + b] + R0[6 + a] * R1[6 + b];
    }
    for (int a = 0; a < 3; a++) { for (int b = 0; b < 3; b++) P[a][b] = R[3 * b + a]; P[a][3] = -(R[a] * t[0] + R[3 + a] * t[1] + R[6 + a] * t[2]); }
    const double* p = pts + 3 * (size_t)(off[f] + j);
    const double inv = 1.0 / sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
    const double fx = p[0] * inv, fy = p[1] * inv, fz = p[2] * inv;   // it_per_frame.point.normalized()
    for (int c = 0; c < 4; c++) { A[2 * j][c] = fx * P[2][c] - fz * P[0][c]; A[2 * j + 1][c] = fy * P[2][c] - fz * P[1][c]; }
  }
  for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) V[a][b] = a == b ? 1.0 : 0.0;
  const int m = 2 * L;
  for (int sweep = 0; sweep < 30; sweep++) {
    double offn = 0;
    for (int p = 0; p < 3; p++)
      for (int q = p + 1; q < 4; q++) {
        double app = 0, aqq = 0, apq = 0;
        for (int r = 0; r < m; r++) { app += A[r][p] * A[r][p]; aqq += A[r][q] * A[r][q]; apq += A[r][p] * A[r][q]; }
        if (fabs(apq) <= 1e-300 || fabs(apq) <= 1e-17 * sqrt(app * aqq)) continue;
        offn = fmax(offn, fabs(apq) / sqrt(app * aqq));
        const double zeta = (aqq - app) / (2.0 * apq);
        const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + tt * tt), s_ = c * tt;
        for (int r = 0; r < m; r++) { const double x = A[r][p], y = A[r][q]; A[r][p] = c * x - s_ * y; A[r][q] = s_ * x + c * y; }
        for (int r = 0; r < 4; r++) { const double x = V[r][p], y = V[r][q]; V[r][p] = c * x - s_ * y; V[r][q] = s_ * x + c * y; }
      }
    if (offn < 1e-15) break;
  }
  int best = 0; double bn = 1e300;
  for (int c = 0; c < 4; c++) { double nn = 0; for (int r = 0; r < m; r++) nn += A[r][c] * A[r][c]; if (nn < bn) { bn = nn; best = c; } }
  double d = V[2][best] / V[3][best];
  if (!(d >= 0)) d = init_depth;                              // triangulation failed (:262-266; NaN also lands here)
  depth[f] = d;
}

// Integer midpoint circle of OpenCV's Circle() (drawing.cpp), fill = 1: half-width of the filled disc per row offset.
void disc_half_widths(int radius, std::vector<int>& hw) {
  hw.assign(radius + 1, -1);
  int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
  while (dx >= dy) {
    hw[dy] = std::max(hw[dy], dx);   // rows center.y +- dy span +-dx
    hw[dx] = std::max(hw[dx], dy);   // rows center.y +- dx span +-dy
    dy++; err += plus; plus += 2;
    const int mask = (err <= 0) - 1;
    err -= minus & mask; dx += mask; minus -= mask & 2;
  }
}
inline int cv_round(double v) { return (int)std::nearbyint(v); }   // cvRound: round half to even (default FP environment)

}  // namespace

struct vils_frontend {
  int rows = 0, cols = 0, device = 0, max_pts = 0;
  cudaStream_t st = nullptr; cudaEvent_t e0 = nullptr, e1 = nullptr;
  uint8_t* d_src = nullptr; uint8_t* d_dst = nullptr; uint8_t* d_lut = nullptr; uint8_t* d_mask = nullptr; uint8_t* h_img = nullptr;
  float* d_eig = nullptr; unsigned long long* d_cand = nullptr; unsigned long long* h_cand = nullptr; unsigned int* d_cnt = nullptr; unsigned int* h_cnt = nullptr; unsigned int* d_hist = nullptr; unsigned long long* d_top = nullptr;
  int* d_centers = nullptr; int* d_hw = nullptr; float* d_uv = nullptr; double* d_ray = nullptr;
  unsigned int cand_cap = 0; int hw_radius = -1;
  bool mask_valid = false;
  float last_ms = 0;
  cudaEvent_t ev_ready = nullptr;   // recorded behind the last kernel of vils_frontend_load: consumers on other streams wait on it
  const uint8_t* d_cur = nullptr;   // vils_frontend_load: the image this frame's steps work on (d_dst when equalised, d_src otherwise), resident
};

extern "C" {

int vils_frontend_create(int32_t rows, int32_t cols, int32_t max_pts, int32_t device, vils_frontend** out) {
  if (!out || rows < 8 || cols < 8 || max_pts <= 0) return vils::fail(VILS_ERR_BAD_ARG, "vils_frontend_create: bad argument");
  int st = vils::require_device(device); if (st) return st;
  vils_frontend* f = new vils_frontend(); f->rows = rows; f->cols = cols; f->device = device; f->max_pts = max_pts;
  const size_t px = (size_t)rows * cols;
  f->cand_cap = 1u << 16;
  cudaError_t e = cudaStreamCreateWithFlags(&f->st, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&f->e0);
  if (e == cudaSuccess) e = cudaEventCreate(&f->e1);
  if (e == cudaSuccess) e = cudaMalloc(&f->d_src, px);
  if (e == cudaSuccess) e = cudaMalloc(&f->d_dst, px);
  if (e == cudaSuccess) e = cudaMalloc(&f->d_mask, px);
  if (e == cudaSuccess) e = cudaMalloc(&f->d_lut, 256 * 64 * 4);
  if (e == cudaSuccess) e = cudaMalloc(&f->d_eig, px * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&f->d_cand, sizeof(unsigned long long) * f->cand_cap);
  if (e == cudaSuccess) e = cudaMalloc(&f->d_cnt, 4 * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMalloc(&f->d_hist, 4096 * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMalloc(&f->d_top, TOPK_SMEM * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(topk_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(TOPK_SMEM * sizeof(unsigned long long)));
  if (e == cudaSuccess) e = cudaMalloc(&f->d_centers, sizeof(int) * 2 * max_pts);
  if (e == cudaSuccess) e = cudaMalloc(&f->d_hw, sizeof(int) * 1024);
  if (e == cudaSuccess) e = cudaMalloc(&f->d_uv, sizeof(float) * 2 * max_pts);
  if (e == cudaSuccess) e = cudaMalloc(&f->d_ray, sizeof(double) * 3 * max_pts);
  if (e == cudaSuccess) e = cudaMallocHost(&f->h_img, px);
  if (e == cudaSuccess) e = cudaMallocHost(&f->h_cand, sizeof(unsigned long long) * f->cand_cap);
  if (e == cudaSuccess) e = cudaMallocHost(&f->h_cnt, 4 * sizeof(unsigned int));
  if (e != cudaSuccess) { vils_frontend_destroy(f); return vils::fail_cuda(e, "vils_frontend_create"); }
  *out = f; return VILS_OK;
}

void vils_frontend_destroy(vils_frontend* f) {
  if (!f) return;
  cudaSetDevice(f->device);
  if (f->st) cudaStreamSynchronize(f->st);
  if (f->ev_ready) cudaEventDestroy(f->ev_ready);
  cudaFree(f->d_src); cudaFree(f->d_dst); cudaFree(f->d_mask); cudaFree(f->d_lut); cudaFree(f->d_eig); cudaFree(f->d_cand); cudaFree(f->d_cnt); cudaFree(f->d_hist); cudaFree(f->d_top);
  cudaFree(f->d_centers); cudaFree(f->d_hw); cudaFree(f->d_uv); cudaFree(f->d_ray); cudaFreeHost(f->h_img); cudaFreeHost(f->h_cand); cudaFreeHost(f->h_cnt);
  if (f->e0) cudaEventDestroy(f->e0);
  if (f->e1) cudaEventDestroy(f->e1);
  if (f->st) cudaStreamDestroy(f->st);
  delete f;
}

static int fe_upload(vils_frontend* f, const uint8_t* img, int stride, uint8_t* dst) {
  if (f->ev_ready) cudaEventSynchronize(f->ev_ready);        // the pinned staging image may still be in flight from an unsynchronised vils_frontend_load
  for (int y = 0; y < f->rows; y++) memcpy(f->h_img + (size_t)y * f->cols, img + (size_t)y * stride, f->cols);
  cudaError_t e = cudaMemcpyAsync(dst, f->h_img, (size_t)f->rows * f->cols, cudaMemcpyHostToDevice, f->st);
  return e == cudaSuccess ? VILS_OK : vils::fail_cuda(e, "frontend upload");
}

static void clahe_launch(vils_frontend* f, double clip_limit, int tiles_x, int tiles_y);

// Device-resident form of the first step of readImage (feature_tracker.cpp:87-93): upload the raw image ONCE, equalise it on the device when
// asked (EQUALIZE), and keep the result there for vils_klt_advance / vils_good_features_resident.  Nothing comes back to the host.
int vils_frontend_load(vils_frontend* f, const uint8_t* src, int32_t stride, int32_t equalize, double clip_limit, int32_t tiles_x, int32_t tiles_y) {
  if (!f || !src || stride < f->cols || (equalize && (tiles_x <= 0 || tiles_y <= 0 || tiles_x * tiles_y > 256))) return vils::fail(VILS_ERR_BAD_ARG, "vils_frontend_load: bad argument");
  cudaSetDevice(f->device);
  int st = fe_upload(f, src, stride, f->d_src); if (st) return st;
  cudaEventRecord(f->e0, f->st);
  if (equalize) clahe_launch(f, clip_limit, tiles_x, tiles_y);
  const cudaError_t le = cudaGetLastError();
  cudaEventRecord(f->e1, f->st);
  // NOT synchronised: the consumers (the KLT handle's stream, this handle's own later calls) order themselves behind ev_ready, so the upload, the
  // equalisation, the pyramid and the tracking of one frame run back to back with a single host synchronisation at the end of vils_klt_advance
  if (!f->ev_ready) cudaEventCreateWithFlags(&f->ev_ready, cudaEventDisableTiming);
  const cudaError_t e = cudaEventRecord(f->ev_ready, f->st);
  if (le != cudaSuccess) return vils::fail_cuda(le, "vils_frontend_load launch");
  if (e != cudaSuccess) return vils::fail_cuda(e, "vils_frontend_load");
  f->d_cur = equalize ? f->d_dst : f->d_src;
  return VILS_OK;
}
int vils_frontend_current(vils_frontend* f, const uint8_t** image_dev, int32_t* pitch_bytes, void** ready_event) {
  if (!f || !image_dev || !pitch_bytes || !f->d_cur) return vils::fail(VILS_ERR_BAD_ARG, "vils_frontend_current: call vils_frontend_load first");
  *image_dev = f->d_cur; *pitch_bytes = f->cols;
  if (ready_event) *ready_event = f->ev_ready;
  return VILS_OK;
}

static void clahe_launch(vils_frontend* f, double clip_limit, int tiles_x, int tiles_y) {
  const int rows = f->rows, cols = f->cols;
  const int ext_w = cols % tiles_x == 0 ? cols : cols + (tiles_x - cols % tiles_x), ext_h = rows % tiles_y == 0 ? rows : rows + (tiles_y - rows % tiles_y);
  const int tw = ext_w / tiles_x, th = ext_h / tiles_y, area = tw * th;
  int clip = 0;
  if (clip_limit > 0.0) { clip = (int)(clip_limit * area / 256); clip = std::max(clip, 1); }
  const float lut_scale = (float)255 / area;
  clahe_lut_kernel<<<tiles_x * tiles_y, 256, 0, f->st>>>(f->d_src, rows, cols, cols, tiles_x, tw, th, clip, lut_scale, f->d_lut);
  dim3 B(32, 8), G((cols + 31) / 32, (rows + 7) / 8);
  clahe_apply_kernel<<<G, B, 0, f->st>>>(f->d_src, rows, cols, cols, tiles_x, tiles_y, 1.0f / tw, 1.0f / th, f->d_lut, f->d_dst, cols);
}

int vils_clahe(vils_frontend* f, const uint8_t* src, int32_t stride, double clip_limit, int32_t tiles_x, int32_t tiles_y, uint8_t* dst, int32_t dst_stride) {
  if (!f || !src || !dst || stride < f->cols || dst_stride < f->cols || tiles_x <= 0 || tiles_y <= 0 || tiles_x * tiles_y > 256)
    return vils::fail(VILS_ERR_BAD_ARG, "vils_clahe: bad argument");
  cudaSetDevice(f->device);
  int st = fe_upload(f, src, stride, f->d_src); if (st) return st;
  const int rows = f->rows, cols = f->cols;
  // clahe.cpp: tiles cover the image extended to a multiple of the grid
  const int ext_w = cols % tiles_x == 0 ? cols : cols + (tiles_x - cols % tiles_x), ext_h = rows % tiles_y == 0 ? rows : rows + (tiles_y - rows % tiles_y);
  const int tw = ext_w / tiles_x, th = ext_h / tiles_y, area = tw * th;
  int clip = 0;
  if (clip_limit > 0.0) { clip = (int)(clip_limit * area / 256); clip = std::max(clip, 1); }
  const float lut_scale = (float)255 / area;
  cudaEventRecord(f->e0, f->st);
  clahe_lut_kernel<<<tiles_x * tiles_y, 256, 0, f->st>>>(f->d_src, rows, cols, cols, tiles_x, tw, th, clip, lut_scale, f->d_lut);
  dim3 B(32, 8), G((cols + 31) / 32, (rows + 7) / 8);
  clahe_apply_kernel<<<G, B, 0, f->st>>>(f->d_src, rows, cols, cols, tiles_x, tiles_y, 1.0f / tw, 1.0f / th, f->d_lut, f->d_dst, cols);
  cudaEventRecord(f->e1, f->st);
  cudaMemcpyAsync(f->h_img, f->d_dst, (size_t)rows * cols, cudaMemcpyDeviceToHost, f->st);
  cudaError_t e = cudaStreamSynchronize(f->st);
  if (e != cudaSuccess) return vils::fail_cuda(e, "vils_clahe");
  cudaEventElapsedTime(&f->last_ms, f->e0, f->e1);
  for (int y = 0; y < rows; y++) memcpy(dst + (size_t)y * dst_stride, f->h_img + (size_t)y * cols, cols);
  return VILS_OK;
}

// FeatureTracker::setMask (:36-69): points in descending track_cnt order; a point survives if its pixel is still unmasked, then blanks a disc
// of `radius` around itself.  keep_idx receives the surviving indices (in pick order).  The device mask is kept for vils_good_features.
int vils_set_mask(vils_frontend* f, const float* xy, const int32_t* track_cnt, int32_t n, int32_t radius, int32_t* keep_idx, int32_t* n_keep) {
  if (!f || n < 0 || n > f->max_pts || (n && (!xy || !track_cnt)) || radius < 0 || radius > 1000 || !keep_idx || !n_keep)
    return vils::fail(VILS_ERR_BAD_ARG, "vils_set_mask: bad argument");
  cudaSetDevice(f->device);
  std::vector<int> hw; disc_half_widths(radius, hw);
  std::vector<int> order(n);
  for (int i = 0; i < n; i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return track_cnt[a] > track_cnt[b]; });
  std::vector<int> centers; centers.reserve(2 * n);
  int nk = 0;
  for (int oi = 0; oi < n; oi++) {
    const int i = order[oi];
    const int px = cv_round(xy[2 * i]), py = cv_round(xy[2 * i + 1]);   // mask.at<uchar>(Point2f) / cv::circle(Point2f): saturate_cast<int>
    bool masked = px < 0 || py < 0 || px >= f->cols || py >= f->rows;   // the reference would read out of bounds; such points are dropped before (inBorder)
    for (size_t k = 0; !masked && k < centers.size(); k += 2) {
      const int dx = std::abs(px - centers[k]), dy = std::abs(py - centers[k + 1]);
      if (dy <= radius && dx <= hw[dy]) masked = true;
    }
    if (masked) continue;
    keep_idx[nk++] = i; centers.push_back(px); centers.push_back(py);
  }
  *n_keep = nk;
  cudaMemsetAsync(f->d_mask, 255, (size_t)f->rows * f->cols, f->st);
  if (nk) {
    cudaMemcpyAsync(f->d_centers, centers.data(), sizeof(int) * 2 * nk, cudaMemcpyHostToDevice, f->st);
    cudaMemcpyAsync(f->d_hw, hw.data(), sizeof(int) * (radius + 1), cudaMemcpyHostToDevice, f->st);
    mask_discs_kernel<<<nk, 256, 0, f->st>>>(f->d_mask, f->rows, f->cols, f->d_centers, nk, radius, f->d_hw);
  }
  cudaError_t e = cudaStreamSynchronize(f->st);   // centers / hw are stack-lifetime host buffers
  if (e != cudaSuccess) return vils::fail_cuda(e, "vils_set_mask");
  f->mask_valid = true;
  return VILS_OK;
}

int vils_get_mask(vils_frontend* f, uint8_t* mask, int32_t stride) {
  if (!f || !mask || stride < f->cols || !f->mask_valid) return vils::fail(VILS_ERR_BAD_ARG, "vils_get_mask: no mask");
  cudaSetDevice(f->device);
  cudaMemcpyAsync(f->h_img, f->d_mask, (size_t)f->rows * f->cols, cudaMemcpyDeviceToHost, f->st);
  cudaError_t e = cudaStreamSynchronize(f->st);
  if (e != cudaSuccess) return vils::fail_cuda(e, "vils_get_mask");
  for (int y = 0; y < f->rows; y++) memcpy(mask + (size_t)y * stride, f->h_img + (size_t)y * f->cols, f->cols);
  return VILS_OK;
}

// cv::goodFeaturesToTrack(img, corners, max_corners, quality, min_distance, mask) with the defaults of the call site (:149): blockSize 3,
// Sobel 3, min-eigenvalue score.  use_mask: 0 = no mask, 1 = the device mask left by vils_set_mask.
static int good_features_impl(vils_frontend* f, const uint8_t* d_img, int32_t max_corners, double quality, double min_distance, int32_t use_mask, float* xy_out, int32_t* n_out);

int vils_good_features(vils_frontend* f, const uint8_t* img, int32_t stride, int32_t max_corners, double quality, double min_distance, int32_t use_mask,
                       float* xy_out, int32_t* n_out) {
  if (!f || !img || stride < f->cols || !xy_out || !n_out || quality <= 0 || min_distance < 0 || (use_mask && !f->mask_valid))
    return vils::fail(VILS_ERR_BAD_ARG, "vils_good_features: bad argument");
  cudaSetDevice(f->device);
  *n_out = 0;
  if (max_corners == 0) return VILS_OK;
  int st = fe_upload(f, img, stride, f->d_src); if (st) return st;
  f->d_cur = nullptr;                                          // d_src no longer holds the frame loaded by vils_frontend_load
  return good_features_impl(f, f->d_src, max_corners, quality, min_distance, use_mask, xy_out, n_out);
}
// The same on the image vils_frontend_load left on the device (no upload).
int vils_good_features_resident(vils_frontend* f, int32_t max_corners, double quality, double min_distance, int32_t use_mask, float* xy_out, int32_t* n_out) {
  if (!f || !f->d_cur || !xy_out || !n_out || quality <= 0 || min_distance < 0 || (use_mask && !f->mask_valid))
    return vils::fail(VILS_ERR_BAD_ARG, "vils_good_features_resident: bad argument (vils_frontend_load first)");
  cudaSetDevice(f->device);
  *n_out = 0;
  if (max_corners == 0) return VILS_OK;
  return good_features_impl(f, f->d_cur, max_corners, quality, min_distance, use_mask, xy_out, n_out);
}
static int good_features_impl(vils_frontend* f, const uint8_t* d_img, int32_t max_corners, double quality, double min_distance, int32_t use_mask, float* xy_out, int32_t* n_out) {
  const int rows = f->rows, cols = f->cols;
  const uint8_t* mask = use_mask ? f->d_mask : nullptr;
  cudaEventRecord(f->e0, f->st);
  cudaMemsetAsync(f->d_cnt, 0, 4 * sizeof(unsigned int), f->st);
  cudaMemsetAsync(f->d_hist, 0, 4096 * sizeof(unsigned int), f->st);
  min_eig_kernel<<<dim3((cols + ET_X - 1) / ET_X, (rows + ET_Y - 1) / ET_Y), dim3(ET_X, ET_Y), 0, f->st>>>(d_img, rows, cols, cols, f->d_eig);
  masked_max_kernel<<<148 * 2, 256, 0, f->st>>>(f->d_eig, mask, rows * cols, f->d_cnt + 1);
  candidates_kernel<<<dim3((cols + 31) / 32, (rows + 7) / 8), dim3(32, 8), 0, f->st>>>(f->d_eig, mask, rows, cols, f->d_cnt + 1, (float)quality, f->d_cand, f->d_cnt,
                                                                                        f->cand_cap, f->d_hist);
  // the greedy pick below stops after max_corners accepts: sort only the strongest few thousand candidates (all of them if max_corners <= 0)
  const unsigned int want = max_corners > 0 ? (unsigned int)std::min<long long>(TOPK_SMEM / 2, 16LL * max_corners + 256) : TOPK_SMEM + 1;
  topk_sort_kernel<<<1, 1024, TOPK_SMEM * sizeof(unsigned long long), f->st>>>(f->d_cand, f->d_hist, want, f->cand_cap, f->d_top, f->d_cnt);
  cudaEventRecord(f->e1, f->st);
  cudaMemcpyAsync(f->h_cnt, f->d_cnt, 4 * sizeof(unsigned int), cudaMemcpyDeviceToHost, f->st);
  cudaError_t e = cudaStreamSynchronize(f->st);
  if (e != cudaSuccess) return vils::fail_cuda(e, "vils_good_features");
  cudaEventElapsedTime(&f->last_ms, f->e0, f->e1);
  const unsigned int total = std::min(f->h_cnt[0], f->cand_cap);
  if (total == 0) return VILS_OK;
  // the greedy minimum-distance pick (featureselect.cpp: grid of cell_size = round(minDistance), 3x3 cell neighbourhood) is sequential by
  // construction; it runs over the sorted candidates on the host, where the reference has it.  Returns true when it ran out of candidates.
  auto pick = [&](unsigned int count, int& n) -> bool {
    n = 0;
    if (min_distance >= 1) {
      const int cell = cv_round(min_distance), gw = (cols + cell - 1) / cell, gh = (rows + cell - 1) / cell;
      std::vector<std::vector<std::pair<int, int>>> grid((size_t)gw * gh);
      const double md2 = min_distance * min_distance;
      for (unsigned int i = 0; i < count; i++) {
        const int ofs = (int)(f->h_cand[i] & 0xffffffffu), y = ofs / cols, x = ofs % cols;
        const int xc = x / cell, yc = y / cell;
        const int x1 = std::max(0, xc - 1), y1 = std::max(0, yc - 1), x2 = std::min(gw - 1, xc + 1), y2 = std::min(gh - 1, yc + 1);
        bool good = true;
        for (int yy = y1; yy <= y2 && good; yy++)
          for (int xx = x1; xx <= x2 && good; xx++)
            for (const auto& p : grid[(size_t)yy * gw + xx]) {
              const float dx = (float)(x - p.first), dy = (float)(y - p.second);
              if (dx * dx + dy * dy < md2) { good = false; break; }
            }
        if (!good) continue;
        grid[(size_t)yc * gw + xc].push_back({x, y});
        xy_out[2 * n] = (float)x; xy_out[2 * n + 1] = (float)y; n++;
        if (max_corners > 0 && n == max_corners) return false;
      }
    } else {
      for (unsigned int i = 0; i < count; i++) {
        const int ofs = (int)(f->h_cand[i] & 0xffffffffu);
        xy_out[2 * n] = (float)(ofs % cols); xy_out[2 * n + 1] = (float)(ofs / cols); n++;
        if (max_corners > 0 && n == max_corners) return false;
      }
    }
    return true;
  };
  int n = 0;
  const unsigned int ntop = f->h_cnt[2];
  bool exhausted = true;
  if (ntop > 0) {
    e = cudaMemcpy(f->h_cand, f->d_top, sizeof(unsigned long long) * ntop, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return vils::fail_cuda(e, "vils_good_features copy");
    exhausted = pick(ntop, n);
  }
  if (exhausted && ntop < total) {   // rare: the strongest candidates were not enough (or did not fit shared memory): sort everything
    sort_desc_kernel<<<1, 1024, 0, f->st>>>(f->d_cand, f->d_cnt, f->cand_cap);
    e = cudaMemcpyAsync(f->h_cand, f->d_cand, sizeof(unsigned long long) * total, cudaMemcpyDeviceToHost, f->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(f->st);
    if (e != cudaSuccess) return vils::fail_cuda(e, "vils_good_features full sort");
    pick(total, n);
  }
  *n_out = n;
  return VILS_OK;
}

// Device-side min-eigenvalue map of the last vils_good_features call (tests / diagnostics).
int vils_frontend_get_eig(vils_frontend* f, float* eig) {
  if (!f || !eig) return vils::fail(VILS_ERR_BAD_ARG, "vils_frontend_get_eig");
  cudaSetDevice(f->device);
  cudaError_t e = cudaMemcpy(eig, f->d_eig, sizeof(float) * (size_t)f->rows * f->cols, cudaMemcpyDeviceToHost);
  return e == cudaSuccess ? VILS_OK : vils::fail_cuda(e, "vils_frontend_get_eig");
}

// PinholeCamera::liftProjective for n pixel positions: rays[3n] = (mx_u, my_u, 1).  cam = {fx, fy, cx, cy, k1, k2, p1, p2}.
int vils_lift_projective(vils_frontend* f, const double cam[8], const float* uv, int32_t n, double* rays) {
  if (!f || !cam || n < 0 || n > f->max_pts || (n && (!uv || !rays))) return vils::fail(VILS_ERR_BAD_ARG, "vils_lift_projective: bad argument");
  if (n == 0) return VILS_OK;
  cudaSetDevice(f->device);
  cudaMemcpyAsync(f->d_uv, uv, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, f->st);
  const int nodist = cam[4] == 0.0 && cam[5] == 0.0 && cam[6] == 0.0 && cam[7] == 0.0;   // m_noDistortion (PinholeCamera.cc:296-306)
  lift_kernel<<<(n + 127) / 128, 128, 0, f->st>>>(f->d_uv, n, cam[0], cam[1], cam[2], cam[3], cam[4], cam[5], cam[6], cam[7], nodist, f->d_ray);
  cudaMemcpyAsync(rays, f->d_ray, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, f->st);
  cudaError_t e = cudaStreamSynchronize(f->st);
  return e == cudaSuccess ? VILS_OK : vils::fail_cuda(e, "vils_lift_projective");
}

// cv::findFundamentalMat(pts1, pts2, cv::FM_RANSAC, threshold, 0.99, status) as used by rejectWithF (:191): status[i] = 1 for inliers.
// n < 8 leaves every point in (the reference only calls it with >= 8 points, :171).  F (may be NULL) receives the winning 3x3, row-major.
int vils_reject_with_f(vils_frontend* f, const float* pts1, const float* pts2, int32_t n, double threshold, uint8_t* status, double* F) {
  if (!f || n < 0 || n > f->max_pts || (n && (!pts1 || !pts2 || !status)) || threshold <= 0) return vils::fail(VILS_ERR_BAD_ARG, "vils_reject_with_f: bad argument");
  if (n < 8) { for (int i = 0; i < n; i++) status[i] = 1; return VILS_OK; }
  cudaSetDevice(f->device);
  // scratch: reuse the candidate buffer (>= 512 KB): pts1 | pts2 | status | info | F
  float* d1 = reinterpret_cast<float*>(f->d_cand); float* d2 = d1 + 2 * n;
  uint8_t* dst = reinterpret_cast<uint8_t*>(d2 + 2 * n); int* dinfo = reinterpret_cast<int*>(dst + ((n + 15) & ~15)); double* dF = reinterpret_cast<double*>(dinfo + 4);
  cudaMemcpyAsync(d1, pts1, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, f->st);
  cudaMemcpyAsync(d2, pts2, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, f->st);
  cudaFuncSetAttribute(ransac_f_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);   // per device; a few microseconds
  if ((size_t)n * 32 > 64 * 1024) return vils::fail(VILS_ERR_CAPACITY, "vils_reject_with_f: too many points");
  cudaEventRecord(f->e0, f->st);
  ransac_f_kernel<<<1, RANSAC_H, (size_t)n * 32, f->st>>>(d1, d2, n, threshold * threshold, 12345u, dst, dinfo, dF);
  cudaEventRecord(f->e1, f->st);
  cudaMemcpyAsync(status, dst, n, cudaMemcpyDeviceToHost, f->st);
  double hF[9];
  cudaMemcpyAsync(hF, dF, sizeof(hF), cudaMemcpyDeviceToHost, f->st);
  cudaError_t e = cudaStreamSynchronize(f->st);
  if (e != cudaSuccess) return vils::fail_cuda(e, "vils_reject_with_f");
  cudaEventElapsedTime(&f->last_ms, f->e0, f->e1);
  if (F) memcpy(F, hF, sizeof(hF));
  return VILS_OK;
}

// FeatureManager::triangulate for n_feat features: feature f has observations pts[off[f]..off[f+1]) (x y z, un-normalised) in keyframes
// start[f], start[f]+1, ...;  Ps (n_kf x 3), Rs (n_kf x 9 row-major), tic(3), ric (9 row-major).  depth[f] = svd_V[2] / svd_V[3], or
// init_depth when negative.  The caller keeps the reference's gating (used_num >= 2, start_frame < WINDOW_SIZE - 2, depth not yet known).
int vils_triangulate(int32_t n_feat, const int32_t* start, const int32_t* off, const double* pts, int32_t n_kf, const double* Ps, const double* Rs,
                     const double tic[3], const double ric[9], double init_depth, double* depth, int32_t device) {
  if (n_feat < 0 || n_kf <= 0 || (n_feat && (!start || !off || !pts || !depth)) || !Ps || !Rs || !tic || !ric) return vils::fail(VILS_ERR_BAD_ARG, "vils_triangulate: bad argument");
  if (n_feat == 0) return VILS_OK;
  for (int f = 0; f < n_feat; f++) if (start[f] < 0 || off[f + 1] < off[f] || start[f] + (off[f + 1] - off[f]) > n_kf) return vils::fail(VILS_ERR_BAD_ARG, "vils_triangulate: observation outside the window");
  int st = vils::require_device(device); if (st) return st;
  const int n_obs = off[n_feat];
  const size_t bytes = sizeof(int32_t) * (2 * (size_t)n_feat + 1) + sizeof(double) * (3 * (size_t)n_obs + 12 * (size_t)n_kf + 12 + n_feat) + 64;
  uint8_t* d = nullptr;
  cudaError_t e = cudaMalloc(&d, bytes);
  if (e != cudaSuccess) return vils::fail_cuda(e, "vils_triangulate alloc");
  std::vector<double> hd(3 * (size_t)n_obs + 12 * (size_t)n_kf + 12);
  memcpy(hd.data(), pts, sizeof(double) * 3 * n_obs); memcpy(hd.data() + 3 * n_obs, Ps, sizeof(double) * 3 * n_kf);
  memcpy(hd.data() + 3 * n_obs + 3 * n_kf, Rs, sizeof(double) * 9 * n_kf);
  memcpy(hd.data() + 3 * n_obs + 12 * n_kf, tic, sizeof(double) * 3); memcpy(hd.data() + 3 * n_obs + 12 * n_kf + 3, ric, sizeof(double) * 9);
  double* dd = reinterpret_cast<double*>(d); double* ddepth = dd + hd.size();
  int32_t* di = reinterpret_cast<int32_t*>(ddepth + n_feat);
  cudaMemcpy(dd, hd.data(), sizeof(double) * hd.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(di, start, sizeof(int32_t) * n_feat, cudaMemcpyHostToDevice);
  cudaMemcpy(di + n_feat, off, sizeof(int32_t) * (n_feat + 1), cudaMemcpyHostToDevice);
  triangulate_kernel<<<(n_feat + 63) / 64, 64>>>(n_feat, di, di + n_feat, dd, n_kf, dd + 3 * n_obs, dd + 3 * n_obs + 3 * n_kf, dd + 3 * n_obs + 12 * n_kf, init_depth, ddepth);
  e = cudaMemcpy(depth, ddepth, sizeof(double) * n_feat, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return e == cudaSuccess ? VILS_OK : vils::fail_cuda(e, "vils_triangulate");
}

int vils_frontend_last_device_ms(vils_frontend* f, float* ms) { if (!f || !ms) return VILS_ERR_BAD_ARG; *ms = f->last_ms; return VILS_OK; }

}  // extern "C"
