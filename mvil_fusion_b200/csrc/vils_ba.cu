// vils_ba.cu — sliding-window BA part of libvils_b200.so: host packer + kernels + C-ABI (include/vils_cabi.h).
// Replaces the body of Estimator::optimization() between vector2double() and double2vector()
// (vils_estimator/src/estimator.cpp:1124-1419) and the factor Evaluate()s it drives through ceres.
// There is no CPU path: every entry point fails with VILS_ERR_NO_DEVICE when no sm_100 device is usable.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include <cooperative_groups.h>
#include <dlfcn.h>
#include <immintrin.h>
#include <nccl.h>

#include "ba_device.cuh"
#include "common.h"
#include "ba_cluster.cuh"
#include "margin_device.cuh"

using namespace vb;

// =================================================================================================================
// kernels
// =================================================================================================================
// IMUFactor's sqrt_info = LLT(covariance.inverse()).matrixL().transpose() (factor/imu_factor.h:64) by ONE WARP.
// Same elimination sequence, pivot rule and per-element operation order as the scalar vf::imu_sqrt_info (partial-pivot LU
// inverse, then lower Cholesky), with the 15 rows / columns spread over lanes 0..14.  wk: 450 doubles of shared memory
// private to the warp.  W: 15x15 row-major upper triangular, NaN-filled on breakdown.
__device__ void warp_imu_sqrt_info(const double* cov, double* W, double* wk) {
  const int lane = threadIdx.x & 31;
  double* a = wk; double* inv = wk + 225;
  for (int e = lane; e < 225; e += 32) a[e] = cov[(e % 15) * 15 + e / 15];   // a[i][j] = cov(i, j), column-major input
  __syncwarp();
  int mypiv = lane;            // lane i < 15: piv[i]
  bool ok = true;
  for (int k = 0; k < 15; k++) {
    // pivot: first row >= k with the largest |a[i][k]|
    double best = (lane >= k && lane < 15) ? fabs(a[lane * 15 + k]) : -1.0; int p = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o); const int op = __shfl_xor_sync(0xffffffffu, p, o);
      if (ob > best || (ob == best && op < p)) { best = ob; p = op; }
    }
    if (best == 0.0 || !(best == best)) { ok = false; break; }
    if (p != k) {
      if (lane < 15) { const double t = a[k * 15 + lane]; a[k * 15 + lane] = a[p * 15 + lane]; a[p * 15 + lane] = t; }
      const int pk = __shfl_sync(0xffffffffu, mypiv, k), pp = __shfl_sync(0xffffffffu, mypiv, p);
      if (lane == k) mypiv = pp; else if (lane == p) mypiv = pk;
    }
    __syncwarp();
    if (lane > k && lane < 15) {
      double* row = a + lane * 15; const double* rk = a + k * 15;
      row[k] /= rk[k]; const double l = row[k];
      for (int j = k + 1; j < 15; j++) row[j] -= l * rk[j];
    }
    __syncwarp();
  }
  if (ok) {
    // lane c < 15 owns column c of the inverse: forward (unit lower) then backward (upper) substitution
    int piv[15];
#pragma unroll
    for (int i = 0; i < 15; i++) piv[i] = __shfl_sync(0xffffffffu, mypiv, i);
    if (lane < 15) {
      double x[15];
#pragma unroll
      for (int i = 0; i < 15; i++) x[i] = (piv[i] == lane) ? 1.0 : 0.0;
#pragma unroll
      for (int i = 0; i < 15; i++) { double sacc = x[i];
#pragma unroll
        for (int k = 0; k < 15; k++) if (k < i) sacc -= a[i * 15 + k] * x[k];
        x[i] = sacc; }
#pragma unroll
      for (int i = 14; i >= 0; i--) { double sacc = x[i];
#pragma unroll
        for (int k = 0; k < 15; k++) if (k > i) sacc -= a[i * 15 + k] * x[k];
        x[i] = sacc / a[i * 15 + i]; }
#pragma unroll
      for (int i = 0; i < 15; i++) inv[i * 15 + lane] = x[i];
    }
    __syncwarp();
    // lower Cholesky of inv (reads the lower triangle like Eigen's LLT): lane i owns row i
    for (int j = 0; j < 15; j++) {
      double d = inv[j * 15 + j];
      for (int k = 0; k < j; k++) d -= inv[j * 15 + k] * inv[j * 15 + k];
      if (!(d > 0.0)) { ok = false; break; }            // warp-uniform: every lane evaluates the same d
      d = sqrt(d);
      __syncwarp();
      if (lane == j) inv[j * 15 + j] = d;
      if (lane > j && lane < 15) { double sacc = inv[lane * 15 + j]; for (int k = 0; k < j; k++) sacc -= inv[lane * 15 + k] * inv[j * 15 + k]; inv[lane * 15 + j] = sacc / d; }
      __syncwarp();
    }
  }
  __syncwarp();
  for (int e = lane; e < 225; e += 32) { const int i = e / 15, j = e % 15; W[e] = ok ? ((j >= i) ? inv[j * 15 + i] : 0.0) : nan(""); }
  __syncwarp();
}

// Once per upload (or at the head of a solve when SolveParams::do_prep is set): IMU sqrt_info (imu_factor.h:64 recomputes
// it in every Evaluate; it only depends on the pre-integration), prior A = J_lin^T J_lin and b0 = J_lin^T r_lin, and the
// zero pattern of E.  work: 450 doubles of shared memory per warp of the block.  Ends with a block barrier.
__device__ void prep_window(const SolveParams& P, const Win& W, double* scr, double* work) {
  const int warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const double* pre = W.d(OFF_IMU);
  for (int k = warp; k < W.h->n_imu; k += nwarp) warp_imu_sqrt_info(pre + (size_t)k * 467 + 242, scr + P.sl.w_imu + (size_t)k * 225, work + warp * 450);
  const int n = W.h->prior_n;
  const double* J = W.d(OFF_PRIOR_J); const double* r = W.d(OFF_PRIOR_R);
  for (int e = threadIdx.x; e < n * n + n; e += blockDim.x) {
    if (e < n * n) {
      const int a = e / n, b = e % n; double s = 0;
      for (int i = 0; i < n; i++) s = fma(J[(size_t)a * n + i], J[(size_t)b * n + i], s);
      scr[P.sl.priorA + e] = s;
    } else {
      const int a = e - n * n; double s = 0;
      for (int i = 0; i < n; i++) s = fma(J[(size_t)a * n + i], r[i], s);
      scr[P.sl.priorb0 + a] = s;
    }
  }
  for (int64_t e = threadIdx.x; e < (int64_t)W.h->n_lm * P.sl.Dv_pad; e += blockDim.x) scr[P.sl.E + e] = 0.0;
  for (int e = threadIdx.x; e < PAIR_LD * PAIR_LD; e += blockDim.x) scr[P.sl.pairpart + (int64_t)(P.Ncap * (P.Ncap - 1) / 2) * PAIR_LD * PAIR_LD + e] = 0.0;
  __syncthreads();
}

__global__ void __launch_bounds__(256) prep_kernel(SolveParams P) {
  __shared__ double work[8 * 450];
  const int slot = P.slot0 + blockIdx.x;
  const Win W = decode(P, slot);
  prep_window(P, W, P.scratch + (size_t)slot * P.sl.total, work);
}

__device__ bool cam_dim_fixed(const SolveParams& P, const Win& W, int d) {
  if (d >= W.D) return true;                      // tile padding
  if (d >= 15 * W.N + 6) return !P.cfg.use_td;
  if (d >= 15 * W.N) return !P.cfg.est_ex;
  return W.u(OFF_FIXED)[W.M + d / 15] != 0;       // kf_fixed
}

// Shared-memory staging area of the Cholesky of a reduced system kept in global memory: the staging union (pair-pass rows / Schur chunk, dead
// by then), if (nb + 1) tiles fit in it.
__device__ __forceinline__ double* chol_stage(const Smem& L, double* sm, int nb) {
  const int have = (L.hv >= 0 ? L.hv : L.imu) - L.uni;
  return (nb + 1) * TSZ + 16 * TLD <= have ? sm + L.uni : nullptr;
}

// Optional phase profiling (VILS_PROF=1): thread 0 of block 0 accumulates SM cycles per phase.
#ifndef VILS_NO_IMU_INLINE
#define VILS_NO_IMU_INLINE 0
#endif
#define PROF_T0() long long prof_t = (P.prof && blockIdx.x == 0 && threadIdx.x == 0) ? clock64() : 0
#define PROF(i) do { if (P.prof && blockIdx.x == 0 && threadIdx.x == 0) { const long long n_ = prof_clock(); P.prof[i] += n_ - prof_t; prof_t = n_; } } while (0)

// Full linearisation at state x: H (tiles, lower), g, hd and the cost (broadcast to all threads).
__device__ double linearize(const SolveParams& P, const Win& W, const Smem& L, double* sm, double* scr, const double* x, double* H, double* Hv,
                            double mu, bool need_cost = true, int jac_mode = 0) {
  double* g = sm + L.g; double* hd = sm + L.hd;
  PROF_T0();
  // IMU factors ride on the idle warps of the pair pass when their staging fits in the (still unused) Hv region
  const bool imu_inline = !VILS_NO_IMU_INLINE && P.hv_in_smem && W.h->n_imu * IMU_SLOT2 <= W.Dvp * W.Dvp + ((P.Ncap + 1) / 2) * 466;
  double c = pair_pass(P, W, x, sm + L.uni, scr, reinterpret_cast<const int*>(sm + L.pid), need_cost, imu_inline ? sm + L.hv : nullptr, reinterpret_cast<const uint16_t*>(sm + L.tbl), sm + L.rot);
  __syncthreads();
  PROF(0);
  landmark_reduce(P, W, sm + L.cinv, sm + L.glam, scr, mu, jac_mode, sm + L.craw, sm + L.scl);
  for (int e = threadIdx.x; e < W.Dvp * W.Dv; e += blockDim.x) Hv[e] = 0.0;
  __syncthreads();
  PROF(1);
  schur_syrk(P, W, sm + L.cinv, sm + L.glam, Hv, sm + L.gv, sm + L.uni, scr, L.hv >= 0 ? L.hv - L.uni : L.imu - L.uni);
  __syncthreads();
  PROF(2);
  for (int e = threadIdx.x; e < tri(W.nb) * TSZ; e += blockDim.x) H[e] = 0.0;
  for (int e = threadIdx.x; e < W.nb * TB; e += blockDim.x) { g[e] = 0.0; hd[e] = 0.0; }
  __syncthreads();
  gather_visual(P, W, Hv, sm + L.gv, H, g, hd, scr, reinterpret_cast<const int*>(sm + L.pid), P.Ncap * (P.Ncap - 1) / 2);
  __syncthreads();
  PROF(3);
  if (imu_inline) imu_add(P, W, H, g, hd, scr);
  else c += imu_pass(P, W, x, H, g, hd, P.hv_in_smem ? sm + L.hv : sm + L.imu, scr, true, P.hv_in_smem != 0);
  PROF(4);
  c += lidar_pass(P, W, x, H, g, hd, true);
  __syncthreads();
  PROF(5);
  c += icp_lps_pass(P, W, x, H, g, hd, sm + L.imu, true);
  c += prior_pass(P, W, x, H, g, hd, sm + L.dx, scr, true);
  __syncthreads();
  c = block_sum(c, sm + L.red);
  PROF(6);
  return c;
}

__device__ double cost_only(const SolveParams& P, const Win& W, const Smem& L, double* sm, double* scr, const double* x) {
  double c = 0;
  for (int f = threadIdx.x; f < W.h->n_proj; f += blockDim.x) c += proj_cost(P, W, x, f);
  c += imu_pass(P, W, x, nullptr, nullptr, nullptr, sm + L.imu, scr, false, false);
  c += lidar_pass(P, W, x, nullptr, nullptr, nullptr, false);
  __syncthreads();
  c += icp_lps_pass(P, W, x, nullptr, nullptr, nullptr, sm + L.imu, false);
  c += prior_pass(P, W, x, nullptr, nullptr, nullptr, sm + L.g, scr, false);   // L.g is dead after the back-substitution; L.dx still holds the GN point (dogleg reuse)
  __syncthreads();
  return block_sum(c, sm + L.red);
}

// Constant blocks -> identity rows/cols (problem.SetParameterBlockConstant, estimator.cpp:1154-1166,1217-1221,1368-1370),
// Levenberg/Jacobi damping d2 = mu clamp(diag), b = -g.  fx: per-dimension constant mask (tile padding included),
// nfix: number of constant dimensions below D (0 in the common case -> no O(D^2) sweep).
// scc / ddc (dogleg): Jacobi scale per dimension (fixed at x0) and the resulting trust-region diagonal dd = clamp(hd s^2) / s^2, kept for the
// dogleg algebra (0 for constant dimensions).
__device__ void damp_and_fix(const Win& W, double* H, double* g, const double* hd, const int* fx, int nfix, double mu, const double* scc = nullptr, double* ddc = nullptr) {
  const int Dp = W.nb * TB;
  for (int i = threadIdx.x; i < Dp; i += blockDim.x) {
    if (fx[i]) { H[tidx(i, i)] = 1.0 + mu; g[i] = 0.0; if (ddc) ddc[i] = 0.0; }
    else {
      double dd = fmin(fmax(hd[i], 1e-6), 1e32);
      if (scc) { const double sj = scc[i]; dd = fmin(fmax(hd[i] * sj * sj, 1e-6), 1e32) / (sj * sj); }
      if (ddc) ddc[i] = dd;
      H[tidx(i, i)] += mu * dd; g[i] = -g[i];
    }
  }
  if (nfix > 0)
    for (int e = threadIdx.x; e < W.D * W.D; e += blockDim.x) {
      const int i = e / W.D, j = e % W.D;
      if (i > j && (fx[i] || fx[j])) H[tidx(i, j)] = 0.0;
    }
}

// dl_f = -cinv_f (g_l + E_f^T dx_v) ; lam += dl ; poses/speed-bias/ex/td (+)= dx.  xout != xin.
__device__ void apply_step(const SolveParams& P, const Win& W, const Smem& L, double* sm, const double* scr, const double* xin, double* xout) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double* dx = sm + L.dx; const double* cinv = sm + L.cinv; const double* glam = sm + L.glam;
  const double* E = scr + P.sl.E; const int32_t* lm_feat = W.lm_feat();
  const int X = 16 * W.N + 8 + W.M;
  for (int k = threadIdx.x; k < X; k += blockDim.x) xout[k] = xin[k];
  __syncthreads();
  // one warp per landmark, two landmarks per trip: their E rows come from L2 and are in flight together
  const int nlm = W.h->n_lm;
  for (int rnk = warp; rnk < nlm; rnk += 2 * SOLVE_WARPS) {
    const int rnk2 = rnk + SOLVE_WARPS; const bool two = rnk2 < nlm;
    const double* e1 = E + (size_t)rnk * W.Dvp; const double* e2 = E + (size_t)(two ? rnk2 : rnk) * W.Dvp;
    double s = 0, s2 = 0;
    for (int a = lane; a < W.Dv; a += 32) { const double v1 = e1[a], v2 = e2[a], d = dx[vis2cam(a, W.N)]; s = fma(v1, d, s); s2 = fma(v2, d, s2); }
    s = warp_sum(s); s2 = warp_sum(s2);
    if (lane == 0) { xout[XL(W.N) + lm_feat[rnk]] += -cinv[rnk] * (glam[rnk] + s); if (two) xout[XL(W.N) + lm_feat[rnk2]] += -cinv[rnk2] * (glam[rnk2] + s2); }
  }
  for (int k = threadIdx.x; k <= W.N; k += blockDim.x) {
    if (k < W.N) {
      vm::pose_plus(xout + XP(k), dx + 15 * k);
      for (int i = 0; i < 9; i++) xout[XS(W.N, k) + i] += dx[15 * k + 6 + i];
    } else {
      vm::pose_plus(xout + XE(W.N), dx + 15 * W.N);
      xout[XT(W.N)] += dx[15 * W.N + 6];
    }
  }
  __syncthreads();
}

__device__ __forceinline__ long long global_ns() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// dogleg: x (+)= step with step_c = stp (already blended), step_l = -a glam / dd_l + b (-cinv (glam + E_f^T y_c)).  Returns this thread's
// share of |step_l|^2 (tangent norm for the parameter-tolerance test).  xout != xin.
__device__ double apply_step_dogleg(const SolveParams& P, const Win& W, const Smem& L, double* sm, const double* scr, const double* xin, double* xout, double a, double b) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double* y = sm + L.dx; const double* stp = sm + L.stp; const double* cinv = sm + L.cinv; const double* glam = sm + L.glam;
  const double* craw = sm + L.craw; const double* scl = sm + L.scl;
  const double* E = scr + P.sl.E; const int32_t* lm_feat = W.lm_feat();
  const int X = 16 * W.N + 8 + W.M;
  double n2 = 0;
  for (int k = threadIdx.x; k < X; k += blockDim.x) xout[k] = xin[k];
  __syncthreads();
  // one warp per landmark, two landmarks per trip (their E rows, read from L2, in flight together; same order of the sums)
  const int nlm = W.h->n_lm;
  for (int rnk = warp; rnk < nlm; rnk += 2 * SOLVE_WARPS) {
    const int rnk2 = rnk + SOLVE_WARPS; const bool two = rnk2 < nlm;
    const double* e1 = E + (size_t)rnk * W.Dvp; const double* e2 = E + (size_t)(two ? rnk2 : rnk) * W.Dvp;
    double s = 0, s2 = 0;
    for (int q = lane; q < W.Dv; q += 32) { const double v1 = e1[q], v2 = e2[q], yy = y[vis2cam(q, W.N)]; s = fma(v1, yy, s); s2 = fma(v2, yy, s2); }
    s = warp_sum(s); s2 = warp_sum(s2);
    if (lane == 0) {
      if (cinv[rnk] != 0.0) {
        const double sj = scl[rnk], dd = fmin(fmax(craw[rnk] * sj * sj, 1e-6), 1e32) / (sj * sj);
        const double st = -a * glam[rnk] / dd + b * (-cinv[rnk] * (glam[rnk] + s));
        xout[XL(W.N) + lm_feat[rnk]] += st; n2 += st * st;
      }
      if (two && cinv[rnk2] != 0.0) {
        const double sj = scl[rnk2], dd = fmin(fmax(craw[rnk2] * sj * sj, 1e-6), 1e32) / (sj * sj);
        const double st = -a * glam[rnk2] / dd + b * (-cinv[rnk2] * (glam[rnk2] + s2));
        xout[XL(W.N) + lm_feat[rnk2]] += st; n2 += st * st;
      }
    }
  }
  for (int k = threadIdx.x; k <= W.N; k += blockDim.x) {
    if (k < W.N) {
      vm::pose_plus(xout + XP(k), stp + 15 * k);
      for (int i = 0; i < 9; i++) xout[XS(W.N, k) + i] += stp[15 * k + 6 + i];
    } else {
      vm::pose_plus(xout + XE(W.N), stp + 15 * W.N);
      xout[XT(W.N)] += stp[15 * W.N + 6];
    }
  }
  __syncthreads();
  return n2;
}

// One CTA = one window for the whole solve.  SMEM_H: the tile-packed H and the visual sub-system Hv live in shared memory
// (windows up to ~12 keyframes); otherwise they sit in the per-window scratch (L2).  TR: the trust-region modes (Levenberg-Marquardt,
// dogleg) are compiled into their own instantiation so that the fixed-count Gauss-Newton kernel keeps its register allocation.
// ---- cluster-assisted linearisation (solve_kernel<.., CL = true>): the trust-region loop runs on CTA 0 of a thread-block cluster exactly as it
// runs on a lone CTA; only the linearisation — factor pass, landmark reduction + partial Schur complements, gather — is dealt to all CTAs
// (the phases of the Gauss-Newton latency kernel).  Protocol: CTA 0 publishes the state and the parameters of the linearisation in the exchange
// area and meets the helpers at a cluster barrier; three more barriers separate the phases; the helpers then wait at the barrier of the next
// command (another linearisation, or stop) while CTA 0 adds the serial terms, factors, and tries steps.
struct ClPtrs { double *hvpart, *gvpart, *lidblk, *gsc, *hdsc, *costp, *xg, *cinvg, *glamg, *crawg, *sclg, *cmd; };
__device__ __forceinline__ ClPtrs cl_ptrs(const SolveParams& P, int slot) {
  double* cx = P.clbuf + (size_t)slot * P.cl.total;
  ClPtrs c; c.hvpart = cx + P.cl.hvpart; c.gvpart = cx + P.cl.gvpart; c.lidblk = cx + P.cl.lidblk; c.gsc = cx + P.cl.gsc; c.hdsc = cx + P.cl.hdsc;
  c.costp = cx + P.cl.costp; c.xg = cx + P.cl.xg; c.cinvg = cx + P.cl.cinvg; c.glamg = cx + P.cl.glamg; c.crawg = cx + P.cl.crawg; c.sclg = cx + P.cl.sclg;
  c.cmd = cx + P.cl.cmd;
  return c;
}

// the share of CTA r (every CTA of the cluster, CTA 0 included), from the command barrier to the barrier after the gather
template <bool SMEM_H>
__device__ void cluster_linearize_share(const SolveParams& P, const Win& W, const Smem& L, double* sm, double* scr, const ClPtrs& C, const double* x,
                                        double* H, double* Hdst, int* own, bool need_cost, double mu, int jac_mode, int r, int G) {
  const int Dp = W.nb * TB;
  const int* pid_s = reinterpret_cast<const int*>(sm + L.pid);
  const uint16_t* tb = reinterpret_cast<const uint16_t*>(sm + L.tbl);
  if (!SMEM_H) for (int e = r + G * threadIdx.x; e < tri(W.nb) * TSZ; e += G * blockDim.x) H[e] = 0.0;
  for (int e = r + G * threadIdx.x; e < Dp; e += G * blockDim.x) { C.gsc[e] = 0.0; C.hdsc[e] = 0.0; }
  double c = pair_pass_cluster(P, W, x, sm + L.uni, scr, need_cost, sm + L.imu, P.cl_imu_slots, tb, sm + L.rot, own, r, G);
  c += lidar_pass_cluster(P, W, x, C.lidblk, true, r, G);
  if (need_cost) { c = block_sum(c, sm + L.red); if (threadIdx.x == 0) C.costp[r] = c; }
  __syncthreads();
  cluster_sync_all();
  landmark_reduce_cluster(P, W, sm + L.cinv, sm + L.glam, scr, mu, r, G, jac_mode, sm + L.craw, sm + L.scl);
  __syncthreads();
  for (int rnk = r + G * threadIdx.x; rnk < W.h->n_lm; rnk += G * blockDim.x) {      // CTA 0 needs every landmark's reduction for the step algebra
    C.cinvg[rnk] = sm[L.cinv + rnk]; C.glamg[rnk] = sm[L.glam + rnk];
    if (jac_mode) { C.crawg[rnk] = sm[L.craw + rnk]; C.sclg[rnk] = sm[L.scl + rnk]; }
  }
  schur_syrk_cluster(P, W, sm + L.cinv, sm + L.glam, C.hvpart + (size_t)r * W.Dvp * W.Dvp, C.gvpart + (size_t)r * W.Dvp, sm + L.uni, scr, L.imu - L.uni, r, G);
  __syncthreads();
  if (SMEM_H && r == 0) for (int e = threadIdx.x; e < tri(W.nb) * TSZ; e += blockDim.x) H[e] = 0.0;
  __syncthreads();
  cluster_sync_all();
  gather_cluster(P, W, C.hvpart, C.gvpart, Hdst, C.gsc, C.hdsc, scr, pid_s, P.Ncap * (P.Ncap - 1) / 2, r, G);
  __syncthreads();
  cluster_sync_all();
}

// CTA 0: command + own share + the serial terms.  Leaves what linearize() leaves: H (lower tiles), g, hd, gv, cinv / glam (/ craw / scl) of every
// landmark in shared memory, E in the scratch; returns the cost when need_cost.
template <bool SMEM_H>
__device__ double linearize_cluster(const SolveParams& P, const Win& W, const Smem& L, double* sm, double* scr, const ClPtrs& C, const double* x,
                                    double* H, int* own, double mu, bool need_cost, int jac_mode, int G) {
  const int X = 16 * W.N + 8 + W.M, Dp = W.nb * TB;
  for (int k = threadIdx.x; k < X; k += blockDim.x) C.xg[k] = x[k];
  if (threadIdx.x == 0) { C.cmd[0] = 1.0; C.cmd[1] = mu; C.cmd[2] = need_cost ? 1.0 : 0.0; C.cmd[3] = (double)jac_mode; }
  __syncthreads();
  cluster_sync_all();                                    // command
  cluster_linearize_share<SMEM_H>(P, W, L, sm, scr, C, x, H, H, own, need_cost, mu, jac_mode, 0, G);
  for (int rnk = threadIdx.x; rnk < W.h->n_lm; rnk += blockDim.x) {
    sm[L.cinv + rnk] = C.cinvg[rnk]; sm[L.glam + rnk] = C.glamg[rnk];
    if (jac_mode) { sm[L.craw + rnk] = C.crawg[rnk]; sm[L.scl + rnk] = C.sclg[rnk]; }
  }
  for (int a = threadIdx.x; a < W.Dv; a += blockDim.x) { double v = 0; for (int q = 0; q < G; q++) v += C.gvpart[(size_t)q * W.Dvp + a]; sm[L.gv + a] = v; }
  for (int e = threadIdx.x; e < Dp; e += blockDim.x) { sm[L.g + e] = C.gsc[e]; sm[L.hd + e] = C.hdsc[e]; }
  __syncthreads();
  imu_add(P, W, H, sm + L.g, sm + L.hd, scr);
  lidar_add(W, H, sm + L.g, sm + L.hd, C.lidblk);
  __syncthreads();
  double c2 = icp_lps_pass(P, W, x, H, sm + L.g, sm + L.hd, sm + L.imu, true);
  c2 += prior_pass(P, W, x, H, sm + L.g, sm + L.hd, sm + L.dx, scr, true);
  __syncthreads();
  c2 = block_sum(c2, sm + L.red);
  if (need_cost) for (int q = 0; q < G; q++) c2 += C.costp[q];
  return c2;
}

// helpers (CTA r > 0): serve linearisations until CTA 0 says stop
template <bool SMEM_H>
__device__ void cluster_helper_loop(const SolveParams& P, const Win& W, const Smem& L, double* sm, double* scr, const ClPtrs& C, double* H, double* Hdst,
                                    int* own, int r, int G) {
  const int X = 16 * W.N + 8 + W.M;
  double* xs = sm + L.xs;
  for (;;) {
    cluster_sync_all();                                  // command
    if (C.cmd[0] != 1.0) return;
    const double mu = C.cmd[1]; const bool need_cost = C.cmd[2] != 0.0; const int jac_mode = (int)C.cmd[3];
    for (int k = threadIdx.x; k < X; k += blockDim.x) xs[k] = C.xg[k];
    __syncthreads();
    cluster_linearize_share<SMEM_H>(P, W, L, sm, scr, C, xs, H, Hdst, own, need_cost, mu, jac_mode, r, G);
    if (!SMEM_H && chol_stage(L, sm, W.nb)) {           // a reduced system in global memory is factored by the whole cluster, if CTA 0 gets that far
      cluster_sync_all();
      if (C.cmd[4] == 1.0) cholesky_tiles_cluster(H, nullptr, nullptr, nullptr, W.nb, nullptr, sm + L.uni, r, G);
    }
  }
}

// CL: launched as one thread-block cluster per window; the loop below runs on CTA 0, the other CTAs serve its linearisations.
template <bool SMEM_H, bool TR, bool CL = false>
__global__ void __launch_bounds__(SOLVE_THREADS, SOLVE_MINB) solve_kernel(SolveParams P) {
  extern __shared__ __align__(16) double sm[];
  __shared__ int chol_flag;
  __shared__ int time_flag;
  __shared__ int own[CL ? 2 + 32 * 31 / 2 : 1];
  const int G = CL ? (int)cluster_size() : 1, r = CL ? (int)cluster_rank() : 0;
  const int slot = P.slot0 + (CL ? (int)blockIdx.x / G : (int)blockIdx.x);
  Win W = decode(P, slot);
  const Smem L = CL ? smem_layout(P.Ncap, P.Mcap, SMEM_H ? 1 : 0, 0) : smem_layout(P.Ncap, P.Mcap, P.h_in_smem, P.hv_in_smem);
  stage_tables(W, L, sm);
  // Scratch (landmark partials, E, pair blocks, IMU products, for large windows H / Hv / tile inverses) lives in a slot owned by the SM, not
  // by the window: one CTA is resident per SM, so n_sm slots are rewritten over and over and stay in L2 instead of streaming
  // (windows x 0.6 MB) of write-backs to HBM.  Everything in it is rebuilt by this kernel (prep_window below).
  unsigned smid; asm("mov.u32 %0, %%smid;" : "=r"(smid));
  double* scr = (P.tscratch && !CL) ? P.tscratch + (size_t)smid * P.sl.total : P.scratch + (size_t)slot * P.sl.total;
  double* H = SMEM_H ? sm + L.uni : scr + P.sl.Hg;
  double* Hv = CL ? nullptr : (SMEM_H ? sm + L.hv : (P.hv_in_smem ? sm + L.hv : scr + P.sl.Hvg));
  ClPtrs C{}; if (CL) C = cl_ptrs(P, slot);
  double* xs = sm + L.xs; double* xc = sm + L.xc;
  double* linv = L.linv >= 0 ? sm + L.linv : scr + P.sl.linvg;   // diagonal-tile inverses: shared memory, or L2 scratch for large windows
  int* fx = reinterpret_cast<int*>(sm + L.fx);
  const int X = 16 * W.N + 8 + W.M;
  const double* x0 = W.d(OFF_X);
  const long long t_start = P.time_cap_ns > 0 ? global_ns() : 0;
  if (CL) { prep_window_cluster(P, W, scr, sm + L.uni, r, G); }
  else if (P.do_prep || P.tscratch) prep_window(P, W, scr, sm + L.uni);     // IMU sqrt_info, prior A / b0, zero pattern of E: into the scratch slot this CTA uses
  for (int k = threadIdx.x; k < X; k += blockDim.x) xs[k] = x0[k];
  { int* pid_s = reinterpret_cast<int*>(sm + L.pid); const int32_t* pid_g = W.i(OFF_PAIR_ID); for (int k = threadIdx.x; k < W.N * W.N; k += blockDim.x) pid_s[k] = pid_g[k]; }
  { uint16_t* tb = reinterpret_cast<uint16_t*>(sm + L.tbl); for (int k = threadIdx.x; k < 450; k += blockDim.x) tb[k] = g_imu_tbl[k]; }
  double nf = 0;
  for (int d = threadIdx.x; d < W.nb * TB; d += blockDim.x) { const bool f = cam_dim_fixed(P, W, d); fx[d] = f; if (f && d < W.D) nf += 1.0; }
  if (threadIdx.x == 0) { chol_flag = 0; time_flag = 0; }
  const int nfix = (int)block_sum(nf, sm + L.red);
  __syncthreads();
  if (CL) {
    cluster_sync_all();                                  // the prep of every CTA is in the scratch
    if (r != 0) {
      double* Hdst = SMEM_H ? cooperative_groups::this_cluster().map_shared_rank(sm + L.uni, 0) : H;
      cluster_helper_loop<SMEM_H>(P, W, L, sm, scr, C, H, Hdst, own, r, G);
      return;
    }
  }

  const bool lm = TR && P.mode == VILS_MODE_LM;
  const bool dl = TR && P.mode == VILS_MODE_DOGLEG;
  int status = VILS_OK, iters = 0, accepted = 0, trials = 0, capped = 0;
  double cost0 = 0, cost = 0, radius = P.lm_radius, decrease = 2.0;
  // dogleg state (ceres DoglegStrategy): mu of the regularised Gauss-Newton solve, reuse of the last GN point / gradient after a rejected step
  double mu_dl = 1e-8, dl_gg = 0, dl_pHp = 0, dl_yy = 0, dl_gy = 0;
  bool reuse = false, scale_set = false; int invalid = 0;

  for (int it = 0;; it++) {
    if (P.time_cap_ns > 0 && it > 0) {   // Solver::Options::max_solver_time_in_seconds: checked after every iteration, like ceres
      if (threadIdx.x == 0) time_flag = (global_ns() - t_start >= P.time_cap_ns) ? 1 : 0;
      __syncthreads();
      if (time_flag) { capped = 1; break; }
    }
    // LM follows ceres TrustRegionMinimizer + LevenbergMarquardtStrategy: (H + diag(clamp(H_ii))/radius) d = -g ; GN uses P.mu.
    const double mu = P.lin_out ? 0.0 : (lm ? 1.0 / radius : (dl ? mu_dl : P.mu));
    bool ok = true;
    double* bsave = sm + L.imu;   // LM: b = -g_r is overwritten by the fused forward solve; ((Ncap+1)/2)*466 doubles >= nb*16
    const bool last = !TR && (iters + 1 >= P.max_iters);
    if (!(dl && reuse)) {
      const double c_lin = CL ? linearize_cluster<SMEM_H>(P, W, L, sm, scr, C, xs, H, own, mu, it == 0, dl ? (scale_set ? 2 : 1) : 0, G)
                              : linearize(P, W, L, sm, scr, xs, H, Hv, mu, it == 0, dl ? (scale_set ? 2 : 1) : 0);   // later costs come from cost_only()
      if (it == 0) { cost0 = c_lin; cost = c_lin; if (TR) iters = 1; }
      const bool coop_chol = CL && !SMEM_H && chol_stage(L, sm, W.nb) != nullptr;   // the helpers wait for the verdict: factor with me, or not
      auto chol_cmd = [&](double v) { if (threadIdx.x == 0) C.cmd[4] = v; __syncthreads(); cluster_sync_all(); };
      if (!isfinite(c_lin)) { if (coop_chol) chol_cmd(0.0); status = VILS_ERR_NOT_FINITE; break; }
      if (P.lin_out) {   // vils_ba_linearize: one linearisation, constant blocks applied, no damping
        damp_and_fix(W, H, sm + L.g, sm + L.hd, fx, nfix, 0.0);
        __syncthreads();
        const int D = W.D;
        for (int e = threadIdx.x; e < D * D; e += blockDim.x) { const int i = e / D, j = e % D; P.lin_out[e] = H[tidx(max(i, j), min(i, j))]; }
        for (int i = threadIdx.x; i < D; i += blockDim.x) P.lin_out[(size_t)D * D + i] = -sm[L.g + i];
        if (threadIdx.x == 0) P.lin_out[(size_t)D * D + D] = c_lin;
        return;
      }
      if (P.max_iters <= 0) { if (coop_chol) chol_cmd(0.0); break; }
      PROF_T0();
      if (dl && !scale_set) {   // Jacobi scaling: s = 1 / (1 + |J column|), from the first linearisation only (ceres jacobi_scaling)
        for (int i = threadIdx.x; i < W.nb * TB; i += blockDim.x) sm[L.scc + i] = 1.0 / (1.0 + sqrt(fmax(sm[L.hd + i], 0.0)));
        scale_set = true;
        __syncthreads();
      }
      damp_and_fix(W, H, sm + L.g, sm + L.hd, fx, nfix, mu, dl ? sm + L.scc : nullptr, dl ? sm + L.ddc : nullptr);
      __syncthreads();
      if (lm) for (int i = threadIdx.x; i < W.nb * TB; i += blockDim.x) bsave[i] = sm[L.g + i];
      if (dl) {   // unreduced camera gradient g_c = g_r + E Cd^-1 g_l = -(b) - gv on the visual dimensions
        for (int i = threadIdx.x; i < W.nb * TB; i += blockDim.x) {
          double gci = 0.0;
          if (i < W.D && !fx[i]) {
            const int k = i / 15, o = i - 15 * k;
            const int va = i < 15 * W.N ? (o < 6 ? 6 * k + o : -1) : 6 * W.N + (i - 15 * W.N);
            gci = -sm[L.g + i] - (va >= 0 ? sm[L.gv + va] : 0.0);
          }
          sm[L.gc + i] = gci;
        }
      }
      if (threadIdx.x == 0) chol_flag = 0;
      __syncthreads();
      PROF(7);
      if (coop_chol) { chol_cmd(1.0); cholesky_tiles_cluster(H, sm + L.g, linv, sm + L.dx, W.nb, &chol_flag, sm + L.uni, 0, G); }
      else cholesky_tiles<SMEM_H>(H, sm + L.g, linv, sm + L.dx, W.nb, &chol_flag, P.prof, chol_stage(L, sm, W.nb));
      ok = chol_flag == 0;
      PROF(8);
      if (!ok && !TR) { status = VILS_ERR_CHOLESKY; break; }
      if (!ok && dl) {   // DoglegStrategy::ComputeGaussNewtonStep: raise mu and solve again (H was factored in place: re-linearise)
        mu_dl *= 10.0;
        if (mu_dl >= 1.0) { status = VILS_ERR_CHOLESKY; break; }
        continue;
      }
      if (ok) { backsub_tiles(H, sm + L.g, linv, sm + L.dx, W.nb, sm + L.red + 32); PROF(9); }
    }
    if (dl) {
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      const double* E = scr + P.sl.E;
      if (!reuse) {
        // scalars of the dogleg algebra over the FULL camera + landmark space (dogleg):
        //   gg = |g / sqrt(dd)|^2, yy = |sqrt(dd) y|^2, gy = g.y, pHp = v^T H v with v = g / dd and
        //   v^T H v = |L^T v_c|^2 - mu sum dd_c v_c^2 + sum_f [cinv_f t_f^2 + 2 t_f v_f + C_f v_f^2], t_f = E_f^T v_c
        double gg = 0, yy = 0, gy = 0, q = 0;
        for (int i = threadIdx.x; i < W.nb * TB; i += blockDim.x) {
          const double dd = sm[L.ddc + i], g_ = sm[L.gc + i], y_ = sm[L.dx + i];
          const double v = dd > 0.0 ? g_ / dd : 0.0;
          sm[L.stp + i] = v;
          if (dd > 0.0) { gg += g_ * v; yy += dd * y_ * y_; gy += g_ * y_; q -= mu * dd * v * v; }
        }
        __syncthreads();
        for (int j = threadIdx.x; j < W.D; j += blockDim.x) {   // (L^T v)_j
          double wj = 0;
          for (int i = j; i < W.D; i++) wj = fma(H[tidx(i, j)], sm[L.stp + i], wj);
          q += wj * wj;
        }
        const int nlm = W.h->n_lm;
        for (int rnk0 = warp; rnk0 < nlm; rnk0 += 2 * SOLVE_WARPS) {   // two landmarks per trip: their E rows (L2) in flight together
          const int rnk1 = rnk0 + SOLVE_WARPS; const bool two = rnk1 < nlm;
          const double* e0 = E + (size_t)rnk0 * W.Dvp; const double* e1 = E + (size_t)(two ? rnk1 : rnk0) * W.Dvp;
          double tv0 = 0, ty0 = 0, tv1 = 0, ty1 = 0;
          for (int e = lane; e < W.Dv; e += 32) {
            const double ev0 = e0[e], ev1 = e1[e]; const int ci = vis2cam(e, W.N); const double sv = sm[L.stp + ci], dv = sm[L.dx + ci];
            tv0 = fma(ev0, sv, tv0); ty0 = fma(ev0, dv, ty0); tv1 = fma(ev1, sv, tv1); ty1 = fma(ev1, dv, ty1);
          }
          tv0 = warp_sum(tv0); ty0 = warp_sum(ty0); tv1 = warp_sum(tv1); ty1 = warp_sum(ty1);
#pragma unroll
          for (int u = 0; u < 2; u++) {
            if (u && !two) break;
            const int rnk = u ? rnk1 : rnk0; const double tv = u ? tv1 : tv0, ty = u ? ty1 : ty0;
            const double ci = sm[L.cinv + rnk];
            if (lane == 0 && ci != 0.0) {
              const double gl = sm[L.glam + rnk], C = sm[L.craw + rnk], sj = sm[L.scl + rnk], dd = fmin(fmax(C * sj * sj, 1e-6), 1e32) / (sj * sj);
              const double yl = -ci * (gl + ty), vl = gl / dd;
              gg += gl * vl; yy += dd * yl * yl; gy += gl * yl;
              q += ci * tv * tv + 2.0 * tv * vl + C * vl * vl;
            }
          }
        }
        dl_gg = block_sum(gg, sm + L.red); dl_yy = block_sum(yy, sm + L.red); dl_gy = block_sum(gy, sm + L.red); dl_pHp = block_sum(q, sm + L.red);
      }
      // dogleg_strategy.cc ComputeTraditionalDoglegStep (every thread, identical arithmetic)
      const double alpha = dl_gg / dl_pHp, gn_norm = sqrt(dl_yy), g_norm = sqrt(dl_gg);
      double a, b, step_norm;
      if (gn_norm <= radius) { a = 0; b = 1; step_norm = gn_norm; }
      else if (g_norm * alpha >= radius) { a = radius / g_norm; b = 0; step_norm = radius; }
      else {
        const double b_dot_a = -alpha * dl_gy, a_sq = (alpha * g_norm) * (alpha * g_norm), bma = a_sq - 2 * b_dot_a + dl_yy, c = b_dot_a - a_sq;
        const double dsc = sqrt(c * c + bma * (radius * radius - a_sq));
        const double beta = (c <= 0) ? (dsc - c) / bma : (radius * radius - a_sq) / (dsc + c);
        a = alpha * (1.0 - beta); b = beta;
        step_norm = sqrt(fmax(0.0, a * a * dl_gg - 2 * a * b * dl_gy + b * b * dl_yy));
      }
      const double g_step = -a * dl_gg + b * dl_gy;
      const double sHs = a * a * dl_pHp + 2 * a * b * (dl_gg + mu * dl_gy) + b * b * (-dl_gy - mu * dl_yy);
      const double model = -g_step - 0.5 * sHs;
      trials++;
      if (!(model > 0)) {   // StepIsInvalid
        mu_dl *= 10.0; reuse = false;
        if (++invalid >= 5 || mu_dl >= 1.0 || trials >= P.max_iters) break;
        continue;
      }
      invalid = 0;
      double dn = 0;
      __syncthreads();
      for (int i = threadIdx.x; i < W.nb * TB; i += blockDim.x) {
        const double dd = sm[L.ddc + i];
        const double st = dd > 0.0 ? -a * sm[L.gc + i] / dd + b * sm[L.dx + i] : 0.0;
        sm[L.stp + i] = st; dn += st * st;
      }
      __syncthreads();
      dn += apply_step_dogleg(P, W, L, sm, scr, xs, xc, a, b);
      const double new_cost = cost_only(P, W, L, sm, scr, xc);
      const double quality = isfinite(new_cost) ? (cost - new_cost) / model : -1;
      if (quality > P.min_rel_dec) {
        double xn = 0;
        for (int k = threadIdx.x; k < X; k += blockDim.x) xn += xs[k] * xs[k];
        xn = block_sum(xn, sm + L.red); dn = block_sum(dn, sm + L.red);
        for (int k = threadIdx.x; k < X; k += blockDim.x) xs[k] = xc[k];
        __syncthreads();
        accepted++;
        if (quality < 0.25) radius *= 0.5;
        if (quality > 0.75) radius = fmax(radius, 3.0 * step_norm);
        mu_dl = fmax(1e-8, 2.0 * mu_dl / 10.0); reuse = false;
        const double change = cost - new_cost; cost = new_cost;
        if (fabs(change) / (cost + 1e-300) < P.f_tol) break;
        if (sqrt(dn) <= P.p_tol * (sqrt(xn) + P.p_tol)) break;
        if (trials >= P.max_iters) break;
        iters++;
      } else {
        radius *= 0.5; reuse = true;
        if (radius < 1e-32 || trials >= P.max_iters) break;
      }
      continue;
    }
    double rho = -1, new_cost = 0;
    if (ok) {
      double model = 0;
      if (lm) {
        // With (H + D2) d = -g over the full camera + landmark system: cost - model(d) = -1/2 g^T d + 1/2 d^T D2 d, where
        // the unreduced camera gradient is g_c = g_r + E Cd^-1 g_l  (Cd = damped landmark diagonal).
        double p2 = 0;
        for (int i = threadIdx.x; i < W.D; i += blockDim.x) {
          const double d = sm[L.dx + i];
          const double d2 = fx[i] ? mu : mu * fmin(fmax(sm[L.hd + i], 1e-6), 1e32);
          p2 += 0.5 * bsave[i] * d + 0.5 * d2 * d * d;
        }
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const double* E = scr + P.sl.E;
        for (int rnk = warp; rnk < W.h->n_lm; rnk += SOLVE_WARPS) {
          double sdot = 0;
          for (int a = lane; a < W.Dv; a += 32) sdot = fma(E[(size_t)rnk * W.Dvp + a], sm[L.dx + vis2cam(a, W.N)], sdot);
          sdot = warp_sum(sdot);
          if (lane == 0 && sm[L.cinv + rnk] != 0.0) {
            const double ci = sm[L.cinv + rnk], gl = sm[L.glam + rnk];
            const double dl_ = -ci * (gl + sdot);
            const double Cd = 1.0 / ci, C = Cd / (1.0 + mu);   // exact while C sits inside the clamp range [1e-6, 1e32]
            p2 += -0.5 * ci * gl * sdot - 0.5 * gl * dl_ + 0.5 * (Cd - C) * dl_ * dl_;
          }
        }
        model = block_sum(p2, sm + L.red);
      }
      PROF_T0();
      apply_step(P, W, L, sm, scr, xs, xc);
      PROF(10);
      if (lm || last) { new_cost = cost_only(P, W, L, sm, scr, xc); PROF(11); }
      if (lm) rho = (isfinite(new_cost) && model > 0) ? (cost - new_cost) / model : -1;
    }
    if (!TR) {   // Gauss-Newton: every step is taken
      for (int k = threadIdx.x; k < X; k += blockDim.x) xs[k] = xc[k];
      __syncthreads();
      iters++; accepted++;
      if (last) { cost = new_cost; if (!isfinite(cost)) status = VILS_ERR_NOT_FINITE; break; }
      continue;
    }
    trials++;
    if (ok && rho > P.min_rel_dec) {
      double xn = 0, dn = 0;
      for (int k = threadIdx.x; k < X; k += blockDim.x) { xn += xs[k] * xs[k]; const double dd = xc[k] - xs[k]; dn += dd * dd; }
      xn = block_sum(xn, sm + L.red); dn = block_sum(dn, sm + L.red);
      for (int k = threadIdx.x; k < X; k += blockDim.x) xs[k] = xc[k];
      __syncthreads();
      accepted++;
      const double t3 = 2.0 * rho - 1.0;
      radius = fmin(radius / fmax(1.0 / 3.0, 1.0 - t3 * t3 * t3), 1e16); decrease = 2.0;
      const double change = cost - new_cost; cost = new_cost;
      if (fabs(change) / (cost + 1e-300) < P.f_tol) break;
      if (sqrt(dn) <= P.p_tol * (sqrt(xn) + P.p_tol)) break;
      if (trials >= P.max_iters) break;
      iters++;
    } else {
      radius /= decrease; decrease *= 2.0;
      if (radius < 1e-32 || trials >= P.max_iters) break;
    }
  }
  if (!TR && capped) {   // a capped Gauss-Newton run reports the cost of the state it stopped at
    cost = cost_only(P, W, L, sm, scr, xs);
  }
  if (CL) {                                              // release the helpers
    if (threadIdx.x == 0) C.cmd[0] = 2.0;
    __syncthreads();
    cluster_sync_all();
  }
  double* xo = P.xout + (size_t)slot * P.xout_stride;
  for (int k = threadIdx.x; k < X; k += blockDim.x) xo[k] = xs[k];
  if (threadIdx.x == 0) {
    vils_summary s; s.status = status; s.iterations = iters; s.accepted = accepted; s.reserved = capped; s.cost_initial = cost0; s.cost_final = cost;
    P.summary[slot] = s;
  }
}

// ---- latency form: one window per thread-block CLUSTER (ba_cluster.cuh) ---------------------------------------------------------------------
// SMEM_H: the tile-packed H lives in the shared memory of CTA 0 and the other CTAs of the cluster write their share of the gather straight into it
// through distributed shared memory (mapa / st.shared::cluster); otherwise (large windows) H sits in the window's L2 scratch.
template <bool SMEM_H>
__global__ void __launch_bounds__(SOLVE_THREADS, 1) solve_cluster_kernel(SolveParams P, ClScratch C, double* clbuf, int imu_slots) {
  extern __shared__ __align__(16) double sm[];
  __shared__ int chol_flag;
  __shared__ int own[2 + 32 * 31 / 2];
  const int G = (int)cluster_size(), r = (int)cluster_rank();
  const int win = blockIdx.x / G, slot = P.slot0 + win;
  Win W = decode(P, slot);
  const Smem L = smem_layout(P.Ncap, P.Mcap, SMEM_H ? 1 : 0, 0);   // the partial Hv always goes through the exchange area (L2)
  stage_tables(W, L, sm);
  double* scr = P.scratch + (size_t)slot * P.sl.total;
  double* cx = clbuf + (size_t)slot * C.total;            // indexed by slot: chunks of one pipelined call run concurrently
  double* hvpart = cx + C.hvpart; double* gvpart = cx + C.gvpart; double* lidblk = cx + C.lidblk; double* gsc = cx + C.gsc; double* hdsc = cx + C.hdsc;
  double* dxg = cx + C.dxg; double* lamg = cx + C.lamg; double* costp = cx + C.costp; double* flagg = cx + C.flagg;
  double* H = SMEM_H ? sm + L.uni : scr + P.sl.Hg;        // CTA 0's copy is the one that is factored
  double* Hdst = H;                                        // where the gather writes: CTA 0's shared memory (DSMEM) or the scratch
  if (SMEM_H) Hdst = cooperative_groups::this_cluster().map_shared_rank(sm + L.uni, 0);
  double* linv = SMEM_H ? sm + L.linv : scr + P.sl.linvg;
  double* xs = sm + L.xs; double* xc = sm + L.xc;
  int* fx = reinterpret_cast<int*>(sm + L.fx);
  int* pid_s = reinterpret_cast<int*>(sm + L.pid);
  const uint16_t* tb = reinterpret_cast<const uint16_t*>(sm + L.tbl);
  const int X = 16 * W.N + 8 + W.M, Dp = W.nb * TB;
  const double* x0 = W.d(OFF_X);
  prep_window_cluster(P, W, scr, sm + L.uni, r, G);
  for (int k = threadIdx.x; k < X; k += blockDim.x) xs[k] = x0[k];
  { const int32_t* pid_g = W.i(OFF_PAIR_ID); for (int k = threadIdx.x; k < W.N * W.N; k += blockDim.x) pid_s[k] = pid_g[k]; }
  { uint16_t* tw = reinterpret_cast<uint16_t*>(sm + L.tbl); for (int k = threadIdx.x; k < 450; k += blockDim.x) tw[k] = g_imu_tbl[k]; }
  double nf = 0;
  for (int d = threadIdx.x; d < Dp; d += blockDim.x) { const bool f = cam_dim_fixed(P, W, d); fx[d] = f; if (f && d < W.D) nf += 1.0; }
  if (threadIdx.x == 0) chol_flag = 0;
  const int nfix = (int)block_sum(nf, sm + L.red);
  for (int k = r + G * threadIdx.x; k < W.M; k += G * blockDim.x) lamg[k] = x0[XL(W.N) + k];
  cluster_sync_all();

  int status = VILS_OK, iters = 0;
  double cost0 = 0, cost = 0;
  const int zblk = P.Ncap * (P.Ncap - 1) / 2;
  long long ct_ = (P.prof && blockIdx.x == 0 && threadIdx.x == 0) ? clock64() : 0;
#define CLPROF(i) do { if (P.prof && blockIdx.x == 0 && threadIdx.x == 0) { const long long n_ = prof_clock(); P.prof[i] += n_ - ct_; ct_ = n_; } } while (0)
  for (int it = 0; it < P.max_iters; it++) {
    CLPROF(11);
    // zero this CTA's share of H / g / hd (filled by the gather after the next two barriers)
    if (!SMEM_H) for (int e = r + G * threadIdx.x; e < tri(W.nb) * TSZ; e += G * blockDim.x) H[e] = 0.0;
    for (int e = r + G * threadIdx.x; e < Dp; e += G * blockDim.x) { gsc[e] = 0.0; hdsc[e] = 0.0; }
    // F: own pairs / IMU factors / LiDAR keyframes
    double c = pair_pass_cluster(P, W, xs, sm + L.uni, scr, it == 0, sm + L.imu, imu_slots, tb, sm + L.rot, own, r, G);
    c += lidar_pass_cluster(P, W, xs, lidblk, true, r, G);
    if (it == 0) { c = block_sum(c, sm + L.red); if (threadIdx.x == 0) costp[r] = c; }
    __syncthreads(); CLPROF(0);
    cluster_sync_all();
    CLPROF(1);
    // L: own landmarks -> partial Schur complement
    landmark_reduce_cluster(P, W, sm + L.cinv, sm + L.glam, scr, P.mu, r, G);
    __syncthreads();
    schur_syrk_cluster(P, W, sm + L.cinv, sm + L.glam, hvpart + (size_t)r * W.Dvp * W.Dvp, gvpart + (size_t)r * W.Dvp, sm + L.uni, scr, L.imu - L.uni, r, G);
    __syncthreads();
    if (SMEM_H && r == 0) for (int e = threadIdx.x; e < tri(W.nb) * TSZ; e += blockDim.x) H[e] = 0.0;   // the staging union is free now: H of CTA 0 starts from zero
    __syncthreads(); CLPROF(2);
    cluster_sync_all();
    CLPROF(3);
    // G
    gather_cluster(P, W, hvpart, gvpart, Hdst, gsc, hdsc, scr, pid_s, zblk, r, G);
    __syncthreads(); CLPROF(4);
    cluster_sync_all();
    CLPROF(5);
    // C: the serial chain on CTA 0
    const bool coop_chol = !SMEM_H && chol_stage(L, sm, W.nb) != nullptr;
    if (!SMEM_H) {     // reduced system in global memory: the prior's n x n block is added by the whole cluster
      prior_add_H_cluster(P, W, H, hdsc, scr, r, G);
      cluster_sync_all();
    }
    if (r == 0) {
      for (int e = threadIdx.x; e < Dp; e += blockDim.x) { sm[L.g + e] = gsc[e]; sm[L.hd + e] = hdsc[e]; }
      __syncthreads();
      imu_add(P, W, H, sm + L.g, sm + L.hd, scr);
      lidar_add(W, H, sm + L.g, sm + L.hd, lidblk);
      __syncthreads();
      double c2 = icp_lps_pass(P, W, xs, H, sm + L.g, sm + L.hd, sm + L.imu, true);
      c2 += prior_pass(P, W, xs, H, sm + L.g, sm + L.hd, sm + L.dx, scr, true, SMEM_H);
      __syncthreads();
      if (it == 0) { c2 = block_sum(c2, sm + L.red); for (int q = 0; q < G; q++) c2 += costp[q]; cost0 = c2; cost = c2; }
      damp_and_fix(W, H, sm + L.g, sm + L.hd, fx, nfix, P.mu);
      if (threadIdx.x == 0) chol_flag = 0;
      __syncthreads();
      CLPROF(6);
      if (!coop_chol) cholesky_tiles<SMEM_H>(H, sm + L.g, linv, sm + L.dx, W.nb, &chol_flag, nullptr, chol_stage(L, sm, W.nb));
    }
    if (coop_chol) {   // reduced system in global memory: the whole cluster factors it (the trailing tiles of every tile row dealt to all CTAs)
      cluster_sync_all();                            // CTA 0's additions to H are visible to the cluster
      cholesky_tiles_cluster(H, sm + L.g, linv, sm + L.dx, W.nb, &chol_flag, sm + L.uni, r, G);
    }
    if (r == 0) {
      CLPROF(7);
      int st = VILS_OK;
      if (chol_flag) st = VILS_ERR_CHOLESKY;
      else if (it == 0 && !isfinite(cost0)) st = VILS_ERR_NOT_FINITE;
      else backsub_tiles(H, sm + L.g, linv, sm + L.dx, W.nb, sm + L.red + 32);
      __syncthreads();
      for (int e = threadIdx.x; e < Dp; e += blockDim.x) dxg[e] = sm[L.dx + e];
      if (threadIdx.x == 0) flagg[0] = (double)st;
      CLPROF(8);
    }
    cluster_sync_all();
    CLPROF(9);
    // U
    const int st = (int)flagg[0];
    if (st != VILS_OK) { status = st; break; }
    for (int e = threadIdx.x; e < Dp; e += blockDim.x) sm[L.dx + e] = dxg[e];
    for (int k = threadIdx.x; k < X; k += blockDim.x) xc[k] = xs[k];
    __syncthreads();
    {
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      const double* E = scr + P.sl.E; const int32_t* lm_feat = W.lm_feat();
      int j = 0;
      for (int rnk = r; rnk < W.h->n_lm; rnk += G, j++) {
        if (j % SOLVE_WARPS != warp) continue;
        double s = 0;
        for (int a = lane; a < W.Dv; a += 32) s = fma(E[(size_t)rnk * W.Dvp + a], sm[L.dx + vis2cam(a, W.N)], s);
        s = warp_sum(s);
        if (lane == 0) { const int feat = lm_feat[rnk]; const double v = xc[XL(W.N) + feat] + -sm[L.cinv + rnk] * (sm[L.glam + rnk] + s); lamg[feat] = v; }
      }
      for (int k = threadIdx.x; k <= W.N; k += blockDim.x) {
        if (k < W.N) {
          vm::pose_plus(xc + XP(k), sm + L.dx + 15 * k);
          for (int i = 0; i < 9; i++) xc[XS(W.N, k) + i] += sm[L.dx + 15 * k + 6 + i];
        } else {
          vm::pose_plus(xc + XE(W.N), sm + L.dx + 15 * W.N);
          xc[XT(W.N)] += sm[L.dx + 15 * W.N + 6];
        }
      }
    }
    cluster_sync_all();
    for (int k = threadIdx.x; k < W.M; k += blockDim.x) xc[XL(W.N) + k] = lamg[k];
    __syncthreads();
    for (int k = threadIdx.x; k < X; k += blockDim.x) xs[k] = xc[k];
    __syncthreads();
    CLPROF(10);
    iters++;
  }
#undef CLPROF
  // final cost, split like the factor pass
  if (status == VILS_OK && P.max_iters > 0) {
    double c = 0;
    for (int f = r + G * threadIdx.x; f < W.h->n_proj; f += G * blockDim.x) c += proj_cost(P, W, xs, f);
    c += lidar_pass_cluster(P, W, xs, lidblk, false, r, G);
    if (r == 0) {
      c += imu_pass(P, W, xs, nullptr, nullptr, nullptr, sm + L.imu, scr, false, false);
      __syncthreads();
      c += icp_lps_pass(P, W, xs, nullptr, nullptr, nullptr, sm + L.imu, false);
      c += prior_pass(P, W, xs, nullptr, nullptr, nullptr, sm + L.g, scr, false);
      __syncthreads();
    }
    c = block_sum(c, sm + L.red);
    if (threadIdx.x == 0) costp[r] = c;
    cluster_sync_all();
    if (r == 0) { cost = 0; for (int q = 0; q < G; q++) cost += costp[q]; if (!isfinite(cost)) status = VILS_ERR_NOT_FINITE; }
  }
  if (r == 0) {
    double* xo = P.xout + (size_t)slot * P.xout_stride;
    for (int k = threadIdx.x; k < X; k += blockDim.x) xo[k] = xs[k];
    if (threadIdx.x == 0) {
      vils_summary s; s.status = status; s.iterations = iters; s.accepted = iters; s.reserved = 0; s.cost_initial = cost0; s.cost_final = cost;
      P.summary[slot] = s;
    }
  }
}

// ---- factor-sharded mode (one window over several GPUs) ------------------------------------------------
// shard_lin_kernel: linearise THIS rank's factors at the device-resident state, leave [H (D x D) | g | hd | cost] in `buf`
// for the caller's all-reduce.  shard_upd_kernel: reduced buffer -> damping, Cholesky, back-substitution (identical on every
// rank), then landmark back-substitution + Plus for the local landmarks.
template <bool SMEM_H>
__global__ void __launch_bounds__(SOLVE_THREADS, 1) shard_lin_kernel(SolveParams P, double* buf, int first, double mu) {
  extern __shared__ __align__(16) double sm[];
  const int slot = P.slot0;
  const Win W = decode(P, slot);
  const Smem L = smem_layout(P.Ncap, P.Mcap, P.h_in_smem, P.hv_in_smem);
  double* scr = P.scratch + (size_t)slot * P.sl.total;
  double* H = SMEM_H ? sm + L.uni : scr + P.sl.Hg;
  double* Hv = SMEM_H ? sm + L.hv : (P.hv_in_smem ? sm + L.hv : scr + P.sl.Hvg);
  double* xs = sm + L.xs;
  const int X = 16 * W.N + 8 + W.M, D = W.D;
  double* xo = P.xout + (size_t)slot * P.xout_stride;
  const double* x0 = first ? W.d(OFF_X) : xo;
  for (int k = threadIdx.x; k < X; k += blockDim.x) { xs[k] = x0[k]; if (first) xo[k] = x0[k]; }
  { int* pid_s = reinterpret_cast<int*>(sm + L.pid); const int32_t* pid_g = W.i(OFF_PAIR_ID); for (int k = threadIdx.x; k < W.N * W.N; k += blockDim.x) pid_s[k] = pid_g[k]; }
  { uint16_t* tb = reinterpret_cast<uint16_t*>(sm + L.tbl); for (int k = threadIdx.x; k < 450; k += blockDim.x) tb[k] = g_imu_tbl[k]; }
  __syncthreads();
  const double c = linearize(P, W, L, sm, scr, xs, H, Hv, mu);
  for (int e = threadIdx.x; e < D * D; e += blockDim.x) { const int i = e / D, j = e % D; buf[e] = H[tidx(max(i, j), min(i, j))]; }
  for (int i = threadIdx.x; i < D; i += blockDim.x) { buf[(size_t)D * D + i] = sm[L.g + i]; buf[(size_t)D * D + D + i] = sm[L.hd + i]; }
  if (threadIdx.x == 0) buf[(size_t)D * D + 2 * D] = c;
  for (int r = threadIdx.x; r < W.h->n_lm; r += blockDim.x) { scr[P.sl.lmsave + r] = sm[L.cinv + r]; scr[P.sl.lmsave + P.Mcap + r] = sm[L.glam + r]; }
}

template <bool SMEM_H>
__global__ void __launch_bounds__(SOLVE_THREADS, 1) shard_upd_kernel(SolveParams P, const double* buf, double mu) {
  extern __shared__ __align__(16) double sm[];
  __shared__ int chol_flag;
  const int slot = P.slot0;
  const Win W = decode(P, slot);
  const Smem L = smem_layout(P.Ncap, P.Mcap, P.h_in_smem, P.hv_in_smem);
  double* scr = P.scratch + (size_t)slot * P.sl.total;
  double* H = SMEM_H ? sm + L.uni : scr + P.sl.Hg;
  double* xs = sm + L.xs; double* xc = sm + L.xc;
  int* fx = reinterpret_cast<int*>(sm + L.fx);
  const int X = 16 * W.N + 8 + W.M, D = W.D;
  double* xo = P.xout + (size_t)slot * P.xout_stride;
  for (int k = threadIdx.x; k < X; k += blockDim.x) xs[k] = xo[k];
  double nf = 0;
  for (int d = threadIdx.x; d < W.nb * TB; d += blockDim.x) { const bool f = cam_dim_fixed(P, W, d); fx[d] = f; if (f && d < W.D) nf += 1.0; }
  if (threadIdx.x == 0) chol_flag = 0;
  const int nfix = (int)block_sum(nf, sm + L.red);
  for (int e = threadIdx.x; e < tri(W.nb) * TSZ; e += blockDim.x) H[e] = 0.0;
  for (int e = threadIdx.x; e < W.nb * TB; e += blockDim.x) { sm[L.g + e] = 0.0; sm[L.hd + e] = 0.0; }
  __syncthreads();
  for (int e = threadIdx.x; e < D * D; e += blockDim.x) { const int i = e / D, j = e % D; if (i >= j) H[tidx(i, j)] = buf[e]; }
  for (int i = threadIdx.x; i < D; i += blockDim.x) { sm[L.g + i] = buf[(size_t)D * D + i]; sm[L.hd + i] = buf[(size_t)D * D + D + i]; }
  for (int r = threadIdx.x; r < W.h->n_lm; r += blockDim.x) { sm[L.cinv + r] = scr[P.sl.lmsave + r]; sm[L.glam + r] = scr[P.sl.lmsave + P.Mcap + r]; }
  __syncthreads();
  damp_and_fix(W, H, sm + L.g, sm + L.hd, fx, nfix, mu);
  __syncthreads();
  double* linv = L.linv >= 0 ? sm + L.linv : scr + P.sl.linvg;
  cholesky_tiles<SMEM_H>(H, sm + L.g, linv, sm + L.dx, W.nb, &chol_flag, nullptr, chol_stage(L, sm, W.nb));
  int status = VILS_OK;
  if (chol_flag) status = VILS_ERR_CHOLESKY;
  else {
    backsub_tiles(H, sm + L.g, linv, sm + L.dx, W.nb, sm + L.red + 32);
    apply_step(P, W, L, sm, scr, xs, xc);
    for (int k = threadIdx.x; k < X; k += blockDim.x) xo[k] = xc[k];
  }
  if (threadIdx.x == 0) {
    vils_summary s = P.summary[slot];
    s.status = status; s.iterations += 1; s.accepted += (status == VILS_OK); s.cost_final = buf[(size_t)D * D + 2 * D];
    P.summary[slot] = s;
  }
}

// ---- materialised evaluation: one thread per factor, outputs in the caller's original factor order -----------------
struct EvalParams {
  SolveParams S;
  double* r_out; double* J_out; int64_t r_stride, J_stride;   // per-slot strides (doubles)
  int32_t apply_loss;
};

// Projection factors: PERSISTENT kernel, grid = SM count x MINB blocks of 128 threads, one thread per factor, work items
// (window, 128-factor chunk) strided over the grid.  The SoA factor tile of the NEXT item (14 constants + 3 indices per
// factor, the 2.5 KB window state and the landmark -> feature table) is staged into shared memory with cp.async while the
// current item is being evaluated, so the HBM latency of the dependent chain header -> {constants, indices, state} is hidden
// behind FP64 math.  Each thread writes its corrected Jacobian row (2x20 doubles, 320 B) into its own 336-byte shared-memory
// row with 16-byte stores (21 x 16 B row pitch: conflict-free) and hands it to the TMA engine as one bulk shared->global
// copy (cp.async.bulk); the wait for that copy is deferred to just before the row is rewritten one item later.
// Output order inside a family is the library's SORTED order; the single-slot host API un-permutes (vils_ba_evaluate),
// the batched device API documents it (vils_ba_evaluate_device).
constexpr int EV_T = 128, EV_PLD = 42, EV_ELD = 21, EV_LLD = 7;
#ifndef EVP_T
#define EVP_T 128    // projection kernel CTA size (96-thread CTAs at 4 per SM balance the 906-factor window better but measured 7 % slower)
#endif
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
struct ProjItem { const uint8_t* base; const WinHdr* h; int slot, chunk, np, N, X, n_lm; };
__device__ __forceinline__ ProjItem proj_item(const SolveParams& P, int item, int pb) {
  ProjItem it; it.slot = P.slot0 + item / pb; it.chunk = item % pb;
  it.base = P.blobs + (size_t)it.slot * P.blob_stride; it.h = reinterpret_cast<const WinHdr*>(it.base);
  it.np = it.h->n_proj; it.N = it.h->n_kf; it.X = 16 * it.N + 8 + it.h->n_feat; it.n_lm = it.h->n_lm;
  return it;
}
// per-thread slots of the factor tile: c[k] at cst[k * 128 + t] (doubles), idx[k] at ist[k * 128 + t] (ints); shared part: xs, feat
__device__ __forceinline__ void proj_prefetch(const ProjItem& it, double* cst, int32_t* ist, double* xs, int32_t* feat, int t) {
  const int f = it.chunk * EVP_T + t;
  if (it.chunk * EVP_T < it.np) {
    const double* x = reinterpret_cast<const double*>(it.base + it.h->off[OFF_X]);
    for (int k = t; k < it.X; k += EVP_T) cp_async8(xs + k, x + k);
    const int32_t* lf = reinterpret_cast<const int32_t*>(it.base + it.h->off[OFF_LM_FEAT]);
    for (int k = t; k < it.n_lm; k += EVP_T) cp_async4(feat + k, lf + k);
    if (f < it.np) {
      const double* c0 = reinterpret_cast<const double*>(it.base + it.h->off[OFF_PROJ]);
      const int32_t* ix = reinterpret_cast<const int32_t*>(it.base + it.h->off[OFF_PROJ_IDX]);
#pragma unroll
      for (int k = 0; k < 14; k++) cp_async8(cst + k * EVP_T + t, c0 + (size_t)k * it.np + f);
#pragma unroll
      for (int k = 0; k < 3; k++) cp_async4(ist + k * EVP_T + t, ix + (size_t)k * it.np + f);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int MINB>
__global__ void __launch_bounds__(EVP_T, MINB) eval_proj_kernel(EvalParams Q, int n_items, int pb, int xs_doubles, int feat_ints) {
  extern __shared__ __align__(16) double st[];
  const SolveParams& P = Q.S;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  // carve-up: rows | cst | xs[2] | ist | feat[2]
  double* rows = st; double* cst = rows + EVP_T * EV_PLD; double* xsb = cst + 14 * EVP_T;
  int32_t* ist = reinterpret_cast<int32_t*>(xsb + 2 * xs_doubles); int32_t* featb = ist + 3 * EVP_T;
  int item = blockIdx.x;
  if (item >= n_items) return;
  const int G = gridDim.x;
  ProjItem cur = proj_item(P, item, pb);
  proj_prefetch(cur, cst, ist, xsb, featb, t);
  ProjItem nx = cur;
  if (item + G < n_items) nx = proj_item(P, item + G, pb);
  double* row = rows + (size_t)t * EV_PLD;
  const double* wrows = rows + (size_t)(32 * warp) * EV_PLD;       // this warp's 32 rows
  for (int n = 0; item < n_items; n++, item += G) {
    // the header of the item AFTER the next one is fetched now and consumed one iteration later (by its proj_prefetch)
    ProjItem nn = nx;
    if (item + 2 * G < n_items) nn = proj_item(P, item + 2 * G, pb);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                            // tile of `cur` complete and visible; previous item fully consumed
    const double* xs = xsb + (n & 1) * xs_doubles; const int32_t* feat = featb + (n & 1) * feat_ints;
    const int fw = cur.chunk * EVP_T + 32 * warp, f = fw + lane;   // fw: first factor of this warp
    const bool act = f < cur.np;
    double c[14]; int i = 0, j = 0, rank = 0;
    if (act) {
#pragma unroll
      for (int k = 0; k < 14; k++) c[k] = cst[k * EVP_T + t];
      i = ist[t]; j = ist[EVP_T + t]; rank = ist[2 * EVP_T + t];
    }
    if (item + G < n_items) proj_prefetch(nx, cst, ist, xsb + ((n + 1) & 1) * xs_doubles, featb + ((n + 1) & 1) * feat_ints, t);   // own slots already copied to registers
    const int n_imu = cur.h->n_imu;
    if (act) {
      const int N = cur.N;
      double r[2];
      // J goes straight into the shared-memory row as it is produced (keeps the 40 values out of the register file)
      vf::proj_eval_rows(P.cfg, c, vm::q2R(vm::ldq(xs + XP(i) + 3)), vm::q2R(vm::ldq(xs + XP(j) + 3)), vm::q2R(vm::ldq(xs + XE(N) + 3)), vm::ld3(xs + XP(i)),
                         vm::ld3(xs + XP(j)), vm::ld3(xs + XE(N)), xs[XL(N) + feat[rank]], xs[XT(N)], r, row, Q.apply_loss ? P.cfg.cauchy_a : 0.0);
      double* R = Q.r_out + (size_t)cur.slot * Q.r_stride + 15 * n_imu + 2 * (size_t)f;
      R[0] = r[0]; R[1] = r[1];
    }
    __syncwarp();
    if (fw < cur.np) {   // the warp's rows leave as one contiguous segment (<= 10 KB) of 16-byte stores
      const int cnt = min(32, cur.np - fw);
      double2* Jo = reinterpret_cast<double2*>(Q.J_out + (size_t)cur.slot * Q.J_stride + (size_t)450 * n_imu + (size_t)40 * fw);   // 16-byte aligned: every term is even
#pragma unroll 4
      for (int e = lane; e < cnt * 20; e += 32) {
        const int fr = e / 20, c2 = e - fr * 20;
        Jo[e] = *reinterpret_cast<const double2*>(wrows + fr * EV_PLD + 2 * c2);
      }
    }
    __syncwarp();
    cur = nx; nx = nn;
  }
}

// IMU factors: one WARP per factor (factors packed across windows).  warp_imu_whitened: one coalesced pass stages the
// pre-integration record, sqrt_info and the four state blocks in shared memory, lane 0 does the quaternion algebra, all
// lanes assemble the 15 x 30 Jacobian through the table and whiten it; the rows then leave as coalesced stores.
// Latency-bound and tiny (9 per window, 7 % of the bytes): launched ahead of the streaming kernels.
#ifndef EVI_WARPS_
#define EVI_WARPS_ 2
#endif
#ifndef EVI_MINB
#define EVI_MINB 8
#endif
constexpr int EVI_WARPS = EVI_WARPS_;
__device__ void eval_prior_row(const EvalParams& Q, int slot, int row);
__global__ void __launch_bounds__(32 * EVI_WARPS, EVI_MINB) eval_imu_kernel(EvalParams Q, int nimu_max, int n, int imu_blocks, int prior_bpw) {
  if ((int)blockIdx.x >= imu_blocks) {   // CTAs past the IMU factors: prior residual rows, prior_bpw CTAs per window
    const int b = blockIdx.x - imu_blocks;
    eval_prior_row(Q, Q.S.slot0 + b / prior_bpw, (b % prior_bpw) * 32 * EVI_WARPS + threadIdx.x);
    return;
  }
  __shared__ double sJ[EVI_WARPS][IMU_SLOT];
  __shared__ double sx[EVI_WARPS][32];
  __shared__ uint16_t stbl[450];
  const SolveParams& P = Q.S;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint16_t treg[8];                                            // table loads fly together with the header loads below
#pragma unroll
  for (int q = 0; q < 8; q++) { const int e = threadIdx.x + 32 * EVI_WARPS * q; treg[q] = e < 450 ? g_imu_tbl[e] : (uint16_t)0; }
  const int fid = blockIdx.x * EVI_WARPS + warp;             // factors packed across windows: no half-empty blocks
  const bool act0 = fid < nimu_max * n;
  const int slot = P.slot0 + (act0 ? fid / nimu_max : 0), k = act0 ? fid % nimu_max : 0;
  const Win W = decode(P, slot);
  const bool act = act0 && k < W.h->n_imu;
  const int N = W.N;
  int i = 0;
  if (act) {   // the four state blocks of the factor, one coalesced fetch: pose_i 0..6 | pose_j 7..13 | sb_i 14..22 | sb_j 23..31
    i = W.i(OFF_IMU_KF)[k];
    const double* x = W.d(OFF_X);
    sx[warp][lane] = x[lane < 14 ? XP(i) + lane : XS(N, i) + (lane - 14)];
  }
#pragma unroll
  for (int q = 0; q < 8; q++) { const int e = threadIdx.x + 32 * EVI_WARPS * q; if (e < 450) stbl[e] = treg[q]; }
  __syncthreads();
  if (!act) return;
  double* R = Q.r_out + (size_t)slot * Q.r_stride; double* Jo = Q.J_out + (size_t)slot * Q.J_stride;
  const double* scr = P.scratch + (size_t)slot * P.sl.total;
  double* J = sJ[warp]; const double* r = J + 450; const double* xw = sx[warp];
  warp_imu_whitened(stbl, W.d(OFF_IMU) + (size_t)k * 467, scr + P.sl.w_imu + (size_t)k * 225, P.cfg.G, xw, xw + 14, xw + 7, xw + 23, J, true);
  if (lane < 15) R[15 * k + lane] = r[lane];
  for (int e = lane; e < 450; e += 32) Jo[(size_t)450 * k + e] = J[e];
}

// LiDAR plane and edge factors (keyframe-sorted order), one thread per factor, results staged per warp in shared memory and
// written as contiguous 16-byte stores.  EDGE = false: LidarPlaneNormFactor (1 + 1x6), 7 KB of staging per CTA so that the SM
// runs at full occupancy; EDGE = true: LidarEdgeFactor (3 + 3x6).
template <bool EDGE>
__global__ void __launch_bounds__(EV_T, EDGE ? 6 : 8) eval_lidar_kernel(EvalParams Q) {
  extern __shared__ __align__(16) double st[];
  constexpr int LD = EDGE ? EV_ELD : EV_LLD, NJ = EDGE ? 18 : 6;
  const SolveParams& P = Q.S;
  const int slot = P.slot0 + blockIdx.y;
  const Win W = decode(P, slot);
  const WinHdr* h = W.h;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int n = EDGE ? h->n_edge : h->n_plane, fw = blockIdx.x * EV_T + 32 * warp, f = fw + lane;   // fw: first factor of this warp
  if (fw >= n) return;
  const double* x = W.d(OFF_X);
  double* R = Q.r_out + (size_t)slot * Q.r_stride + 15 * h->n_imu + 2 * (size_t)h->n_proj + (EDGE ? h->n_plane : 0);
  double* Jo = Q.J_out + (size_t)slot * Q.J_stride + (size_t)450 * h->n_imu + (size_t)40 * h->n_proj + (EDGE ? (size_t)6 * h->n_plane : 0);
  double* rows = st + (size_t)(32 * warp) * LD;
  if (f < n) {
    const double* c = W.d(EDGE ? OFF_EDGE : OFF_PLANE); const int32_t* ix = W.i(EDGE ? OFF_EDGE_IDX : OFF_PLANE_IDX);
    double* o = rows + lane * LD;
    if (!EDGE) {
      double J[6];
      const double r = vf::plane_eval(x + XP(ix[f]), vm::mk(c[f], c[(size_t)n + f], c[(size_t)2 * n + f]),
                                      vm::mk(c[(size_t)3 * n + f], c[(size_t)4 * n + f], c[(size_t)5 * n + f]), c[(size_t)6 * n + f], J);
      double w = 1.0;
      if (Q.apply_loss) { double rho; vf::huber(P.cfg.huber_a, r * r, rho, w); }
      R[f] = r * w;
#pragma unroll
      for (int e = 0; e < 6; e++) o[e] = J[e] * w;
    } else {
      double r[3], J[18];
      vf::edge_eval(x + XP(ix[f]), vm::mk(c[f], c[(size_t)n + f], c[(size_t)2 * n + f]),
                    vm::mk(c[(size_t)3 * n + f], c[(size_t)4 * n + f], c[(size_t)5 * n + f]),
                    vm::mk(c[(size_t)6 * n + f], c[(size_t)7 * n + f], c[(size_t)8 * n + f]), r, J);
      double w = 1.0;
      if (Q.apply_loss) { double rho; vf::huber(P.cfg.huber_a, r[0] * r[0] + r[1] * r[1] + r[2] * r[2], rho, w); }
#pragma unroll
      for (int e = 0; e < 18; e++) o[e] = J[e] * w;
#pragma unroll
      for (int e = 0; e < 3; e++) o[18 + e] = r[e] * w;
    }
  }
  __syncwarp();
  const int cnt = min(32, n - fw);
  if (EDGE) for (int e = lane; e < cnt * 3; e += 32) R[(size_t)3 * fw + e] = rows[(e / 3) * LD + 18 + e % 3];
  double2* J2 = reinterpret_cast<double2*>(Jo + (size_t)NJ * fw);     // 16-byte aligned: every term of the offset is even
  for (int e = lane; e < cnt * (NJ / 2); e += 32) {
    const int fr = e / (NJ / 2), c2 = e - fr * (NJ / 2);
    const double* sp = rows + fr * LD + 2 * c2;
    J2[e] = make_double2(sp[0], sp[1]);
  }
}

// MarginalizationFactor residual rows (marginalization_factor.cpp:364-383): one thread per row.  Runs as extra CTAs of the IMU
// launch (both are tiny and latency-bound; a launch of its own would cost more than the work).
__device__ void eval_prior_row(const EvalParams& Q, int slot, int row) {
  const SolveParams& P = Q.S;
  const Win W = decode(P, slot);
  const WinHdr* h = W.h;
  const int n = h->prior_n, N = W.N;
  if (row >= n) return;
  const double* x = W.d(OFF_X);
  double* R = Q.r_out + (size_t)slot * Q.r_stride;
  const int rbase = 15 * h->n_imu + 2 * h->n_proj + h->n_plane + 3 * h->n_edge + 3 * h->n_icp + 3 * h->n_lps;
  const int32_t* blk = W.i(OFF_PRIOR_BLK);
  const double* x0 = W.d(OFF_PRIOR_X0); const double* Jl = W.d(OFF_PRIOR_J);
  double r = W.d(OFF_PRIOR_R)[row];
  for (int b = 0; b < h->prior_nblk; b++) {
    const int type = blk[4 * b], idx = blk[4 * b + 1], xo = blk[4 * b + 2], c0 = blk[4 * b + 3];
    const double* xb = type == VILS_BLK_POSE ? x + XP(idx) : type == VILS_BLK_SPEEDBIAS ? x + XS(N, idx) : type == VILS_BLK_EXPOSE ? x + XE(N) : x + XT(N);
    double dx[9]; int sz;
    if (type == VILS_BLK_POSE || type == VILS_BLK_EXPOSE) { vf::prior_dx_pose(xb, x0 + xo, dx); sz = 6; }
    else { sz = type == VILS_BLK_SPEEDBIAS ? 9 : 1; for (int k = 0; k < sz; k++) dx[k] = xb[k] - x0[xo + k]; }
    for (int k = 0; k < sz; k++) r = fma(Jl[(size_t)(c0 + k) * n + row], dx[k], r);
  }
  R[rbase + row] = r;
}

// ICP / LPS constraints (forward-mode duals, register hungry): their own tiny kernel so that they do not set the register
// budget of the IMU kernel.  One thread per constraint.
__global__ void eval_cons_kernel(EvalParams Q) {
  const SolveParams& P = Q.S;
  const int slot = P.slot0 + blockIdx.y;
  const Win W = decode(P, slot);
  const WinHdr* h = W.h;
  const double* x = W.d(OFF_X);
  double* R = Q.r_out + (size_t)slot * Q.r_stride; double* Jo = Q.J_out + (size_t)slot * Q.J_stride;
  int t = threadIdx.x;
  int rbase; int64_t jbase;
  rbase = 15 * h->n_imu + 2 * h->n_proj + h->n_plane;
  jbase = (int64_t)450 * h->n_imu + (int64_t)40 * h->n_proj + (int64_t)6 * h->n_plane;
  rbase += 3 * h->n_edge; jbase += (int64_t)18 * h->n_edge;
  if (t < h->n_icp) {
    const double* c = W.d(OFF_ICP) + 14 * t; double r[3], J[72];
    vf::icp_eval(c, x + XP((int)c[10]), x + XP((int)c[11]), x + XP((int)c[12]), x + XP((int)c[13]), r, J);
    double w = 1.0;
    if (Q.apply_loss) { double rho; vf::cauchy(P.cfg.cauchy_a, r[0] * r[0] + r[1] * r[1] + r[2] * r[2], rho, w); }
    for (int e = 0; e < 3; e++) R[rbase + 3 * t + e] = r[e] * w;
    for (int e = 0; e < 72; e++) Jo[jbase + (int64_t)72 * t + e] = J[e] * w;
    return;
  }
  t -= h->n_icp; rbase += 3 * h->n_icp; jbase += (int64_t)72 * h->n_icp;
  if (t < h->n_lps) {
    const double* c = W.d(OFF_LPS) + 9 * t; double r[3], J[36];
    vf::lps_eval(c, x + XP((int)c[7]), x + XP((int)c[8]), r, J);
    double w = 1.0;
    if (Q.apply_loss) { double rho; vf::cauchy(P.cfg.cauchy_a, r[0] * r[0] + r[1] * r[1] + r[2] * r[2], rho, w); }
    for (int e = 0; e < 3; e++) R[rbase + 3 * t + e] = r[e] * w;
    for (int e = 0; e < 36; e++) Jo[jbase + (int64_t)36 * t + e] = J[e] * w;
    return;
  }
}

// =================================================================================================================
// host
// =================================================================================================================
struct SlotMeta { int n_kf = 0, n_feat = 0, n_res = 0; int64_t n_jac = 0; int items = 0; bool set = false; bool on_device = false; bool prepped = false; int bytes = 0; };
struct PackPool;
static void pack_pool_destroy(PackPool* p);   // defined with the pool (the type is incomplete up here: a plain delete would skip the destructor and leak the threads)

struct vils_ba {
  vils_config cfg{};
  int max_windows = 0;
  cudaStream_t stream = nullptr, stream2 = nullptr;
  static constexpr int NPIPE = 8;
  cudaStream_t pipe[NPIPE] = {};                        // vils_ba_solve: chunked upload / solve / download pipeline
  cudaStream_t copy_stream = nullptr;                   // all uploads of the pipeline, back to back, ahead of the solves
  std::vector<cudaEvent_t> ev_chunk;                    // "chunk c is on the device"
  int n_sm = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_fork = nullptr, ev_join = nullptr;
  size_t blob_stride = 0;
  uint8_t* h_blob = nullptr; uint8_t* d_blob = nullptr;
  ScratchLayout sl{};
  double* d_scratch = nullptr;
  double* d_tscratch = nullptr; int n_smid = 0;        // per-SM scratch of solve_kernel (n_smid slots, indexed by %smid)
  size_t launch_smem = 0;                              // >= half the SM's shared memory: exactly one solve CTA per SM
  int64_t xstride = 0;
  double* d_xout = nullptr; double* h_xout = nullptr;
  vils_summary* d_sum = nullptr; vils_summary* h_sum = nullptr;
  double* d_lin = nullptr; double* h_lin = nullptr;
  double* d_er = nullptr; double* d_eJ = nullptr; int64_t er_stride = 0, eJ_stride = 0;
  std::vector<SlotMeta> meta;
  int h_in_smem = 0, hv_in_smem = 0; size_t smem_bytes = 0;
  float last_ms = 0; int last_launches = 0, last_cluster = 1; size_t last_h2d = 0, last_d2h = 0;
  double* d_shard = nullptr; size_t shard_doubles = 0;
  double* d_mws = nullptr; int32_t* d_miws = nullptr; MargParams mq{}; int64_t mws_doubles = 0; int mi_ints = 0;
  PackPool* pool = nullptr;                            // host packer threads (created on first use)
  ClScratch cl{}; double* d_clbuf = nullptr; int cl_windows = 0, cl_imu_slots = 0; size_t cl_smem = 0; bool cl_smem_h = false; int cluster_pref = 0;   // latency mode (0 auto, 1 off, 2/4/8/16 forced)
  bool cl16 = false;                                   // the device can co-schedule a cluster of 16 of these CTAs (non-portable size)
  ncclComm_t comm = nullptr; int comm_rank = 0, comm_size = 0;   // factor-sharded mode: one rank per GPU (vils_ba_sharded_init)
  double* d_lamred = nullptr;                          // [lam * own | own] for the final inverse-depth exchange
};

static inline size_t al16(size_t x) { return (x + 15) & ~size_t(15); }

static size_t blob_capacity(const vils_config& c) {
  const size_t N = c.max_kf, M = c.max_feat, Pn = c.max_proj, Ln = c.max_lidar, n = 15 * N + 7;
  size_t b = al16(sizeof(WinHdr));
  b += al16((16 * N + 8 + M) * 8) + al16(M + N) + al16(N * 467 * 8) + al16(N * 4);
  b += al16(14 * Pn * 8) + al16(4 * Pn * 4) + al16((M + 1) * 4) + al16(M * 4) + al16(N * N * 4 * 4) + al16(Pn * 4) + al16(N * N * 4);
  b += al16(7 * Ln * 8) + al16(2 * Ln * 4) + al16((N + 1) * 4) + al16(9 * Ln * 8) + al16(2 * Ln * 4) + al16((N + 1) * 4);
  b += al16(16 * 14 * 8) + al16(16 * 9 * 8);
  b += al16(n * n * 8) + al16(n * 8) + al16((2 * N + 2) * 9 * 8) + al16((2 * N + 2) * 16) + al16(n * 4);
  return (b + 255) & ~size_t(255);
}

static void launch_solve(vils_ba* ba, const SolveParams& P, int n, cudaStream_t s, bool allow_cluster = true);
extern "C" { static void nccl_comm_destroy(ncclComm_t c); }
static SolveParams make_params(vils_ba* ba, const vils_solve_opts* o) {
  SolveParams P{};
  P.blobs = ba->d_blob; P.blob_stride = (int64_t)ba->blob_stride;
  P.scratch = ba->d_scratch; P.sl = ba->sl;
  P.xout = ba->d_xout; P.xout_stride = ba->xstride; P.summary = ba->d_sum;
  const vils_config& c = ba->cfg;
  P.cfg.s_info = c.focal_length / 2.0;
  for (int i = 0; i < 3; i++) P.cfg.G[i] = c.gravity[i];
  P.cfg.tr_over_row = c.tr / c.row; P.cfg.half_row = c.row / 2.0;
  P.cfg.cauchy_a = c.cauchy_visual; P.cfg.huber_a = c.huber_lidar;
  P.cfg.use_td = c.estimate_td; P.cfg.est_ex = c.estimate_extrinsic;
  if (o) {
    P.mode = o->mode; P.max_iters = o->max_iters; P.mu = o->mu; P.lm_radius = o->lm_initial_radius; P.f_tol = o->function_tolerance;
    P.p_tol = o->parameter_tolerance; P.min_rel_dec = o->min_relative_decrease;
    P.time_cap_ns = (long long)(o->max_solver_time * 1e9);
  }
  P.Ncap = c.max_kf; P.Mcap = c.max_feat; P.h_in_smem = ba->h_in_smem; P.hv_in_smem = ba->hv_in_smem;
  P.lin_out = nullptr; P.slot0 = 0; P.prof = nullptr; P.do_prep = 0;
  P.tscratch = ba->d_tscratch;
  P.cl = ba->cl; P.clbuf = ba->d_clbuf; P.cl_imu_slots = ba->cl_imu_slots;
  return P;
}

// Cluster size for n windows: as many SMs per window as the device has to spare (8, 4 or 2), 1 = the one-CTA-per-window kernel.
static int pick_cluster(const vils_ba* ba, const SolveParams& P, int n) {
  if (P.slot0 + n > ba->cl_windows) return 1;               // the exchange areas cover slots [0, cl_windows)
  static const bool prof_cluster = getenv("VILS_PROF_CLUSTER") != nullptr;
  if (P.lin_out || (P.prof && !prof_cluster) || P.max_iters <= 0 || ba->cluster_pref == 1 || n > ba->cl_windows) return 1;
  if ((P.mode != VILS_MODE_GN || P.time_cap_ns > 0) && P.prof) return 1;      // the cluster phase profile belongs to the Gauss-Newton latency kernel
  const int gmax = ba->cl16 ? CL_MAX : CL_PORTABLE;
  if (ba->cluster_pref >= 2) { const int g = std::min(ba->cluster_pref, gmax); return n * g <= ba->n_sm ? g : 1; }
  for (int g = gmax; g >= 2; g >>= 1) if (n * g <= ba->n_sm) return g;
  return 1;
}
static void launch_solve(vils_ba* ba, const SolveParams& P, int n, cudaStream_t s, bool allow_cluster) {
  const int G = allow_cluster ? pick_cluster(ba, P, n) : 1;
  if (G > 1) {   // latency form: one thread-block cluster per window (ba_cluster.cuh)
    cudaLaunchConfig_t cfg{}; cudaLaunchAttribute at[1];
    cfg.gridDim = dim3(n * G); cfg.blockDim = dim3(SOLVE_THREADS); cfg.dynamicSmemBytes = ba->cl_smem; cfg.stream = s;
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const bool tr = P.mode != VILS_MODE_GN;
    if (tr || P.time_cap_ns > 0) {
      // trust-region solves (and time-capped ones): the one-CTA loop on CTA 0 of the cluster, linearisations served by all CTAs
      if (ba->cl_smem_h) { if (tr) cudaLaunchKernelEx(&cfg, solve_kernel<true, true, true>, P); else cudaLaunchKernelEx(&cfg, solve_kernel<true, false, true>, P); }
      else { if (tr) cudaLaunchKernelEx(&cfg, solve_kernel<false, true, true>, P); else cudaLaunchKernelEx(&cfg, solve_kernel<false, false, true>, P); }
    } else if (ba->cl_smem_h) cudaLaunchKernelEx(&cfg, solve_cluster_kernel<true>, P, ba->cl, ba->d_clbuf, ba->cl_imu_slots);
    else cudaLaunchKernelEx(&cfg, solve_cluster_kernel<false>, P, ba->cl, ba->d_clbuf, ba->cl_imu_slots);
    ba->last_cluster = G;
    return;
  }
  ba->last_cluster = 1;
  const bool both = ba->h_in_smem && ba->hv_in_smem, tr = P.mode != VILS_MODE_GN;
  const size_t sm = ba->launch_smem;
  if (both) { if (tr) solve_kernel<true, true><<<n, SOLVE_THREADS, sm, s>>>(P); else solve_kernel<true, false><<<n, SOLVE_THREADS, sm, s>>>(P); }
  else { if (tr) solve_kernel<false, true><<<n, SOLVE_THREADS, sm, s>>>(P); else solve_kernel<false, false><<<n, SOLVE_THREADS, sm, s>>>(P); }
}
__global__ void nsmid_kernel(int* out) { unsigned n; asm("mov.u32 %0, %%nsmid;" : "=r"(n)); *out = (int)n; }
// prep_kernel for the slots of [slot0, slot0 + n) whose per-window scratch (IMU sqrt_info, prior A / b0) is not up to date: the evaluate,
// marginalization and sharded kernels read it, solve_kernel does not write it any more (it works in per-SM scratch).
static void ensure_prepped(vils_ba* ba, int slot0, int n, cudaStream_t s) {
  SolveParams P = make_params(ba, nullptr);
  for (int k = slot0; k < slot0 + n;) {
    if (ba->meta[k].prepped) { k++; continue; }
    int e = k; while (e < slot0 + n && !ba->meta[e].prepped) e++;
    P.slot0 = k;
    prep_kernel<<<e - k, 256, 0, s>>>(P);
    for (int q = k; q < e; q++) ba->meta[q].prepped = true;
    k = e;
  }
}

extern "C" {

int vils_abi_version(void) { return VILS_ABI_VERSION; }
const char* vils_last_error(void) { return vils::last_error().c_str(); }

void vils_default_config(vils_config* cfg) {
  if (!cfg) return;
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->focal_length = 460.0;                                   // parameters.h:11
  cfg->gravity[2] = 9.795;                                     // config/mynteye_leishen_indoor.yaml:102
  cfg->tr = 0.0; cfg->row = 480.0;                             // yaml:12,118
  cfg->cauchy_visual = 1.0; cfg->huber_lidar = 0.1;            // estimator.cpp:1129 ; localMapping.cpp:597
  const double rli[9] = {-0.0320631, 0.000946093, -0.999485, -0.999482, -0.00274554, 0.0320604, -0.0027138, 0.999996, 0.00103363};  // yaml:43-47
  for (int i = 0; i < 9; i++) cfg->rlb[i] = rli[i];
  cfg->tlb[0] = 0.2; cfg->tlb[1] = -0.005; cfg->tlb[2] = -0.1;  // yaml:48-52
  cfg->estimate_extrinsic = 1; cfg->estimate_td = 1;            // yaml:25,112
  cfg->max_kf = 10; cfg->max_feat = 150; cfg->max_proj = 1400; cfg->max_lidar = 2000; cfg->device = 0;
  cfg->imu_noise[0] = 0.02065; cfg->imu_noise[1] = 0.00519; cfg->imu_noise[2] = 0.00667; cfg->imu_noise[3] = 0.00088056;   // yaml:81-86
}

void vils_default_solve_opts(vils_solve_opts* o) {
  if (!o) return;
  o->mode = VILS_MODE_GN; o->max_iters = 5; o->mu = 1e-8; o->lm_initial_radius = 1e4; o->function_tolerance = 1e-6;
  o->parameter_tolerance = 1e-8; o->min_relative_decrease = 1e-3; o->max_solver_time = 0.0;
}

int vils_ba_create(const vils_config* cfg, int32_t max_windows, vils_ba** out) {
  if (!cfg || !out || max_windows <= 0 || cfg->max_kf < 2 || cfg->max_kf > 32 || cfg->max_feat < 0 || cfg->max_proj < 0 || cfg->max_lidar < 0)
    return vils::fail(VILS_ERR_BAD_ARG, "vils_ba_create: bad argument");
  int st = vils::require_device(cfg->device);
  if (st) return st;
  vils_ba* ba = new vils_ba();
  ba->cfg = *cfg; ba->max_windows = max_windows;
  ba->meta.resize(max_windows);
  const int N = cfg->max_kf, M = cfg->max_feat, D = 15 * N + 7, Dv = 6 * N + 7, Dvp = (Dv + 3) & ~3, nb = (D + TB - 1) / TB;
  ba->blob_stride = blob_capacity(*cfg);
  // scratch layout
  ScratchLayout& s = ba->sl; int64_t o = 0;
  auto take = [&](int64_t n) { int64_t r = o; o += (n + 1) & ~int64_t(1); return r; };
  s.Dv_pad = Dvp;
  s.w_imu = take((int64_t)N * 225); s.E = take((int64_t)M * Dvp); s.part = take((int64_t)cfg->max_proj * PART_LD);
  s.pairpart = take(((int64_t)N * (N - 1) / 2 + 1) * PAIR_LD * PAIR_LD);   // + one all-zero block
  s.pairctx = take((int64_t)N * (N - 1) / 2 * vf::PCTX_LD);
  s.priorA = take((int64_t)D * D); s.priorb0 = take(D); s.lmsave = take(2 * (int64_t)M); s.imuprod = take((int64_t)N * IMU_PROD_LD);
  // shared-memory plan: prefer H and Hv both in shared memory, then Hv only, then neither
  int dev_smem = 0; cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device);
  const size_t budget = (size_t)dev_smem - 1024;
  // VILS_SMEM_BUDGET (experiments): cap the per-CTA shared-memory plan, e.g. to make two windows resident per SM
  const size_t plan_budget = getenv("VILS_SMEM_BUDGET") ? std::min(budget, (size_t)atol(getenv("VILS_SMEM_BUDGET"))) : budget;
  ba->h_in_smem = 1; ba->hv_in_smem = 1;
  if ((size_t)smem_layout(N, M, 1, 1).total * 8 > plan_budget) { ba->h_in_smem = 0; }
  if ((size_t)smem_layout(N, M, ba->h_in_smem, 1).total * 8 > plan_budget) { ba->hv_in_smem = 0; }
  ba->smem_bytes = (size_t)smem_layout(N, M, ba->h_in_smem, ba->hv_in_smem).total * 8;
  if (ba->smem_bytes > budget) { delete ba; return vils::fail(VILS_ERR_CAPACITY, "vils_ba_create: window too large for shared memory plan"); }
  s.Hg = take((int64_t)tri(nb) * TSZ);                      // also used by the cluster (latency) kernel, which keeps H in L2
  s.Hvg = ba->hv_in_smem ? 0 : take((int64_t)Dvp * Dvp);
  s.linvg = take((int64_t)nb * TB * TB);
  s.total = o;
  ba->xstride = 16 * N + 8 + M;
  {   // exchange area of the cluster (latency) kernel: one per window that can be solved that way at a time
    ClScratch& c = ba->cl; int64_t q = 0;
    auto tk = [&](int64_t n) { int64_t r = q; q += (n + 1) & ~int64_t(1); return r; };
    c.hvpart = tk((int64_t)CL_MAX * Dvp * Dvp); c.gvpart = tk((int64_t)CL_MAX * Dvp); c.lidblk = tk((int64_t)N * 28); c.gsc = tk(nb * TB); c.hdsc = tk(nb * TB);
    c.dxg = tk(nb * TB); c.lamg = tk(std::max(M, 1)); c.costp = tk(CL_MAX); c.flagg = tk(2);
    c.xg = tk(16 * N + 8 + M); c.cinvg = tk(std::max(M, 1)); c.glamg = tk(std::max(M, 1)); c.crawg = tk(std::max(M, 1)); c.sclg = tk(std::max(M, 1)); c.cmd = tk(8);
    c.total = q;
    ba->cl_smem_h = (size_t)smem_layout(N, M, 1, 0).total * 8 + 4096 <= budget;     // H of CTA 0 in shared memory when it fits
    const Smem Lc = smem_layout(N, M, ba->cl_smem_h ? 1 : 0, 0);
    ba->cl_smem = (size_t)Lc.total * 8;
    ba->cl_imu_slots = std::min(IMU_IDLE, (((N + 1) / 2) * 466) / IMU_SLOT2);
    if (const char* e = getenv("VILS_CLUSTER")) ba->cluster_pref = atoi(e);
  }
  const int64_t n_lin = (int64_t)D * D + D + 1;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { vils_ba_destroy(ba); return vils::fail_cuda(e_, #x); } } while (0)
  CK(cudaStreamCreateWithFlags(&ba->stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&ba->ev0)); CK(cudaEventCreate(&ba->ev1));
  CK(cudaStreamCreateWithFlags(&ba->stream2, cudaStreamNonBlocking));
  for (int k = 0; k < vils_ba::NPIPE; k++) CK(cudaStreamCreateWithFlags(&ba->pipe[k], cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&ba->copy_stream, cudaStreamNonBlocking));
  CK(cudaDeviceGetAttribute(&ba->n_sm, cudaDevAttrMultiProcessorCount, cfg->device));
  CK(cudaEventCreateWithFlags(&ba->ev_fork, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ba->ev_join, cudaEventDisableTiming));
  CK(cudaHostAlloc(&ba->h_blob, ba->blob_stride * max_windows, (getenv("VILS_WC") && atoi(getenv("VILS_WC"))) ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
  CK(cudaMalloc(&ba->d_blob, ba->blob_stride * max_windows));
  CK(cudaMalloc(&ba->d_scratch, (size_t)s.total * 8 * max_windows));
  { int* d_n = nullptr; CK(cudaMalloc(&d_n, sizeof(int))); nsmid_kernel<<<1, 1>>>(d_n); CK(cudaMemcpy(&ba->n_smid, d_n, sizeof(int), cudaMemcpyDeviceToHost)); cudaFree(d_n); }
  // VILS_SM_SCRATCH=1 (experiment, measured round 2: no change in DRAM traffic — the write-backs are local-memory traffic, not scratch — and 1.7 %
  // slower because prep runs in every launch): solve_kernel uses a scratch slot per SM instead of per window
  if (getenv("VILS_SM_SCRATCH") && atoi(getenv("VILS_SM_SCRATCH"))) CK(cudaMalloc(&ba->d_tscratch, (size_t)s.total * 8 * std::max(ba->n_smid, 1)));
  ba->launch_smem = std::max(ba->smem_bytes, (size_t)dev_smem / 2 + 1024);   // two solve CTAs can never share an SM (and its scratch slot)
  CK(cudaMalloc(&ba->d_xout, (size_t)ba->xstride * 8 * max_windows));
  CK(cudaDeviceGetAttribute(&ba->n_sm, cudaDevAttrMultiProcessorCount, cfg->device));
  ba->cl_windows = (ba->cl_smem + 4096 <= budget && ba->cl_imu_slots >= 1) ? std::min(max_windows, std::max(1, ba->n_sm / 2)) : 0;
  if (ba->cl_windows) CK(cudaMalloc(&ba->d_clbuf, (size_t)ba->cl.total * 8 * ba->cl_windows));
  CK(cudaMallocHost(&ba->h_xout, (size_t)ba->xstride * 8 * max_windows));
  CK(cudaMalloc(&ba->d_sum, sizeof(vils_summary) * max_windows));
  CK(cudaMallocHost(&ba->h_sum, sizeof(vils_summary) * max_windows));
  CK(cudaMalloc(&ba->d_lin, n_lin * 8)); CK(cudaMallocHost(&ba->h_lin, n_lin * 8));
  // cudaFuncAttributeMaxDynamicSharedMemorySize is per function AND per device, shared by every handle on that device: it is
  // raised once to the device's opt-in limit (what a launch actually uses is its own smem_bytes), so handles of different
  // capacities can live side by side.
  {
    static std::mutex attr_m; static bool attr_done[64] = {};
    std::lock_guard<std::mutex> l(attr_m);
    if (cfg->device >= 64 || !attr_done[cfg->device]) {
      const int lim = (int)budget;
      CK(cudaFuncSetAttribute(solve_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
      CK(cudaFuncSetAttribute(solve_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
      CK(cudaFuncSetAttribute(solve_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
      CK(cudaFuncSetAttribute(solve_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
      CK(cudaFuncSetAttribute(shard_lin_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
      CK(cudaFuncSetAttribute(shard_lin_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
      CK(cudaFuncSetAttribute(shard_upd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
      CK(cudaFuncSetAttribute(shard_upd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
      CK(cudaFuncSetAttribute(eval_proj_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
      CK(cudaFuncSetAttribute(eval_proj_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
      CK(cudaFuncSetAttribute(eval_proj_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
      CK(cudaFuncSetAttribute(margin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
      CK(cudaFuncSetAttribute(solve_cluster_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim - 4096));   // 2 KB of static shared memory (pair ownership table)
      CK(cudaFuncSetAttribute(solve_cluster_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim - 4096));
      CK(cudaFuncSetAttribute(solve_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim - 4096));
      CK(cudaFuncSetAttribute(solve_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim - 4096));
      CK(cudaFuncSetAttribute(solve_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim - 4096));
      CK(cudaFuncSetAttribute(solve_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim - 4096));
      cudaFuncSetAttribute(solve_kernel<true, true, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      cudaFuncSetAttribute(solve_kernel<true, false, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      cudaFuncSetAttribute(solve_kernel<false, true, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      cudaFuncSetAttribute(solve_kernel<false, false, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      cudaFuncSetAttribute(solve_cluster_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);     // clusters of 16: best effort
      cudaFuncSetAttribute(solve_cluster_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      cudaGetLastError();
      if (cfg->device < 64) attr_done[cfg->device] = true;
    }
  }
#undef CK
  {   // clusters of 16 CTAs are beyond the portable limit: used only if the device says it can co-schedule one
    cudaLaunchConfig_t qc{}; cudaLaunchAttribute at[1];
    qc.gridDim = dim3(16); qc.blockDim = dim3(SOLVE_THREADS); qc.dynamicSmemBytes = ba->cl_smem;
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    qc.attrs = at; qc.numAttrs = 1;
    int ncl = 0;
    const cudaError_t qe = ba->cl_smem_h ? cudaOccupancyMaxActiveClusters(&ncl, solve_cluster_kernel<true>, &qc) : cudaOccupancyMaxActiveClusters(&ncl, solve_cluster_kernel<false>, &qc);
    ba->cl16 = qe == cudaSuccess && ncl >= 1 && !(getenv("VILS_CLUSTER16") && atoi(getenv("VILS_CLUSTER16")) == 0);
    cudaGetLastError();
  }
  { uint16_t tbl[450]; vf::imu_build_table(tbl); cudaError_t e_ = cudaMemcpyToSymbol(g_imu_tbl, tbl, sizeof(tbl)); if (e_ != cudaSuccess) { vils_ba_destroy(ba); return vils::fail_cuda(e_, "imu table"); } }
  std::memset(ba->h_blob, 0, ba->blob_stride * max_windows);
  std::memset(ba->h_sum, 0, sizeof(vils_summary) * max_windows);
  *out = ba;
  return VILS_OK;
}

void vils_ba_destroy(vils_ba* ba) {
  if (!ba) return;
  cudaSetDevice(ba->cfg.device);
  if (ba->stream) cudaStreamSynchronize(ba->stream);
  cudaFreeHost(ba->h_blob); cudaFree(ba->d_blob); cudaFree(ba->d_scratch); cudaFree(ba->d_tscratch); cudaFree(ba->d_clbuf); cudaFree(ba->d_xout); cudaFreeHost(ba->h_xout);
  cudaFree(ba->d_sum); cudaFreeHost(ba->h_sum); cudaFree(ba->d_lin); cudaFreeHost(ba->h_lin); cudaFree(ba->d_er); cudaFree(ba->d_eJ); cudaFree(ba->d_mws); cudaFree(ba->d_miws); cudaFree(ba->d_shard);
  if (ba->ev0) cudaEventDestroy(ba->ev0);
  if (ba->ev1) cudaEventDestroy(ba->ev1);
  if (ba->ev_fork) cudaEventDestroy(ba->ev_fork);
  if (ba->ev_join) cudaEventDestroy(ba->ev_join);
  if (ba->stream2) cudaStreamDestroy(ba->stream2);
  for (int k = 0; k < vils_ba::NPIPE; k++) if (ba->pipe[k]) cudaStreamDestroy(ba->pipe[k]);
  if (ba->copy_stream) cudaStreamDestroy(ba->copy_stream);
  for (cudaEvent_t v : ba->ev_chunk) cudaEventDestroy(v);
  if (ba->stream) cudaStreamDestroy(ba->stream);
  pack_pool_destroy(ba->pool);
  if (ba->comm) nccl_comm_destroy(ba->comm);
  cudaFree(ba->d_lamred);
  delete ba;
}

// ---- host packer ------------------------------------------------------------------------------------------------------
// One window -> one blob in pinned staging.  Thread-safe for distinct slots (touches only its slot's blob and SlotMeta and the
// caller's PackScratch); never calls CUDA.  All orderings are stable counting sorts (landmark, keyframe pair, keyframe): no
// allocation after the first use of a scratch.  Errors come back as (code, message) so that worker threads can hand them on.
// 8-byte streaming store (movnti): the staging blob is written once and read only by the DMA engine, so it should not cost a
// read-for-ownership nor evict the caller's arrays from the cache.  VILS_PACK_NT=0 falls back to plain stores.
static const bool g_pack_nt = !(getenv("VILS_PACK_NT") && atoi(getenv("VILS_PACK_NT")) == 0);
static inline void nt_store(double* p, double v) {
  if (g_pack_nt) { long long b; std::memcpy(&b, &v, 8); _mm_stream_si64(reinterpret_cast<long long*>(p), b); }
  else *p = v;
}
// dst[s] = src[stride * idx[s] + c] for s < n as one sequential stream of 16-byte streaming stores fed by 4-wide AVX2 gathers (the packer is
// bound by its ~28 k scalar load / store pairs per window, not by memory bandwidth: the sources of one window sit in the L1 / L2 cache)
__attribute__((target("avx2"))) static void gather_stream(double* dst, const double* src, const int* idx, int n, int stride, int c) {
  int s = 0;
  if (!g_pack_nt) { for (; s < n; s++) dst[s] = src[(size_t)stride * idx[s] + c]; return; }
  while (s < n && (reinterpret_cast<uintptr_t>(dst + s) & 15)) { nt_store(dst + s, src[(size_t)stride * idx[s] + c]); s++; }
  const __m128i vstride = _mm_set1_epi32(stride), vc = _mm_set1_epi32(c);
  for (; s + 4 <= n; s += 4) {
    __m128i vi = _mm_loadu_si128(reinterpret_cast<const __m128i*>(idx + s));
    vi = _mm_add_epi32(_mm_mullo_epi32(vi, vstride), vc);
    const __m256d v = _mm256_i32gather_pd(src, vi, 8);
    _mm_stream_pd(dst + s, _mm256_castpd256_pd128(v)); _mm_stream_pd(dst + s + 2, _mm256_extractf128_pd(v, 1));
  }
  for (; s < n; s++) nt_store(dst + s, src[(size_t)stride * idx[s] + c]);
}
struct PackScratch {
  std::vector<int> cnt, ord, rank_of, lm_start, lm_feat, pord, pairs, pair_id, plo, pls, edo, eds, pblk, pcol;
};

static int pack_window(vils_ba* ba, int32_t slot, const vils_window* w, PackScratch& S, std::string& err) {
#define PFAIL(code, msg) do { err = (msg); return (code); } while (0)
  if (!ba || !w || slot < 0 || slot >= ba->max_windows) PFAIL(VILS_ERR_BAD_ARG, "vils_ba_set_window: bad slot / null");
  const vils_config& c = ba->cfg;
  const int N = w->n_kf, M = w->n_feat, np = w->n_proj, npl = w->n_plane, ned = w->n_edge;
  if (N < 2 || N > c.max_kf || M < 0 || M > c.max_feat || np < 0 || np > c.max_proj || npl < 0 || ned < 0 || npl + ned > c.max_lidar ||
      w->n_imu < 0 || w->n_imu > N - 1 + 0 || w->n_icp < 0 || w->n_icp > 16 || w->n_lps < 0 || w->n_lps > 16 || w->prior_n < 0 || w->prior_n > 15 * N + 7 ||
      w->prior_nblk < 0 || w->prior_nblk > 2 * N + 2)
    PFAIL(VILS_ERR_CAPACITY, "vils_ba_set_window: window exceeds the handle's capacity");
  if (!w->pose || !w->speedbias || !w->ex_pose || (M && !w->inv_depth) || (w->n_imu && (!w->imu || !w->imu_kf)) ||
      (np && (!w->pts_i || !w->pts_j || !w->vel_i || !w->vel_j || !w->td_i || !w->td_j || !w->row_i || !w->row_j || !w->kf_i || !w->kf_j || !w->feat)) ||
      (npl && (!w->plane_p || !w->plane_n || !w->plane_d || !w->plane_kf)) || (ned && (!w->edge_p || !w->edge_a || !w->edge_b || !w->edge_kf)) ||
      (w->n_icp && !w->icp) || (w->n_lps && !w->lps) || (w->prior_n && (!w->prior_J || !w->prior_r || !w->prior_blk || !w->prior_x0)))
    PFAIL(VILS_ERR_BAD_ARG, "vils_ba_set_window: null array");
  for (int k = 0; k < w->n_imu; k++) if (w->imu_kf[k] < 0 || w->imu_kf[k] + 1 >= N) PFAIL(VILS_ERR_BAD_ARG, "imu_kf out of range");
  for (int k = 0; k < np; k++)
    if (w->kf_i[k] < 0 || w->kf_i[k] >= N || w->kf_j[k] < 0 || w->kf_j[k] >= N || w->kf_i[k] == w->kf_j[k] || w->feat[k] < 0 || w->feat[k] >= M)
      PFAIL(VILS_ERR_BAD_ARG, "projection factor index out of range");
  for (int k = 0; k < npl; k++) if (w->plane_kf[k] < 0 || w->plane_kf[k] >= N) PFAIL(VILS_ERR_BAD_ARG, "plane_kf out of range");
  for (int k = 0; k < ned; k++) if (w->edge_kf[k] < 0 || w->edge_kf[k] >= N) PFAIL(VILS_ERR_BAD_ARG, "edge_kf out of range");
  for (int k = 0; k < w->n_icp; k++) for (int b = 0; b < 4; b++) if (w->icp[k].kf[b] < 0 || w->icp[k].kf[b] >= N) PFAIL(VILS_ERR_BAD_ARG, "icp kf out of range");
  for (int k = 0; k < w->n_lps; k++) for (int b = 0; b < 2; b++) if (w->lps[k].kf[b] < 0 || w->lps[k].kf[b] >= N) PFAIL(VILS_ERR_BAD_ARG, "lps kf out of range");
  // the prior's block table is validated BEFORE anything is written: a rejected window leaves the slot exactly as it was
  const int n = w->prior_n;
  int gs = 0, lsz = 0;
  S.pblk.assign(4 * (size_t)w->prior_nblk, 0); S.pcol.clear();
  for (int b = 0; b < w->prior_nblk; b++) {
    const int id = w->prior_blk[b], type = VILS_BLK_TYPE(id), idx = VILS_BLK_INDEX(id);
    if (type < 0 || type > 3 || ((type == VILS_BLK_POSE || type == VILS_BLK_SPEEDBIAS) && idx >= N)) PFAIL(VILS_ERR_BAD_ARG, "prior block id out of range");
    const int g = (type == VILS_BLK_POSE || type == VILS_BLK_EXPOSE) ? 7 : type == VILS_BLK_SPEEDBIAS ? 9 : 1;
    const int l = (type == VILS_BLK_POSE || type == VILS_BLK_EXPOSE) ? 6 : type == VILS_BLK_SPEEDBIAS ? 9 : 1;
    const int off = type == VILS_BLK_POSE ? 15 * idx : type == VILS_BLK_SPEEDBIAS ? 15 * idx + 6 : type == VILS_BLK_EXPOSE ? 15 * N : 15 * N + 6;
    S.pblk[4 * b] = type; S.pblk[4 * b + 1] = idx; S.pblk[4 * b + 2] = gs; S.pblk[4 * b + 3] = lsz;
    for (int a = 0; a < l; a++) S.pcol.push_back(off + a);
    gs += g; lsz += l;
  }
  if (lsz != n) PFAIL(VILS_ERR_BAD_ARG, "prior_n does not match the local sizes of prior_blk");

  // ---- orderings (stable counting sorts)
  // landmark order: factors by feature index, original order inside a feature
  S.cnt.assign((size_t)M + 1, 0);
  for (int k = 0; k < np; k++) S.cnt[w->feat[k] + 1]++;
  S.lm_feat.clear(); S.lm_start.clear();
  for (int f = 0; f < M; f++) { if (S.cnt[f + 1]) { S.lm_feat.push_back(f); S.lm_start.push_back(S.cnt[f]); } S.cnt[f + 1] += S.cnt[f]; }
  S.lm_start.push_back(np);
  const int nlm = (int)S.lm_feat.size();
  S.ord.resize(np); S.rank_of.resize(np);
  for (int k = 0; k < np; k++) S.ord[S.cnt[w->feat[k]]++] = k;
  for (int r = 0; r < nlm; r++) {
    const int s0 = S.lm_start[r], s1 = S.lm_start[r + 1], anchor = w->kf_i[S.ord[s0]];
    for (int s = s0; s < s1; s++) {
      if (w->kf_i[S.ord[s]] != anchor) PFAIL(VILS_ERR_BAD_ARG, "factors of one feature must share the anchor keyframe");
      // the gather assumes anchor < observer (a feature's anchor is its first observation, estimator.cpp:1201-1206)
      if (w->kf_i[S.ord[s]] >= w->kf_j[S.ord[s]]) PFAIL(VILS_ERR_BAD_ARG, "anchor keyframe must precede the observing keyframe");
      S.rank_of[s] = r;
    }
  }
  // pair order: landmark-sorted positions grouped by (kf_i, kf_j)
  S.cnt.assign((size_t)N * N + 1, 0);
  for (int s = 0; s < np; s++) S.cnt[w->kf_i[S.ord[s]] * N + w->kf_j[S.ord[s]] + 1]++;
  S.pairs.clear(); S.pair_id.assign((size_t)N * N, -1);
  for (int k = 0; k < N * N; k++) {
    if (S.cnt[k + 1]) { S.pair_id[k] = (int)S.pairs.size() / 4; S.pairs.push_back(S.cnt[k]); S.pairs.push_back(S.cnt[k + 1]); S.pairs.push_back(k / N); S.pairs.push_back(k % N); }
    S.cnt[k + 1] += S.cnt[k];
  }
  const int npair = (int)S.pairs.size() / 4;
  S.pord.resize(np);
  for (int s = 0; s < np; s++) S.pord[S.cnt[w->kf_i[S.ord[s]] * N + w->kf_j[S.ord[s]]]++] = s;
  auto sort_by_kf = [&](int cntf, const int32_t* kf, std::vector<int>& o, std::vector<int>& start) {
    start.assign((size_t)N + 1, 0);
    for (int k = 0; k < cntf; k++) start[kf[k] + 1]++;
    for (int k = 0; k < N; k++) start[k + 1] += start[k];
    o.resize(cntf); S.cnt.assign(start.begin(), start.end());
    for (int k = 0; k < cntf; k++) o[S.cnt[kf[k]]++] = k;
  };
  sort_by_kf(npl, w->plane_kf, S.plo, S.pls); sort_by_kf(ned, w->edge_kf, S.edo, S.eds);

  // ---- blob
  uint8_t* base = ba->h_blob + (size_t)slot * ba->blob_stride;
  WinHdr* h = reinterpret_cast<WinHdr*>(base);
  std::memset(h, 0, sizeof(WinHdr));
  h->n_kf = N; h->n_feat = M; h->n_imu = w->n_imu; h->n_proj = np; h->n_plane = npl; h->n_edge = ned; h->n_icp = w->n_icp; h->n_lps = w->n_lps;
  h->prior_n = w->prior_n; h->prior_nblk = w->prior_nblk; h->n_lm = nlm; h->n_pair = npair;
  size_t o = al16(sizeof(WinHdr));
  auto place = [&](int which, size_t bytes) { h->off[which] = (int32_t)o; uint8_t* p = base + o; o += al16(bytes); return p; };
  double* x = (double*)place(OFF_X, (size_t)(16 * N + 8 + M) * 8);
  std::memcpy(x, w->pose, sizeof(double) * 7 * N); std::memcpy(x + 7 * N, w->speedbias, sizeof(double) * 9 * N);
  std::memcpy(x + 16 * N, w->ex_pose, sizeof(double) * 7); x[16 * N + 7] = w->td;
  if (M) std::memcpy(x + 16 * N + 8, w->inv_depth, sizeof(double) * M);
  uint8_t* fx = place(OFF_FIXED, (size_t)M + N);
  for (int f = 0; f < M; f++) fx[f] = w->depth_fixed ? w->depth_fixed[f] : 0;
  for (int k = 0; k < N; k++) fx[M + k] = w->kf_fixed ? w->kf_fixed[k] : 0;
  double* imu = (double*)place(OFF_IMU, (size_t)w->n_imu * 467 * 8);
  if (w->n_imu) std::memcpy(imu, w->imu, sizeof(vils_preint) * w->n_imu);
  int32_t* ikf = (int32_t*)place(OFF_IMU_KF, (size_t)w->n_imu * 4);
  for (int k = 0; k < w->n_imu; k++) ikf[k] = w->imu_kf[k];
  // SoA arrays are written one field at a time: each destination is ONE sequential stream (streaming stores, the blob is only ever read
  // by the copy engine), the sources are read nearly in order (callers hand factors over grouped by feature)
  double* pj = (double*)place(OFF_PROJ, (size_t)14 * np * 8);
  int32_t* pix = (int32_t*)place(OFF_PROJ_IDX, (size_t)4 * np * 4);
  {
    const int* od = S.ord.data();
    auto field = [&](int a, const double* src, int stride, int c) { gather_stream(pj + (size_t)a * np, src, od, np, stride, c); };
    for (int c2 = 0; c2 < 3; c2++) { field(c2, w->pts_i, 3, c2); field(3 + c2, w->pts_j, 3, c2); }
    for (int c2 = 0; c2 < 2; c2++) { field(6 + c2, w->vel_i, 2, c2); field(8 + c2, w->vel_j, 2, c2); }
    field(10, w->td_i, 1, 0); field(11, w->td_j, 1, 0); field(12, w->row_i, 1, 0); field(13, w->row_j, 1, 0);
    for (int s = 0; s < np; s++) pix[s] = w->kf_i[od[s]];
    for (int s = 0; s < np; s++) pix[np + s] = w->kf_j[od[s]];
    for (int s = 0; s < np; s++) pix[2 * np + s] = S.rank_of[s];
    for (int s = 0; s < np; s++) pix[3 * np + s] = od[s];
  }
  int32_t* ls = (int32_t*)place(OFF_LM_START, (size_t)(nlm + 1) * 4); for (int k = 0; k <= nlm; k++) ls[k] = S.lm_start[k];
  int32_t* lf = (int32_t*)place(OFF_LM_FEAT, (size_t)nlm * 4); for (int k = 0; k < nlm; k++) lf[k] = S.lm_feat[k];
  int32_t* pr = (int32_t*)place(OFF_PAIR, (size_t)npair * 16);
  for (int k = 0; k < npair; k++) { pr[4 * k] = S.pairs[4 * k]; pr[4 * k + 1] = S.pairs[4 * k + 1]; pr[4 * k + 2] = S.pairs[4 * k + 2]; pr[4 * k + 3] = S.pairs[4 * k + 3]; }
  int32_t* pp = (int32_t*)place(OFF_PAIR_PERM, (size_t)np * 4); for (int k = 0; k < np; k++) pp[k] = S.pord[k];
  int32_t* pi = (int32_t*)place(OFF_PAIR_ID, (size_t)N * N * 4); for (int k = 0; k < N * N; k++) pi[k] = S.pair_id[k];
  // LiDAR points go to the body frame once, here: p_b = RLB^T (p_l - TLB)   (estimator.cpp:449-451)
  auto to_body = [&](const double* p, double* out) {
    const double d[3] = {p[0] - c.tlb[0], p[1] - c.tlb[1], p[2] - c.tlb[2]};
    for (int a = 0; a < 3; a++) out[a] = c.rlb[0 * 3 + a] * d[0] + c.rlb[1 * 3 + a] * d[1] + c.rlb[2 * 3 + a] * d[2];
  };
  double* pl = (double*)place(OFF_PLANE, (size_t)7 * npl * 8);
  int32_t* plx = (int32_t*)place(OFF_PLANE_IDX, (size_t)2 * npl * 4);
  for (int s = 0; s < npl; s++) {
    const int k = S.plo[s]; double pb[3]; to_body(w->plane_p + 3 * k, pb);
    for (int a = 0; a < 3; a++) nt_store(pl + (size_t)a * npl + s, pb[a]);
    plx[s] = w->plane_kf[k]; plx[npl + s] = k;
  }
  for (int a = 0; a < 3; a++) gather_stream(pl + (size_t)(3 + a) * npl, w->plane_n, S.plo.data(), npl, 3, a);
  gather_stream(pl + (size_t)6 * npl, w->plane_d, S.plo.data(), npl, 1, 0);
  int32_t* plst = (int32_t*)place(OFF_PLANE_START, (size_t)(N + 1) * 4); for (int k = 0; k <= N; k++) plst[k] = S.pls[k];
  double* ed = (double*)place(OFF_EDGE, (size_t)9 * ned * 8);
  int32_t* edx = (int32_t*)place(OFF_EDGE_IDX, (size_t)2 * ned * 4);
  for (int s = 0; s < ned; s++) {
    const int k = S.edo[s]; double pb[3]; to_body(w->edge_p + 3 * k, pb);
    for (int a = 0; a < 3; a++) nt_store(ed + (size_t)a * ned + s, pb[a]);
    edx[s] = w->edge_kf[k]; edx[ned + s] = k;
  }
  for (int a = 0; a < 3; a++) {
    gather_stream(ed + (size_t)(3 + a) * ned, w->edge_a, S.edo.data(), ned, 3, a);
    gather_stream(ed + (size_t)(6 + a) * ned, w->edge_b, S.edo.data(), ned, 3, a);
  }
  int32_t* edst = (int32_t*)place(OFF_EDGE_START, (size_t)(N + 1) * 4); for (int k = 0; k <= N; k++) edst[k] = S.eds[k];
  double* ic = (double*)place(OFF_ICP, (size_t)w->n_icp * 14 * 8);
  for (int k = 0; k < w->n_icp; k++) {
    const vils_icp& q = w->icp[k]; double* d = ic + 14 * k;
    d[0] = q.ta; d[1] = q.tb; d[2] = q.tc; d[3] = q.td; d[4] = q.ti; d[5] = q.tj; d[6] = q.trans_t[0]; d[7] = q.trans_t[1]; d[8] = q.trans_t[2]; d[9] = q.sqrt_info;
    for (int b = 0; b < 4; b++) d[10 + b] = q.kf[b];
  }
  double* lp = (double*)place(OFF_LPS, (size_t)w->n_lps * 9 * 8);
  for (int k = 0; k < w->n_lps; k++) {
    const vils_lps& q = w->lps[k]; double* d = lp + 9 * k;
    d[0] = q.tl; d[1] = q.tr; d[2] = q.tk; for (int a = 0; a < 4; a++) d[3 + a] = q.q[a]; d[7] = q.kf[0]; d[8] = q.kf[1];
  }
  double* PJ = (double*)place(OFF_PRIOR_J, (size_t)n * n * 8); if (n) std::memcpy(PJ, w->prior_J, sizeof(double) * n * n);
  double* PR = (double*)place(OFF_PRIOR_R, (size_t)n * 8); if (n) std::memcpy(PR, w->prior_r, sizeof(double) * n);
  double* PX = (double*)place(OFF_PRIOR_X0, (size_t)gs * 8); if (gs) std::memcpy(PX, w->prior_x0, sizeof(double) * gs);
  int32_t* PB = (int32_t*)place(OFF_PRIOR_BLK, (size_t)4 * w->prior_nblk * 4); for (size_t k = 0; k < S.pblk.size(); k++) PB[k] = S.pblk[k];
  int32_t* PC = (int32_t*)place(OFF_PRIOR_COL, (size_t)n * 4); for (int k = 0; k < n; k++) PC[k] = S.pcol[k];
  SlotMeta& m = ba->meta[slot];
  if (o > ba->blob_stride) { m.set = false; PFAIL(VILS_ERR_CAPACITY, "vils_ba_set_window: blob overflow"); }   // cannot happen for windows inside the capacities checked above
  h->bytes = (int32_t)o;
  _mm_sfence();                                             // streaming stores visible before the copy engine is pointed at the blob
  m.set = true; m.on_device = false; m.prepped = false; m.n_kf = N; m.n_feat = M; m.bytes = (int)o;
  m.n_res = 15 * w->n_imu + 2 * np + npl + 3 * ned + 3 * w->n_icp + 3 * w->n_lps + n;
  m.n_jac = (int64_t)450 * w->n_imu + (int64_t)40 * np + 6 * npl + 18 * ned + 72 * w->n_icp + 36 * w->n_lps;
  m.items = w->n_imu + np + npl + ned + w->n_icp + w->n_lps + n;
  return VILS_OK;
#undef PFAIL
}

int vils_ba_set_window(vils_ba* ba, int32_t slot, const vils_window* w) {
  static thread_local PackScratch S;
  std::string err;
  const int st = pack_window(ba, slot, w, S, err);
  return st == VILS_OK ? VILS_OK : vils::fail(st, err);
}

// Persistent host threads that pack windows (vils_ba_set_windows / vils_ba_solve_windows).  Work items are window indices
// handed out in ascending order by an atomic counter; the calling thread packs too, so a pool of T threads gives T + 1 packers.
struct PackPool {
  std::vector<std::thread> th;
  std::mutex m; std::condition_variable cv;
  uint64_t generation = 0; bool stop = false;
  // current job
  vils_ba* ba = nullptr; const vils_window* ws = nullptr; int slot0 = 0, n = 0;
  std::vector<int> chunk_of;                               // window -> chunk (chunks may have different sizes: small ones first)
  std::atomic<int> next{0}, active{0};
  std::vector<std::atomic<int>> chunk_done;
  std::atomic<int> err_code{0}; std::string err_msg;
  explicit PackPool(int T) : chunk_done(0) { for (int t = 0; t < T; t++) th.emplace_back([this] { run(); }); }
  ~PackPool() { { std::lock_guard<std::mutex> l(m); stop = true; } cv.notify_all(); for (auto& t : th) t.join(); }
  void work(PackScratch& S) {
    for (;;) {
      const int k = next.fetch_add(1, std::memory_order_relaxed);
      if (k >= n) break;
      if (err_code.load(std::memory_order_relaxed) == 0) {
        std::string e;
        const int st = pack_window(ba, slot0 + k, ws + k, S, e);
        if (st != VILS_OK) { std::lock_guard<std::mutex> l(m); if (err_code.load() == 0) { err_msg = e; err_code.store(st); } }
      }
      chunk_done[chunk_of[k]].fetch_add(1, std::memory_order_release);
    }
  }
  void run() {
    PackScratch S; uint64_t seen = 0;
    for (;;) {
      { std::unique_lock<std::mutex> l(m); cv.wait(l, [&] { return stop || generation != seen; }); if (stop) return; seen = generation; active.fetch_add(1); }
      work(S);
      active.fetch_sub(1, std::memory_order_release);
    }
  }
  // bounds: chunk c covers windows [bounds[c], bounds[c + 1])
  void begin(vils_ba* b, const vils_window* w, int s0, int cnt, const std::vector<int>& bounds) {
    while (active.load(std::memory_order_acquire) != 0) std::this_thread::yield();      // stragglers of the previous job
    { std::lock_guard<std::mutex> l(m);
      ba = b; ws = w; slot0 = s0; n = cnt; err_code.store(0); err_msg.clear();
      const int nc = (int)bounds.size() - 1;
      chunk_of.resize(cnt);
      for (int c = 0; c < nc; c++) for (int k = bounds[c]; k < bounds[c + 1]; k++) chunk_of[k] = c;
      if ((int)chunk_done.size() < nc) { std::vector<std::atomic<int>> v(nc); chunk_done.swap(v); }
      for (auto& c : chunk_done) c.store(0, std::memory_order_relaxed);
      next.store(0); generation++; }
    cv.notify_all();
  }
  // the caller packs until chunk c is complete (or nothing is left to hand out, then it yields)
  void wait_chunk(int c, int cn, PackScratch& S) {
    while (chunk_done[c].load(std::memory_order_acquire) < cn) {
      const int k = next.fetch_add(1, std::memory_order_relaxed);
      if (k < n) {
        if (err_code.load(std::memory_order_relaxed) == 0) {
          std::string e;
          const int st = pack_window(ba, slot0 + k, ws + k, S, e);
          if (st != VILS_OK) { std::lock_guard<std::mutex> l(m); if (err_code.load() == 0) { err_msg = e; err_code.store(st); } }
        }
        chunk_done[chunk_of[k]].fetch_add(1, std::memory_order_release);
      } else std::this_thread::yield();
    }
  }
};

extern "C++" { static void pack_pool_destroy(PackPool* p) { delete p; } }

static PackPool* pack_pool(vils_ba* ba) {
  if (!ba->pool) {
    // one rank per GPU on a node: the host cores are shared by LOCAL_WORLD_SIZE processes (torchrun exports it); oversubscribing them with a
    // full-size pool per rank made the 8-GPU end-to-end rate collapse (round 2: 204 k solves/s with 8 x 32 threads on 32 cores)
    int T = (int)std::thread::hardware_concurrency();
    if (const char* lw = getenv("LOCAL_WORLD_SIZE")) { const int nw = atoi(lw); if (nw > 1) T = std::max(1, T / nw); }
    if (const char* e = getenv("VILS_PACK_THREADS")) T = atoi(e);
    T = std::max(1, std::min(T, 64)) - 1;                         // the calling thread is one of the packers
    ba->pool = new PackPool(T);
  }
  return ba->pool;
}

// Stage n windows into slots [slot0, slot0 + n) on all host threads.
int vils_ba_set_windows(vils_ba* ba, int32_t slot0, int32_t n, const vils_window* ws) {
  if (!ba || !ws || n <= 0 || slot0 < 0 || slot0 + n > ba->max_windows) return vils::fail(VILS_ERR_BAD_ARG, "vils_ba_set_windows: bad argument");
  static thread_local PackScratch S;
  if (n == 1) return vils_ba_set_window(ba, slot0, ws);
  PackPool* pool = pack_pool(ba);
  pool->begin(ba, ws, slot0, n, std::vector<int>{0, n});
  pool->wait_chunk(0, n, S);
  if (pool->err_code.load()) return vils::fail(pool->err_code.load(), pool->err_msg);
  return VILS_OK;
}

static int check_n(vils_ba* ba, int n, const char* who) {
  if (!ba || n <= 0 || n > ba->max_windows) return vils::fail(VILS_ERR_BAD_ARG, std::string(who) + ": bad window count");
  for (int k = 0; k < n; k++) if (!ba->meta[k].set) return vils::fail(VILS_ERR_BAD_ARG, std::string(who) + ": slot not set");
  return cudaSetDevice(ba->cfg.device) == cudaSuccess ? VILS_OK : vils::fail(VILS_ERR_CUDA, "cudaSetDevice");
}
// Device-side entry points work on the UPLOADED blob: a slot re-staged by vils_ba_set_window since its last upload is stale.
static int check_on_device(vils_ba* ba, int slot0, int n, const char* who) {
  for (int k = slot0; k < slot0 + n; k++)
    if (!ba->meta[k].on_device) return vils::fail(VILS_ERR_BAD_ARG, std::string(who) + ": slot staged but not uploaded (call vils_ba_upload / vils_ba_solve first)");
  return VILS_OK;
}
// A launch-configuration error is reported by cudaGetLastError() right after the launch, never by the stream synchronisation.
#define VILS_LAUNCH_CHECK(what) do { const cudaError_t le_ = cudaGetLastError(); if (le_ != cudaSuccess) return vils::fail_cuda(le_, what); } while (0)

int vils_ba_upload(vils_ba* ba, int32_t n) {
  int st = check_n(ba, n, "vils_ba_upload"); if (st) return st;
  size_t width = 0; for (int k = 0; k < n; k++) width = std::max(width, (size_t)ba->meta[k].bytes);
  ba->last_h2d = width * n;
  cudaError_t e = cudaMemcpy2DAsync(ba->d_blob, ba->blob_stride, ba->h_blob, ba->blob_stride, width, n, cudaMemcpyHostToDevice, ba->stream);
  if (e != cudaSuccess) return vils::fail_cuda(e, "upload");
  SolveParams P = make_params(ba, nullptr);
  prep_kernel<<<n, 256, 0, ba->stream>>>(P);
  VILS_LAUNCH_CHECK("prep_kernel launch");
  e = cudaStreamSynchronize(ba->stream);
  if (e != cudaSuccess) return vils::fail_cuda(e, "upload sync");
  for (int k = 0; k < n; k++) { ba->meta[k].on_device = true; ba->meta[k].prepped = true; }
  return VILS_OK;
}

int vils_ba_solve_device(vils_ba* ba, int32_t n, const vils_solve_opts* opts) {
  int st = check_n(ba, n, "vils_ba_solve_device"); if (st) return st;
  if (!opts || opts->max_iters < 0 || (opts->mode != VILS_MODE_GN && opts->mode != VILS_MODE_LM && opts->mode != VILS_MODE_DOGLEG) || opts->max_solver_time < 0)
    return vils::fail(VILS_ERR_BAD_ARG, "solve: bad options");
  st = check_on_device(ba, 0, n, "vils_ba_solve_device"); if (st) return st;
  SolveParams P = make_params(ba, opts);
  static const bool prof = getenv("VILS_PROF") != nullptr;
  long long* d_prof = nullptr;
  if (prof) { cudaMalloc(&d_prof, 24 * sizeof(long long)); cudaMemset(d_prof, 0, 24 * sizeof(long long)); P.prof = d_prof; }
  cudaEventRecord(ba->ev0, ba->stream);
  launch_solve(ba, P, n, ba->stream);
  VILS_LAUNCH_CHECK("solve_kernel launch");
  cudaEventRecord(ba->ev1, ba->stream);
  cudaError_t e = cudaStreamSynchronize(ba->stream);
  if (e != cudaSuccess) return vils::fail_cuda(e, "solve_kernel");
  cudaEventElapsedTime(&ba->last_ms, ba->ev0, ba->ev1);
  if (prof) {
    long long h[24]; cudaMemcpy(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost); cudaFree(d_prof);
    static const char* names[24] = {"pair_pass", "landmark_reduce", "schur_syrk", "zero+gather", "imu", "lidar", "icp+prior+sum", "damp_fix", "cholesky", "backsub", "apply_step", "final_cost", "  chol:diag(w0)", "  chol:phaseB-wait", "  chol:panel", "  chol:phaseA", "  pair:eval(t0)", "  pair:eval-wait", "  pair:accum(w0)", "  pair:accum-wait", "  imu:loads+core", "  imu:assemble", "  imu:whiten", "  imu:JtJ+store"};
    if (ba->last_cluster > 1) {
      static const char* cn[12] = {"F factor pass", "  sync", "L landmarks+schur", "  sync", "G gather", "  sync", "C adds+damp", "C cholesky", "C backsub+dx", "  sync (others wait for C)", "U update+exchange", "zero H"};
      long long t2 = 0; for (int i = 0; i < 12; i++) t2 += h[i];
      fprintf(stderr, "[VILS_PROF] cluster of %d, CTA 0, %d windows, %.3f ms; SM cycles per phase (sum over iterations):\n", ba->last_cluster, n, ba->last_ms);
      for (int i = 0; i < 12; i++) fprintf(stderr, "  %-28s %10lld  %5.1f%%\n", cn[i], h[i], 100.0 * h[i] / (t2 ? t2 : 1));
      ba->last_launches = 1;
      return VILS_OK;
    }
    long long tot = 0; for (int i = 0; i < 12; i++) tot += h[i];
    fprintf(stderr, "[VILS_PROF] block 0, %d windows, %.3f ms; SM cycles per phase (sum over iterations):\n", n, ba->last_ms);
    for (int i = 0; i < 24; i++) fprintf(stderr, "  %-16s %10lld  %5.1f%%\n", names[i], h[i], 100.0 * h[i] / (tot ? tot : 1));
  }
  ba->last_launches = 1;
  return VILS_OK;
}

int vils_ba_download(vils_ba* ba, int32_t n) {
  int st = check_n(ba, n, "vils_ba_download"); if (st) return st;
  ba->last_d2h = (size_t)ba->xstride * 8 * n + sizeof(vils_summary) * n;
  cudaMemcpyAsync(ba->h_xout, ba->d_xout, (size_t)ba->xstride * 8 * n, cudaMemcpyDeviceToHost, ba->stream);
  cudaMemcpyAsync(ba->h_sum, ba->d_sum, sizeof(vils_summary) * n, cudaMemcpyDeviceToHost, ba->stream);
  cudaError_t e = cudaStreamSynchronize(ba->stream);
  return e == cudaSuccess ? VILS_OK : vils::fail_cuda(e, "download");
}

// The call a host makes per optimization(): host buffers in, host buffers out.  Slots are processed in chunks of half
// the SM count (one CTA per window, one CTA per SM), each chunk on one of three streams: the host->device copy of chunk
// c+1 runs on the copy engine while chunk c is being solved, the small device->host copies ride behind each solve, and a
// single synchronisation closes the call (a one-window call therefore pays one sync instead of three).
// ws != nullptr (vils_ba_solve_windows): the windows are packed from the caller's arrays INSIDE the pipeline — the host threads
// of the pack pool stage chunk c+1 while chunk c is being copied and solved — so the call runs from vils_window arrays to states.
static int solve_pipeline(vils_ba* ba, int32_t n, const vils_window* ws, const vils_solve_opts* opts, const char* who) {
  if (!ba || n <= 0 || n > ba->max_windows) return vils::fail(VILS_ERR_BAD_ARG, std::string(who) + ": bad window count");
  if (!opts || opts->max_iters < 0 || (opts->mode != VILS_MODE_GN && opts->mode != VILS_MODE_LM && opts->mode != VILS_MODE_DOGLEG) || opts->max_solver_time < 0)
    return vils::fail(VILS_ERR_BAD_ARG, "solve: bad options");
  if (!ws) for (int k = 0; k < n; k++) if (!ba->meta[k].set) return vils::fail(VILS_ERR_BAD_ARG, std::string(who) + ": slot not set");
  if (cudaSetDevice(ba->cfg.device) != cudaSuccess) return vils::fail(VILS_ERR_CUDA, "cudaSetDevice");
  static const int chunk_env = getenv("VILS_CHUNK") ? atoi(getenv("VILS_CHUNK")) : 0;
  static const int ns_env = getenv("VILS_STREAMS") ? atoi(getenv("VILS_STREAMS")) : 0;
  const int chunk = chunk_env > 0 ? chunk_env : std::max(1, ba->n_sm / 2);
  // Uploads go back to back on their own stream, ahead of the solves (an upload queued on a solve stream would wait for the previous chunk of that
  // stream to finish: measured, the copy engine then idles 60 % of the time and the SMs wait for data); the solve streams only need to be enough
  // for the chunks in flight to cover every SM.
  static const bool own_copy_stream = !(getenv("VILS_COPY_STREAM") && atoi(getenv("VILS_COPY_STREAM")) == 0);
  const int NS = std::min<int>(vils_ba::NPIPE, ns_env > 0 ? ns_env : ws ? vils_ba::NPIPE : std::max(3, (ba->n_sm + chunk - 1) / chunk + 1));
  SolveParams P = make_params(ba, opts);
  const bool both = ba->h_in_smem && ba->hv_in_smem;
  size_t h2d = 0; int launches = 0;
  static thread_local PackScratch S;
  // Chunk boundaries.  With in-pipeline packing the chunks repeat the pattern n_sm / 8, n_sm / 4, n_sm / 2, rest of the device: the GPU starts after
  // ~0.1 ms of packing instead of waiting for a half-device chunk, and because one pattern covers every SM exactly once, the chunk that arrives
  // next is as large as the one that retires next (uniform chunks after a small head leave a ragged tail: measured 6.45 vs 6.1 ms per 592 windows).
  std::vector<int> bounds{0};
  if (ws && n > chunk && !chunk_env) {
    const int a = std::max(1, ba->n_sm / 8), b = std::max(1, ba->n_sm / 4), c2 = std::max(1, ba->n_sm / 2);
    const int pat[4] = {a, b, c2, std::max(1, ba->n_sm - a - b - c2)};
    for (int k = 0; bounds.back() < n; k++) bounds.push_back(std::min(n, bounds.back() + pat[k & 3]));
  }
  while (bounds.back() < n) bounds.push_back(std::min(n, bounds.back() + chunk));
  const int nchunks = (int)bounds.size() - 1;
  PackPool* pool = nullptr;
  if (ws && n > 1) { pool = pack_pool(ba); pool->begin(ba, ws, 0, n, bounds); }
  cudaEventRecord(ba->ev_fork, ba->stream);                 // order after anything still queued on the handle's main stream
  for (int k = 0; k < NS; k++) cudaStreamWaitEvent(ba->pipe[k], ba->ev_fork, 0);
  if (own_copy_stream) {
    cudaStreamWaitEvent(ba->copy_stream, ba->ev_fork, 0);
    while ((int)ba->ev_chunk.size() < nchunks) { cudaEvent_t v; if (cudaEventCreateWithFlags(&v, cudaEventDisableTiming) != cudaSuccess) return vils::fail(VILS_ERR_CUDA, "cudaEventCreate"); ba->ev_chunk.push_back(v); }
  }
  int err = VILS_OK; std::string err_msg; cudaError_t le = cudaSuccess;
  static const bool trace = getenv("VILS_TRACE") != nullptr;
  const auto t_begin = std::chrono::steady_clock::now();
  auto now_ms = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
  std::vector<double> t_ready, t_issued;
  std::vector<cudaEvent_t> tev;                              // VILS_TRACE only: device timeline of every chunk (copied / solved / read back)
  if (trace) { tev.resize(3 * nchunks + 1); for (auto& v : tev) cudaEventCreate(&v); cudaEventRecord(tev[3 * nchunks], ba->stream); }
  for (int c = 0; c < nchunks; c++) {
    const int c0 = bounds[c], cn = bounds[c + 1] - bounds[c];
    if (pool) {
      pool->wait_chunk(c, cn, S);
      if (trace) t_ready.push_back(now_ms());
      if (pool->err_code.load()) { err = pool->err_code.load(); break; }
    } else if (ws) {
      err = pack_window(ba, 0, ws, S, err_msg);
      if (err) break;
    }
    cudaStream_t s = ba->pipe[c % NS];
    size_t width = 0; for (int k = c0; k < c0 + cn; k++) width = std::max(width, (size_t)ba->meta[k].bytes);
    h2d += width * cn;
    // pitched copy of the USED part of every blob.  (One contiguous copy of the chunk, slack included, was measured: 25 % more bytes made the
    // pre-packed pipeline 48 % slower, 6.0 -> 8.9 ms per 592 windows: the pipeline is bound by the solve kernel and by host memory traffic, not by
    // the PCIe rate of a single copy.)
    cudaStream_t sc = own_copy_stream ? ba->copy_stream : s;
    cudaMemcpy2DAsync(ba->d_blob + (size_t)c0 * ba->blob_stride, ba->blob_stride, ba->h_blob + (size_t)c0 * ba->blob_stride, ba->blob_stride, width, cn,
                      cudaMemcpyHostToDevice, sc);
    if (trace) cudaEventRecord(tev[3 * c], sc);
    if (own_copy_stream) { cudaEventRecord(ba->ev_chunk[c], sc); cudaStreamWaitEvent(s, ba->ev_chunk[c], 0); }
    P.slot0 = c0; P.do_prep = 1;
    launch_solve(ba, P, cn, s, cn == n);               // clusters only when the whole call is one chunk (no concurrent chunk launches)
    if (le == cudaSuccess) le = cudaGetLastError();
    if (trace) cudaEventRecord(tev[3 * c + 1], s);
    launches += 1;
    cudaMemcpyAsync(ba->h_xout + (size_t)c0 * ba->xstride, ba->d_xout + (size_t)c0 * ba->xstride, (size_t)ba->xstride * 8 * cn, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(ba->h_sum + c0, ba->d_sum + c0, sizeof(vils_summary) * cn, cudaMemcpyDeviceToHost, s);
    if (trace) { cudaEventRecord(tev[3 * c + 2], s); t_issued.push_back(now_ms()); }
  }
  cudaError_t e = cudaSuccess;
  for (int k = 0; k < NS; k++) { const cudaError_t ek = cudaStreamSynchronize(ba->pipe[k]); if (e == cudaSuccess) e = ek; }
  if (own_copy_stream) { const cudaError_t ek = cudaStreamSynchronize(ba->copy_stream); if (e == cudaSuccess) e = ek; }
  if (trace) {
    fprintf(stderr, "[VILS_TRACE] %s n=%d: ready/issued (ms):", who, n);
    for (size_t k = 0; k < t_issued.size(); k++) fprintf(stderr, " %.2f/%.2f", k < t_ready.size() ? t_ready[k] : 0.0, t_issued[k]);
    fprintf(stderr, " | all streams done %.2f\n", now_ms());
    fprintf(stderr, "[VILS_TRACE]   device copied/solved/read back (ms after the call's fork event):");
    for (int c = 0; c < nchunks; c++) {
      float a = 0, b = 0, d = 0;
      cudaEventElapsedTime(&a, tev[3 * nchunks], tev[3 * c]); cudaEventElapsedTime(&b, tev[3 * nchunks], tev[3 * c + 1]); cudaEventElapsedTime(&d, tev[3 * nchunks], tev[3 * c + 2]);
      fprintf(stderr, " %.2f/%.2f/%.2f", a, b, d);
    }
    fprintf(stderr, "\n");
    for (auto& v : tev) cudaEventDestroy(v);
  }
  if (pool) {   // drain the job so that the pool is idle again (after an error the remaining windows are skipped, not packed)
    for (int c = 0; c < nchunks; c++) pool->wait_chunk(c, bounds[c + 1] - bounds[c], S);
    if (pool->err_code.load()) { err = pool->err_code.load(); err_msg = pool->err_msg; }
  }
  if (err) return vils::fail(err, err_msg);
  if (le != cudaSuccess) return vils::fail_cuda(le, "solve_kernel launch");
  if (e != cudaSuccess) return vils::fail_cuda(e, who);
  for (int k = 0; k < n; k++) ba->meta[k].on_device = true;
  ba->last_h2d = h2d; ba->last_d2h = (size_t)ba->xstride * 8 * n + sizeof(vils_summary) * n;
  ba->last_launches = launches;
  return VILS_OK;
}

int vils_ba_solve(vils_ba* ba, int32_t n, const vils_solve_opts* opts) { return solve_pipeline(ba, n, nullptr, opts, "vils_ba_solve"); }

// vector2double() + AddResidualBlock(...) + ceres::Solve in one call (estimator.cpp:1169-1414) for n independent windows given as
// the caller's own arrays: pack (all host threads) | H2D | solve | D2H, pipelined chunk by chunk.  Window k lands in slot k.
int vils_ba_solve_windows(vils_ba* ba, int32_t n, const vils_window* ws, const vils_solve_opts* opts) {
  if (!ws) return vils::fail(VILS_ERR_BAD_ARG, "vils_ba_solve_windows: null windows");
  return solve_pipeline(ba, n, ws, opts, "vils_ba_solve_windows");
}

int vils_ba_get_state(vils_ba* ba, int32_t slot, double* pose, double* sb, double* ex, double* lam, double* td, vils_summary* sum) {
  if (!ba || slot < 0 || slot >= ba->max_windows || !ba->meta[slot].set) return vils::fail(VILS_ERR_BAD_ARG, "vils_ba_get_state: bad slot");
  const int N = ba->meta[slot].n_kf, M = ba->meta[slot].n_feat;
  const double* x = ba->h_xout + (size_t)slot * ba->xstride;
  if (pose) std::memcpy(pose, x, sizeof(double) * 7 * N);
  if (sb) std::memcpy(sb, x + 7 * N, sizeof(double) * 9 * N);
  if (ex) std::memcpy(ex, x + 16 * N, sizeof(double) * 7);
  if (td) *td = x[16 * N + 7];
  if (lam && M) std::memcpy(lam, x + 16 * N + 8, sizeof(double) * M);
  if (sum) *sum = ba->h_sum[slot];
  return ba->h_sum[slot].status;
}

int vils_ba_put_state(vils_ba* ba, int32_t slot, const double* pose, const double* sb, const double* ex, const double* lam, double td) {
  if (!ba || slot < 0 || slot >= ba->max_windows || !ba->meta[slot].set || !pose || !sb || !ex) return vils::fail(VILS_ERR_BAD_ARG, "vils_ba_put_state: bad argument");
  { const int sd = check_on_device(ba, slot, 1, "vils_ba_put_state"); if (sd) return sd; }
  if (cudaSetDevice(ba->cfg.device) != cudaSuccess) return vils::fail(VILS_ERR_CUDA, "cudaSetDevice");
  const int N = ba->meta[slot].n_kf, M = ba->meta[slot].n_feat;
  if (M && !lam) return vils::fail(VILS_ERR_BAD_ARG, "vils_ba_put_state: null inv_depth");
  double* x = ba->h_xout + (size_t)slot * ba->xstride;
  std::memcpy(x, pose, sizeof(double) * 7 * N); std::memcpy(x + 7 * N, sb, sizeof(double) * 9 * N);
  std::memcpy(x + 16 * N, ex, sizeof(double) * 7); x[16 * N + 7] = td;
  if (M) std::memcpy(x + 16 * N + 8, lam, sizeof(double) * M);
  cudaError_t e = cudaMemcpyAsync(ba->d_xout + (size_t)slot * ba->xstride, x, sizeof(double) * (16 * N + 8 + M), cudaMemcpyHostToDevice, ba->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ba->stream);
  return e == cudaSuccess ? VILS_OK : vils::fail_cuda(e, "vils_ba_put_state");
}

static int ensure_eval_buffers(vils_ba* ba) {
  if (ba->d_er) return VILS_OK;
  const vils_config& c = ba->cfg; const int N = c.max_kf;
  ba->er_stride = 15 * N + 2 * (int64_t)c.max_proj + 3 * (int64_t)c.max_lidar + 3 * 32 + 15 * N + 7;
  ba->eJ_stride = 450 * (int64_t)N + 40 * (int64_t)c.max_proj + 18 * (int64_t)c.max_lidar + 72 * 16 + 36 * 16;
  cudaError_t e = cudaMalloc(&ba->d_er, (size_t)ba->er_stride * 8 * ba->max_windows);
  if (e == cudaSuccess) e = cudaMalloc(&ba->d_eJ, (size_t)ba->eJ_stride * 8 * ba->max_windows);
  return e == cudaSuccess ? VILS_OK : vils::fail_cuda(e, "eval buffers");
}

static int launch_eval(vils_ba* ba, int slot0, int n, int apply_loss) {
  int st = ensure_eval_buffers(ba); if (st) return st;
  int np = 0, npl = 0, ned = 0, nimu = 0, nprior = 0, cons = 0;
  for (int k = slot0; k < slot0 + n; k++) {
    const WinHdr* h = reinterpret_cast<const WinHdr*>(ba->h_blob + (size_t)k * ba->blob_stride);
    np = std::max(np, h->n_proj); npl = std::max(npl, h->n_plane); ned = std::max(ned, h->n_edge);
    nimu = std::max(nimu, h->n_imu); nprior = std::max(nprior, h->prior_n); cons = std::max(cons, h->n_icp + h->n_lps);
  }
  EvalParams Q{}; Q.S = make_params(ba, nullptr); Q.S.slot0 = slot0;
  Q.r_out = ba->d_er; Q.J_out = ba->d_eJ; Q.r_stride = ba->er_stride; Q.J_stride = ba->eJ_stride; Q.apply_loss = apply_loss;
  const int xs_doubles = (16 * ba->cfg.max_kf + 8 + ba->cfg.max_feat + 1) & ~1, feat_ints = (ba->cfg.max_feat + 3) & ~3;
  const size_t proj_smem = (size_t)(EVP_T * EV_PLD + 14 * EVP_T + 2 * xs_doubles) * 8 + (size_t)(3 * EVP_T + 2 * feat_ints) * 4;
  static const int minb = getenv("VILS_EV_MINB") ? atoi(getenv("VILS_EV_MINB")) : 3;
  ensure_prepped(ba, slot0, n, ba->stream);
  cudaEventRecord(ba->ev0, ba->stream);
  int launches = 0;
  // Two streams: the latency-bound IMU warps and then the persistent projection kernel on one, the streaming LiDAR / prior /
  // constraint kernels on the other.
  cudaEventRecord(ba->ev_fork, ba->stream); cudaStreamWaitEvent(ba->stream2, ba->ev_fork, 0);
  // VILS_EV_LAYOUT (experiments): 0 = IMU ahead of the projection kernel on stream 1 (default); 1 = IMU first on stream 2, next to the
  // projection kernel; 2 = IMU last on stream 2
  static const int layout = getenv("VILS_EV_LAYOUT") ? atoi(getenv("VILS_EV_LAYOUT")) : 0;
  auto launch_imu = [&](cudaStream_t st) {
    if (!(nimu || nprior)) return;
    const int ib = (nimu * n + EVI_WARPS - 1) / EVI_WARPS, pbw = (nprior + 32 * EVI_WARPS - 1) / (32 * EVI_WARPS);
    eval_imu_kernel<<<ib + pbw * n, 32 * EVI_WARPS, 0, st>>>(Q, std::max(nimu, 1), n, ib, std::max(pbw, 1)); launches++;
  };
  if (layout == 0) launch_imu(ba->stream);
  const int pb = (np + EVP_T - 1) / EVP_T;
  if (pb) {
    const int items = pb * n, g = std::min(items, ba->n_sm * minb);
    if (minb == 2) eval_proj_kernel<2><<<g, EVP_T, proj_smem, ba->stream>>>(Q, items, pb, xs_doubles, feat_ints);
    else if (minb == 4) eval_proj_kernel<4><<<g, EVP_T, proj_smem, ba->stream>>>(Q, items, pb, xs_doubles, feat_ints);
    else eval_proj_kernel<3><<<g, EVP_T, proj_smem, ba->stream>>>(Q, items, pb, xs_doubles, feat_ints);
    launches++;
  }
  if (layout == 1) launch_imu(ba->stream2);
  const int pc = (npl + EV_T - 1) / EV_T, ec = (ned + EV_T - 1) / EV_T;
  if (pc) { eval_lidar_kernel<false><<<dim3(pc, n), EV_T, EV_T * EV_LLD * 8, ba->stream2>>>(Q); launches++; }
  if (ec) { eval_lidar_kernel<true><<<dim3(ec, n), EV_T, EV_T * EV_ELD * 8, ba->stream2>>>(Q); launches++; }
  if (cons) { eval_cons_kernel<<<dim3(1, n), 32, 0, ba->stream2>>>(Q); launches++; }
  if (layout == 2) launch_imu(ba->stream2);
  const cudaError_t le = cudaGetLastError();
  cudaEventRecord(ba->ev_join, ba->stream2);
  cudaStreamWaitEvent(ba->stream, ba->ev_join, 0);
  cudaEventRecord(ba->ev1, ba->stream);
  cudaError_t e = cudaStreamSynchronize(ba->stream);
  if (le != cudaSuccess) return vils::fail_cuda(le, "eval kernel launch");
  if (e != cudaSuccess) return vils::fail_cuda(e, "eval kernels");
  cudaEventElapsedTime(&ba->last_ms, ba->ev0, ba->ev1);
  ba->last_launches = launches;
  return VILS_OK;
}

int vils_ba_evaluate_device(vils_ba* ba, int32_t n, int32_t apply_loss) {
  int st = check_n(ba, n, "vils_ba_evaluate_device"); if (st) return st;
  st = check_on_device(ba, 0, n, "vils_ba_evaluate_device"); if (st) return st;
  return launch_eval(ba, 0, n, apply_loss);
}

int vils_ba_evaluate(vils_ba* ba, int32_t slot, int32_t apply_loss, double* residuals, double* jacobians) {
  if (!ba || slot < 0 || slot >= ba->max_windows || !ba->meta[slot].set) return vils::fail(VILS_ERR_BAD_ARG, "vils_ba_evaluate: bad slot");
  { const int sd = check_on_device(ba, slot, 1, "vils_ba_evaluate"); if (sd) return sd; }
  if (cudaSetDevice(ba->cfg.device) != cudaSuccess) return vils::fail(VILS_ERR_CUDA, "cudaSetDevice");
  int st = launch_eval(ba, slot, 1, apply_loss); if (st) return st;
  const SlotMeta& m = ba->meta[slot];
  // The kernels write each family in the library's sorted order; hand the caller its own factor order back.
  std::vector<double> hr(m.n_res), hj((size_t)std::max<int64_t>(m.n_jac, 1));
  cudaError_t e = cudaMemcpy(hr.data(), ba->d_er + (size_t)slot * ba->er_stride, sizeof(double) * m.n_res, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && m.n_jac) e = cudaMemcpy(hj.data(), ba->d_eJ + (size_t)slot * ba->eJ_stride, sizeof(double) * m.n_jac, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return vils::fail_cuda(e, "evaluate copy");
  const uint8_t* base = ba->h_blob + (size_t)slot * ba->blob_stride;
  const WinHdr* h = reinterpret_cast<const WinHdr*>(base);
  auto unperm = [&](int n, const int32_t* orig, int nr, int nj, size_t ro, size_t jo) {
    std::vector<double> tr((size_t)n * nr), tj((size_t)n * nj);
    for (int s2 = 0; s2 < n; s2++) {
      std::copy(hr.begin() + ro + (size_t)s2 * nr, hr.begin() + ro + (size_t)(s2 + 1) * nr, tr.begin() + (size_t)orig[s2] * nr);
      std::copy(hj.begin() + jo + (size_t)s2 * nj, hj.begin() + jo + (size_t)(s2 + 1) * nj, tj.begin() + (size_t)orig[s2] * nj);
    }
    std::copy(tr.begin(), tr.end(), hr.begin() + ro); std::copy(tj.begin(), tj.end(), hj.begin() + jo);
  };
  size_t ro = (size_t)15 * h->n_imu, jo = (size_t)450 * h->n_imu;
  unperm(h->n_proj, reinterpret_cast<const int32_t*>(base + h->off[OFF_PROJ_IDX]) + 3 * h->n_proj, 2, 40, ro, jo);
  ro += (size_t)2 * h->n_proj; jo += (size_t)40 * h->n_proj;
  unperm(h->n_plane, reinterpret_cast<const int32_t*>(base + h->off[OFF_PLANE_IDX]) + h->n_plane, 1, 6, ro, jo);
  ro += h->n_plane; jo += (size_t)6 * h->n_plane;
  unperm(h->n_edge, reinterpret_cast<const int32_t*>(base + h->off[OFF_EDGE_IDX]) + h->n_edge, 3, 18, ro, jo);
  if (residuals) std::copy(hr.begin(), hr.end(), residuals);
  if (jacobians && m.n_jac) std::copy(hj.begin(), hj.begin() + m.n_jac, jacobians);
  return VILS_OK;
}

int vils_ba_linearize(vils_ba* ba, int32_t slot, double* S, double* g, double* cost) {
  if (!ba || slot < 0 || slot >= ba->max_windows || !ba->meta[slot].set) return vils::fail(VILS_ERR_BAD_ARG, "vils_ba_linearize: bad slot");
  { const int sd = check_on_device(ba, slot, 1, "vils_ba_linearize"); if (sd) return sd; }
  if (cudaSetDevice(ba->cfg.device) != cudaSuccess) return vils::fail(VILS_ERR_CUDA, "cudaSetDevice");
  vils_solve_opts o; vils_default_solve_opts(&o); o.mu = 0;
  SolveParams P = make_params(ba, &o); P.slot0 = slot; P.lin_out = ba->d_lin;
  launch_solve(ba, P, 1, ba->stream);
  VILS_LAUNCH_CHECK("solve_kernel (linearize) launch");
  const int D = 15 * ba->meta[slot].n_kf + 7;
  cudaMemcpyAsync(ba->h_lin, ba->d_lin, ((size_t)D * D + D + 1) * 8, cudaMemcpyDeviceToHost, ba->stream);
  cudaError_t e = cudaStreamSynchronize(ba->stream);
  if (e != cudaSuccess) return vils::fail_cuda(e, "linearize");
  if (S) std::memcpy(S, ba->h_lin, sizeof(double) * D * D);
  if (g) std::memcpy(g, ba->h_lin + (size_t)D * D, sizeof(double) * D);
  if (cost) *cost = ba->h_lin[(size_t)D * D + D];
  return VILS_OK;
}


// Marginalization of one slot at its SOLVED state (the last vils_ba_solve / solve_device of that slot).
int vils_ba_marginalize(vils_ba* ba, int32_t slot, int32_t flag, vils_prior_out* out) {
  if (!ba || !out || slot < 0 || slot >= ba->max_windows || !ba->meta[slot].set || (flag != VILS_MARGIN_OLD && flag != VILS_MARGIN_SECOND_NEW))
    return vils::fail(VILS_ERR_BAD_ARG, "vils_ba_marginalize: bad argument");
  { const int sd = check_on_device(ba, slot, 1, "vils_ba_marginalize"); if (sd) return sd; }
  if (cudaSetDevice(ba->cfg.device) != cudaSuccess) return vils::fail(VILS_ERR_CUDA, "cudaSetDevice");
  const vils_config& c = ba->cfg;
  const int D = 15 * c.max_kf + 7, Tcap = D + c.max_feat;
  if (!ba->d_mws) {
    MargParams& q = ba->mq; int64_t o = 0;
    auto take = [&](int64_t n) { int64_t r = o; o += (n + 1) & ~int64_t(1); return r; };
    const int64_t T2 = (int64_t)Tcap * Tcap;
    q.oH = take(T2); q.oG = take(Tcap); q.oA = take(T2); q.oB = take(Tcap); q.oV = take(T2); q.oW = take(Tcap); q.oAinv = take(T2); q.oArm = take(T2);
    q.oAr = take((int64_t)D * D); q.oBr = take(Tcap); q.oV2 = take(T2); q.oS = take(Tcap);
    q.oStage = take(std::max<int64_t>((int64_t)c.max_proj * 44, 512)); q.oJout = take((int64_t)D * D); q.oRout = take(D); q.oX0 = take((2 * c.max_kf + 2) * 9);
    ba->mws_doubles = o; ba->mi_ints = 8 + 5 * Tcap + c.max_feat + 2 * (2 * c.max_kf + 2);
    q.Tcap = Tcap; q.Mcap = c.max_feat;
    cudaError_t e = cudaMalloc(&ba->d_mws, (size_t)o * 8);
    if (e == cudaSuccess) e = cudaMalloc(&ba->d_miws, sizeof(int32_t) * ba->mi_ints);
    if (e != cudaSuccess) return vils::fail_cuda(e, "marginalize workspace");
    q.ws = ba->d_mws; q.iws = ba->d_miws;
  }
  MargParams q = ba->mq; q.slot = slot; q.flag = flag;
  SolveParams P = make_params(ba, nullptr);
  ensure_prepped(ba, slot, 1, ba->stream);
  cudaMemsetAsync(ba->d_miws, 0, sizeof(int32_t) * 8, ba->stream);
  margin_kernel<<<1, SOLVE_THREADS, 16384, ba->stream>>>(P, q);
  VILS_LAUNCH_CHECK("margin_kernel launch");
  int32_t hdr[8];
  cudaMemcpyAsync(hdr, ba->d_miws, sizeof(hdr), cudaMemcpyDeviceToHost, ba->stream);
  cudaError_t e = cudaStreamSynchronize(ba->stream);
  if (e != cudaSuccess) return vils::fail_cuda(e, "margin_kernel");
  const int n = hdr[0], m = hdr[1], nb = hdr[2], nx0 = hdr[4];
  out->n = n; out->m = m; out->nblk = nb;
  if (n == 0) return VILS_OK;
  if (n > out->capacity_n || !out->J || !out->r || !out->blk || !out->x0) return vils::fail(VILS_ERR_CAPACITY, "vils_ba_marginalize: output capacity too small");
  std::vector<int32_t> blk(2 * nb);
  const int N = ba->meta[slot].n_kf; (void)N;
  e = cudaMemcpy(out->J, ba->d_mws + q.oJout, sizeof(double) * n * n, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(out->r, ba->d_mws + q.oRout, sizeof(double) * n, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(out->x0, ba->d_mws + q.oX0, sizeof(double) * nx0, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(blk.data(), ba->d_miws + 8 + 3 * Tcap + c.max_feat, sizeof(int32_t) * 2 * nb, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return vils::fail_cuda(e, "marginalize copy");
  for (int b = 0; b < nb; b++) out->blk[b] = blk[2 * b];     // ids already re-addressed to the slid window
  return VILS_OK;
}


// ---- factor-sharded mode ---------------------------------------------------------------------------------------------
static int ensure_shard_buffer(vils_ba* ba) {
  if (ba->d_shard) return VILS_OK;
  const int D = 15 * ba->cfg.max_kf + 7;
  ba->shard_doubles = (size_t)D * D + 2 * (size_t)D + 1;
  cudaError_t e = cudaMalloc(&ba->d_shard, ba->shard_doubles * 8);
  return e == cudaSuccess ? VILS_OK : vils::fail_cuda(e, "sharded buffer");
}
int vils_ba_sharded_buffer(vils_ba* ba, void** dev_ptr, size_t* n_doubles) {
  if (!ba || !dev_ptr || !n_doubles || !ba->meta[0].set) return vils::fail(VILS_ERR_BAD_ARG, "vils_ba_sharded_buffer: stage the window in slot 0 first");
  if (cudaSetDevice(ba->cfg.device) != cudaSuccess) return vils::fail(VILS_ERR_CUDA, "cudaSetDevice");
  int st = ensure_shard_buffer(ba); if (st) return st;
  const int D = 15 * ba->meta[0].n_kf + 7;
  *dev_ptr = ba->d_shard; *n_doubles = (size_t)D * D + 2 * (size_t)D + 1;
  return VILS_OK;
}
int vils_ba_sharded_linearize(vils_ba* ba, int32_t iteration, const vils_solve_opts* opts) {
  if (!ba || !opts || iteration < 0 || !ba->meta[0].set) return vils::fail(VILS_ERR_BAD_ARG, "vils_ba_sharded_linearize: bad argument");
  { const int sd = check_on_device(ba, 0, 1, "vils_ba_sharded_linearize"); if (sd) return sd; }
  if (cudaSetDevice(ba->cfg.device) != cudaSuccess) return vils::fail(VILS_ERR_CUDA, "cudaSetDevice");
  int st = ensure_shard_buffer(ba); if (st) return st;
  SolveParams P = make_params(ba, opts);
  ensure_prepped(ba, 0, 1, ba->stream);
  if (iteration == 0) cudaMemsetAsync(ba->d_sum, 0, sizeof(vils_summary), ba->stream);
  cudaEventRecord(ba->ev0, ba->stream);
  if (ba->h_in_smem && ba->hv_in_smem) shard_lin_kernel<true><<<1, SOLVE_THREADS, ba->smem_bytes, ba->stream>>>(P, ba->d_shard, iteration == 0, opts->mu);
  else shard_lin_kernel<false><<<1, SOLVE_THREADS, ba->smem_bytes, ba->stream>>>(P, ba->d_shard, iteration == 0, opts->mu);
  VILS_LAUNCH_CHECK("shard_lin_kernel launch");
  cudaEventRecord(ba->ev1, ba->stream);
  cudaError_t e = cudaStreamSynchronize(ba->stream);      // the caller's all-reduce runs on its own stream: hand over a finished buffer
  if (e != cudaSuccess) return vils::fail_cuda(e, "shard_lin_kernel");
  cudaEventElapsedTime(&ba->last_ms, ba->ev0, ba->ev1);
  return VILS_OK;
}
int vils_ba_sharded_update(vils_ba* ba, const vils_solve_opts* opts) {
  if (!ba || !opts || !ba->d_shard) return vils::fail(VILS_ERR_BAD_ARG, "vils_ba_sharded_update: linearize first");
  if (cudaSetDevice(ba->cfg.device) != cudaSuccess) return vils::fail(VILS_ERR_CUDA, "cudaSetDevice");
  SolveParams P = make_params(ba, opts);
  cudaEventRecord(ba->ev0, ba->stream);
  if (ba->h_in_smem && ba->hv_in_smem) shard_upd_kernel<true><<<1, SOLVE_THREADS, ba->smem_bytes, ba->stream>>>(P, ba->d_shard, opts->mu);
  else shard_upd_kernel<false><<<1, SOLVE_THREADS, ba->smem_bytes, ba->stream>>>(P, ba->d_shard, opts->mu);
  VILS_LAUNCH_CHECK("shard_upd_kernel launch");
  cudaEventRecord(ba->ev1, ba->stream);
  cudaError_t e = cudaStreamSynchronize(ba->stream);
  if (e != cudaSuccess) return vils::fail_cuda(e, "shard_upd_kernel");
  cudaEventElapsedTime(&ba->last_ms, ba->ev0, ba->ev1);
  return VILS_OK;
}
// Host-mediated access to the partial system (tests, or transports other than NCCL).
int vils_ba_sharded_read(vils_ba* ba, double* host) {
  if (!ba || !host || !ba->d_shard) return vils::fail(VILS_ERR_BAD_ARG, "vils_ba_sharded_read");
  const int D = 15 * ba->meta[0].n_kf + 7;
  cudaError_t e = cudaMemcpy(host, ba->d_shard, ((size_t)D * D + 2 * D + 1) * 8, cudaMemcpyDeviceToHost);
  return e == cudaSuccess ? VILS_OK : vils::fail_cuda(e, "sharded read");
}
int vils_ba_sharded_write(vils_ba* ba, const double* host) {
  if (!ba || !host || !ba->d_shard) return vils::fail(VILS_ERR_BAD_ARG, "vils_ba_sharded_write");
  const int D = 15 * ba->meta[0].n_kf + 7;
  cudaError_t e = cudaMemcpy(ba->d_shard, host, ((size_t)D * D + 2 * D + 1) * 8, cudaMemcpyHostToDevice);
  return e == cudaSuccess ? VILS_OK : vils::fail_cuda(e, "sharded write");
}

// ---- native factor-sharded solve over NCCL ----------------------------------------------------------------------------------
// libnccl is bound lazily (dlopen) the first time the sharded API is used: a process that already carries an NCCL (torch's bundled copy)
// shares it, everything else in the library works on a box without NCCL.
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
static NcclApi* nccl_api() {
  static NcclApi api; static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) return;
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
    api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy && api.GetErrorString;
  });
  return api.ok ? &api : nullptr;
}
static void nccl_comm_destroy(ncclComm_t c) { if (NcclApi* a = nccl_api()) a->CommDestroy(c); }
static int fail_nccl(ncclResult_t r, const char* what) { NcclApi* a = nccl_api(); return vils::fail(VILS_ERR_CUDA, std::string(what) + ": " + (a ? a->GetErrorString(r) : "NCCL unavailable")); }

int vils_nccl_unique_id(uint8_t id[128]) {
  NcclApi* a = nccl_api();
  if (!a) return vils::fail(VILS_ERR_NO_DEVICE, "vils_nccl_unique_id: libnccl.so.2 not found");
  if (!id) return vils::fail(VILS_ERR_BAD_ARG, "vils_nccl_unique_id: null");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId u; const ncclResult_t r = a->GetUniqueId(&u);
  if (r != ncclSuccess) return fail_nccl(r, "ncclGetUniqueId");
  std::memcpy(id, &u, 128);
  return VILS_OK;
}

int vils_ba_sharded_init(vils_ba* ba, int32_t rank, int32_t nranks, const uint8_t id[128]) {
  NcclApi* a = nccl_api();
  if (!a) return vils::fail(VILS_ERR_NO_DEVICE, "vils_ba_sharded_init: libnccl.so.2 not found");
  if (!ba || !id || nranks < 1 || rank < 0 || rank >= nranks) return vils::fail(VILS_ERR_BAD_ARG, "vils_ba_sharded_init: bad argument");
  if (cudaSetDevice(ba->cfg.device) != cudaSuccess) return vils::fail(VILS_ERR_CUDA, "cudaSetDevice");
  if (ba->comm) { a->CommDestroy(ba->comm); ba->comm = nullptr; }
  ncclUniqueId u; std::memcpy(&u, id, 128);
  const ncclResult_t r = a->CommInitRank(&ba->comm, nranks, u, rank);
  if (r != ncclSuccess) { ba->comm = nullptr; return fail_nccl(r, "ncclCommInitRank"); }
  ba->comm_rank = rank; ba->comm_size = nranks;
  return VILS_OK;
}

// out[0..M) = inverse depth of the landmarks THIS rank holds projection factors for (0 elsewhere), out[M..2M) = 1 / 0 ownership.
__global__ void lam_pack_kernel(SolveParams P, double* out) {
  const Win W = decode(P, P.slot0);
  const double* x = P.xout + (size_t)P.slot0 * P.xout_stride;
  const int32_t* lm_feat = W.i(OFF_LM_FEAT);
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < 2 * W.M; f += gridDim.x * blockDim.x) out[f] = 0.0;
  __syncthreads();     // one block
  for (int r = threadIdx.x; r < W.h->n_lm; r += blockDim.x) { const int f = lm_feat[r]; out[f] = x[XL(W.N) + f]; out[W.M + f] = 1.0; }
}
// after the sum over ranks: landmarks somebody owns take the owner's value, the others keep the local (= initial) one
__global__ void lam_unpack_kernel(SolveParams P, const double* in) {
  const Win W = decode(P, P.slot0);
  double* x = P.xout + (size_t)P.slot0 * P.xout_stride;
  for (int f = threadIdx.x; f < W.M; f += blockDim.x) if (in[W.M + f] > 0.0) x[XL(W.N) + f] = in[f] / in[W.M + f];
}

// All Gauss-Newton iterations of the factor-sharded solve of slot 0, enqueued on the handle's stream without a host synchronisation in
// between: shard_lin_kernel -> ncclAllReduce (in place, D^2 + 2D + 1 doubles, NVLink / NVSwitch) -> shard_upd_kernel, max_iters times; then
// the inverse depths (each landmark lives on exactly one rank) are exchanged with one more all-reduce, so every rank ends with the full state.
int vils_ba_sharded_solve(vils_ba* ba, const vils_solve_opts* opts, vils_summary* summary) {
  NcclApi* a = nccl_api();
  if (!ba || !opts || !ba->meta[0].set) return vils::fail(VILS_ERR_BAD_ARG, "vils_ba_sharded_solve: stage the window in slot 0 first");
  if (!a || !ba->comm) return vils::fail(VILS_ERR_BAD_ARG, "vils_ba_sharded_solve: call vils_ba_sharded_init first");
  if (opts->mode != VILS_MODE_GN || opts->max_iters <= 0) return vils::fail(VILS_ERR_BAD_ARG, "vils_ba_sharded_solve: Gauss-Newton mode only");
  { const int sd = check_on_device(ba, 0, 1, "vils_ba_sharded_solve"); if (sd) return sd; }
  if (cudaSetDevice(ba->cfg.device) != cudaSuccess) return vils::fail(VILS_ERR_CUDA, "cudaSetDevice");
  int st = ensure_shard_buffer(ba); if (st) return st;
  if (!ba->d_lamred) { const cudaError_t e = cudaMalloc(&ba->d_lamred, sizeof(double) * 2 * std::max(1, ba->cfg.max_feat)); if (e != cudaSuccess) return vils::fail_cuda(e, "sharded depth buffer"); }
  SolveParams P = make_params(ba, opts);
  const bool both = ba->h_in_smem && ba->hv_in_smem;
  const int D = 15 * ba->meta[0].n_kf + 7, M = ba->meta[0].n_feat;
  const size_t cnt = (size_t)D * D + 2 * (size_t)D + 1;
  cudaStream_t s = ba->stream;
  ensure_prepped(ba, 0, 1, s);
  cudaMemsetAsync(ba->d_sum, 0, sizeof(vils_summary), s);
  cudaEventRecord(ba->ev0, s);
  cudaError_t le = cudaSuccess; ncclResult_t nr = ncclSuccess;
  for (int it = 0; it < opts->max_iters && nr == ncclSuccess; it++) {
    if (both) shard_lin_kernel<true><<<1, SOLVE_THREADS, ba->smem_bytes, s>>>(P, ba->d_shard, it == 0, opts->mu);
    else shard_lin_kernel<false><<<1, SOLVE_THREADS, ba->smem_bytes, s>>>(P, ba->d_shard, it == 0, opts->mu);
    if (le == cudaSuccess) le = cudaGetLastError();
    nr = a->AllReduce(ba->d_shard, ba->d_shard, cnt, ncclDouble, ncclSum, ba->comm, s);
    if (both) shard_upd_kernel<true><<<1, SOLVE_THREADS, ba->smem_bytes, s>>>(P, ba->d_shard, opts->mu);
    else shard_upd_kernel<false><<<1, SOLVE_THREADS, ba->smem_bytes, s>>>(P, ba->d_shard, opts->mu);
    if (le == cudaSuccess) le = cudaGetLastError();
  }
  if (nr == ncclSuccess && M > 0) {
    lam_pack_kernel<<<1, 256, 0, s>>>(P, ba->d_lamred);
    nr = a->AllReduce(ba->d_lamred, ba->d_lamred, 2 * (size_t)M, ncclDouble, ncclSum, ba->comm, s);
    lam_unpack_kernel<<<1, 256, 0, s>>>(P, ba->d_lamred);
    if (le == cudaSuccess) le = cudaGetLastError();
  }
  cudaEventRecord(ba->ev1, s);
  cudaMemcpyAsync(ba->h_xout, ba->d_xout, (size_t)ba->xstride * 8, cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(ba->h_sum, ba->d_sum, sizeof(vils_summary), cudaMemcpyDeviceToHost, s);
  const cudaError_t e = cudaStreamSynchronize(s);               // the only host synchronisation of the call
  if (nr != ncclSuccess) return fail_nccl(nr, "ncclAllReduce");
  if (le != cudaSuccess) return vils::fail_cuda(le, "sharded kernel launch");
  if (e != cudaSuccess) return vils::fail_cuda(e, "vils_ba_sharded_solve");
  cudaEventElapsedTime(&ba->last_ms, ba->ev0, ba->ev1);
  ba->last_launches = 2 * opts->max_iters + 2;
  if (summary) *summary = ba->h_sum[0];
  return ba->h_sum[0].status;
}

int vils_ba_set_cluster(vils_ba* ba, int32_t cluster_size) {
  if (!ba || !(cluster_size == 0 || cluster_size == 1 || cluster_size == 2 || cluster_size == 4 || cluster_size == 8 || cluster_size == 16)) return vils::fail(VILS_ERR_BAD_ARG, "vils_ba_set_cluster: 0 (auto), 1 (off), 2, 4, 8 or 16");
  ba->cluster_pref = cluster_size; return VILS_OK;
}
int vils_ba_last_cluster(vils_ba* ba, int32_t* cluster_size) { if (!ba || !cluster_size) return VILS_ERR_BAD_ARG; *cluster_size = ba->last_cluster; return VILS_OK; }

int vils_ba_last_device_ms(vils_ba* ba, float* ms) { if (!ba || !ms) return VILS_ERR_BAD_ARG; *ms = ba->last_ms; return VILS_OK; }
int vils_ba_last_transfer_bytes(vils_ba* ba, size_t* h2d, size_t* d2h) { if (!ba) return VILS_ERR_BAD_ARG; if (h2d) *h2d = ba->last_h2d; if (d2h) *d2h = ba->last_d2h; return VILS_OK; }
int vils_ba_last_launches(vils_ba* ba, int32_t* n) { if (!ba || !n) return VILS_ERR_BAD_ARG; *n = ba->last_launches; return VILS_OK; }

// Estimator::double2vector gauge re-anchoring (estimator.cpp:962-1011): host arithmetic on 7N + 9N doubles, kept on the
// host exactly where the reference has it (it needs the pre-solve Rs[0]/Ps[0] the caller owns).
static void q2R_h(const double* q, double R[9]) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w); R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w); R[7] = 2 * (y * z + x * w); R[8] = 1 - 2 * (x * x + y * y);
}
static void R2ypr_h(const double R[9], double ypr[3]) {   // utility.h:66-81, degrees
  const double nx = R[0], ny = R[3], nz = R[6], ox = R[1], oy = R[4], ax = R[2], ay = R[5];
  const double y = atan2(ny, nx), p = atan2(-nz, nx * cos(y) + ny * sin(y)), r = atan2(ax * sin(y) - ay * cos(y), -ox * sin(y) + oy * cos(y));
  ypr[0] = y / M_PI * 180.0; ypr[1] = p / M_PI * 180.0; ypr[2] = r / M_PI * 180.0;
}
static void R2q_h(const double m[9], double q[4]) {   // Eigen Quaternion(Matrix3)
  double t = m[0] + m[4] + m[8];
  if (t > 0) { t = sqrt(t + 1.0); q[3] = 0.5 * t; t = 0.5 / t; q[0] = (m[7] - m[5]) * t; q[1] = (m[2] - m[6]) * t; q[2] = (m[3] - m[1]) * t; }
  else {
    int i = 0; if (m[4] > m[0]) i = 1; if (m[8] > m[i * 4]) i = 2; const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m[i * 4] - m[j * 4] - m[k * 4] + 1.0); q[i] = 0.5 * t; t = 0.5 / t;
    q[3] = (m[k * 3 + j] - m[j * 3 + k]) * t; q[j] = (m[j * 3 + i] + m[i * 3 + j]) * t; q[k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
  }
}
int vils_double2vector(int32_t n_kf, const double* pose0_before, double* pose, double* sb) {
  if (n_kf <= 0 || !pose0_before || !pose || !sb) return VILS_ERR_BAD_ARG;
  double R0[9], R00[9], y0[3], y00[3], rot[9];
  q2R_h(pose0_before + 3, R0); q2R_h(pose + 3, R00); R2ypr_h(R0, y0); R2ypr_h(R00, y00);
  const double yd = (y0[0] - y00[0]) / 180.0 * M_PI;
  rot[0] = cos(yd); rot[1] = -sin(yd); rot[2] = 0; rot[3] = sin(yd); rot[4] = cos(yd); rot[5] = 0; rot[6] = 0; rot[7] = 0; rot[8] = 1;
  if (fabs(fabs(y0[1]) - 90) < 1.0 || fabs(fabs(y00[1]) - 90) < 1.0)   // :979-988
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += R0[i * 3 + k] * R00[j * 3 + k]; rot[i * 3 + j] = s; }
  const double P0[3] = {pose[0], pose[1], pose[2]};
  for (int k = 0; k < n_kf; k++) {
    double* p = pose + 7 * k; double qn[4]; const double nn = sqrt(p[3] * p[3] + p[4] * p[4] + p[5] * p[5] + p[6] * p[6]);
    for (int a = 0; a < 4; a++) qn[a] = p[3 + a] / nn;
    double R[9], RR[9]; q2R_h(qn, R);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int m = 0; m < 3; m++) s += rot[i * 3 + m] * R[m * 3 + j]; RR[i * 3 + j] = s; }
    const double d[3] = {p[0] - P0[0], p[1] - P0[1], p[2] - P0[2]}; double* v = sb + 9 * k; const double v0[3] = {v[0], v[1], v[2]};
    for (int i = 0; i < 3; i++) { p[i] = rot[i * 3] * d[0] + rot[i * 3 + 1] * d[1] + rot[i * 3 + 2] * d[2] + pose0_before[i]; v[i] = rot[i * 3] * v0[0] + rot[i * 3 + 1] * v0[1] + rot[i * 3 + 2] * v0[2]; }
    R2q_h(RR, p + 3);
  }
  return VILS_OK;
}

}  // extern "C"
