// ba_device.cuh — device side of the sliding-window solve: one CTA owns one window for the whole Gauss-Newton loop.
//
// Per iteration (all in one kernel, no host round trip):
//   V  pair pass      : projection factors grouped by keyframe pair (i,j); a warp evaluates <=16 factors at a time
//                       (ProjectionTdFactor + Cauchy corrector), stages the weighted 2x19 Jacobian rows + residual
//                       in shared memory and accumulates the pair-local 20x20 block A^T A (direct terms, gradient);
//                       per-factor landmark partials (C, g_l, E rows) go to the per-window scratch in L2.
//   L  landmark reduce: one thread per landmark sums its factors' partials in a fixed order (deterministic).
//   S  Schur SYRK     : Hv = -E diag(1/C) E^T, gv = -E g_l / C over the (6N+7)-dim visual sub-system (the only dense
//                       contraction of the path), E streamed through shared memory in 32-landmark chunks.
//   G  gather         : Hv + pair blocks -> tile-packed H (16x16 tiles, lower), g, diag for damping.
//   I  IMU / LiDAR / ICP / LPS / prior contributions straight into H.
//   C  blocked Cholesky with tile inverses, forward solve fused in as an extra row, blocked back-substitution.
//   U  landmark back-substitution, PoseLocalParameterization::Plus.
// No atomics anywhere: every accumulator has exactly one owner, so results are bit-reproducible run to run.
#pragma once
#include <cuda_runtime.h>

#include "../../include/vils_cabi.h"
#include "ba_layout.h"
#include "factors.cuh"

namespace vb {

constexpr int TB = 16;            // Cholesky tile
constexpr int TLD = 17;           // padded tile row stride (doubles): lanes walking rows hit distinct shared-memory banks
constexpr int TSZ = TB * TLD + 2; // padded tile stride: neighbouring tiles start 2 doubles (4 banks) apart
#ifndef VILS_SOLVE_THREADS
#define VILS_SOLVE_THREADS 384   // 12 warps (must be a multiple of 4 warps: Cholesky look-ahead).  512 threads x 128 registers spilled ~1 KB per thread and
                                 // 31 % more DRAM traffic for +2 % speed (round 2: 4.75 -> 3.29 MB per window, 116.8 k -> 114.5 k solves/s)
#endif
constexpr int SOLVE_THREADS = VILS_SOLVE_THREADS;
constexpr int SOLVE_WARPS = SOLVE_THREADS / 32;
static_assert(SOLVE_THREADS % 128 == 0, "the Cholesky look-ahead parks one warp per SM sub-partition: whole groups of 4 warps only");
constexpr int STAGE_LD = 25;      // pair-pass staging row: 19 Jacobian cols + residual + pad (odd stride: fewer bank conflicts)
#ifndef VILS_PAIR_CHUNK
#define VILS_PAIR_CHUNK 224      // 7 warps evaluate projection factors per round, 5 are free for the IMU stages
#endif
#ifndef VILS_SOLVE_MINB
#define VILS_SOLVE_MINB 1
#endif
constexpr int PAIR_CHUNK = VILS_PAIR_CHUNK;   // projection factors evaluated per round (one per thread), staged in shared memory
static_assert(VILS_PAIR_CHUNK <= VILS_SOLVE_THREADS, "one thread per projection factor of a round");
constexpr int SOLVE_MINB = VILS_SOLVE_MINB;   // resident CTAs (windows) per SM the solve kernels are compiled for
constexpr int PAIR_LD = 20;       // pair-local block: [pose_i 6 | pose_j 6 | ex 6 | td | r]
constexpr int ECHUNK = 32;        // landmarks per Schur chunk
constexpr int PART_LD = 16;       // per-factor landmark partial: C, g_l, e_i(6), e_ex(6), e_td, pad

// offsets (doubles) into a cluster's exchange area (latency kernels, ba_cluster.cuh)
struct ClScratch { int64_t hvpart, gvpart, lidblk, gsc, hdsc, dxg, lamg, costp, flagg, xg, cinvg, glamg, crawg, sclg, cmd, total; };

struct SolveParams {
  const uint8_t* blobs; int64_t blob_stride;
  double* scratch; ScratchLayout sl;
  double* tscratch;               // != null: solve_kernel works in a per-SM scratch slot (indexed by %smid) instead of the per-window one
  double* xout; int64_t xout_stride;
  vils_summary* summary;
  vf::BaCfg cfg;
  int32_t mode, max_iters;
  double mu, lm_radius, f_tol, p_tol, min_rel_dec;
  long long time_cap_ns;          // 0 = none: stop iterating once this window's solve has run that long (max_solver_time_in_seconds)
  int32_t Ncap, Mcap;             // capacities the shared-memory carve-up is sized for
  int32_t h_in_smem, hv_in_smem;
  double* lin_out;                // != null: dump [S (D x D row-major) | g (D) | cost] of the first linearisation and stop
  int32_t slot0;                  // first slot handled by blockIdx.x == 0
  int32_t do_prep;                // solve_kernel runs prep_window itself (vils_ba_solve's pipelined path)
  long long* prof;                // optional: per-phase SM cycle counters of block 0 (VILS_PROF=1)
  ClScratch cl; double* clbuf; int32_t cl_imu_slots;   // cluster-assisted solves (solve_kernel<.., CL = true>): exchange area of the window's cluster
};

struct Smem {   // offsets in doubles into dynamic shared memory, computed identically on host and device
  int xs, xc, g, dx, hd, fx, pid, gv, hdv, cinv, glam, red, linv, hv, uni, imu, tbl, rot, scc, ddc, gc, stp, craw, scl, hdr, ptab, lmst, lmft, total;
  int ntile_rows;
};

__host__ __device__ inline int tri(int n) { return n * (n + 1) / 2; }

__host__ __device__ inline Smem smem_layout(int Ncap, int Mcap, int h_in_smem, int hv_in_smem) {
  Smem s; int o = 0;
  const int X = 16 * Ncap + 8 + Mcap, D = 15 * Ncap + 7, Dv = 6 * Ncap + 7, nb = (D + TB - 1) / TB, Dvp = (Dv + 3) & ~3;
  auto take = [&](int n) { int r = o; o += (n + 1) & ~1; return r; };
  s.xs = take(X); s.xc = take(X);
  s.g = take(nb * TB); s.dx = take(nb * TB); s.hd = take(nb * TB); s.fx = take(nb * TB / 2 + 1); s.pid = take(Ncap * Ncap / 2 + 1);
  s.gv = take(Dvp); s.hdv = take(Dvp);
  s.cinv = take(Mcap); s.glam = take(Mcap);
  // dogleg (VILS_MODE_DOGLEG): Jacobi scale and trust-region diagonal per camera dimension, unreduced camera gradient, combined step;
  // raw landmark diagonal C and landmark Jacobi scale
  s.scc = take(nb * TB); s.ddc = take(nb * TB); s.gc = take(nb * TB); s.stp = take(nb * TB); s.craw = take(Mcap); s.scl = take(Mcap);
  s.red = take(64 + SOLVE_WARPS * 2);
  // shared-memory copies of the window header and of the small index tables every phase walks (pair list, landmark CSR, landmark -> feature):
  // dependent loads of these from global memory were the critical path of the pair-accumulate and landmark phases
  s.hdr = take((int)((sizeof(WinHdr) + 7) / 8)); s.ptab = take(2 * (Ncap * (Ncap - 1) / 2) + 2); s.lmst = take(Mcap / 2 + 2); s.lmft = take(Mcap / 2 + 2);
  s.tbl = take(114);                                         // IMU Jacobian assembly table (450 x uint16)
  s.rot = take((Ncap + 1) * 9);                              // rotation matrices of the keyframes + ric (pair pass)
  s.linv = h_in_smem ? take(nb * TB * TB) : -1;             // large windows keep the tile inverses in the L2 scratch
  int uni = PAIR_CHUNK * 2 * STAGE_LD;                       // pair-pass staging
  if (ECHUNK * Dvp > uni) uni = ECHUNK * Dvp;                // Schur chunk
  if (h_in_smem && tri(nb) * TSZ > uni) uni = tri(nb) * TSZ;
  s.uni = take(uni);
  // hv and imu are adjacent on purpose: once the gather has consumed Hv, the IMU pass stages ALL factors at once in
  // [hv, imu end) (needs (Ncap-1)*466 doubles; Dvp^2 + ceil((Ncap+1)/2)*466 always covers it)
  s.hv = hv_in_smem ? take(Dvp * Dvp) : -1;
  s.imu = take(((Ncap + 1) / 2) * 466);
  s.total = o; s.ntile_rows = nb;
  return s;
}

// Asynchronous global -> shared copy of one double (LDGSTS): a staging loop written as `smem[k] = gmem[k]` is one L2 round trip per element
// and thread (the store waits for its load and the next load is issued after the store), a loop of these has every element in flight at once.
__device__ __forceinline__ void cp_async_f64(double* dst_shared, const double* src_global) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((uint32_t)__cvta_generic_to_shared(dst_shared)), "l"(src_global) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }

// Clock read for the in-kernel phase profile.  BAR.SYNC blocks lazily (the warp keeps issuing independent instructions, a plain clock read
// among them, until it needs the barrier), which booked barrier waits on the phase AFTER the barrier; a shared-memory load cannot pass the
// barrier and the clock read is issued after it.
__device__ __forceinline__ long long prof_clock() {
  unsigned dummy; long long c;
  asm volatile("ld.shared.u32 %0, [%2];\n\tmov.u64 %1, %%clock64;" : "=r"(dummy), "=l"(c) : "r"(0u) : "memory");
  return c + (dummy & 0u);
}

// ---- tile-packed lower-triangular matrix -----------------------------------------------------------------------
__device__ __forceinline__ int tidx(int i, int j) {  // requires i >= j
  const int bi = i >> 4, bj = j >> 4;
  return (bi * (bi + 1) / 2 + bj) * TSZ + (i & 15) * TLD + (j & 15);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct Win {   // decoded blob
  const uint8_t* base; const WinHdr* h;
  int N, M, D, Dv, Dvp, nb;
  const int32_t* pairs_s = nullptr; const int32_t* lmst_s = nullptr; const int32_t* lmft_s = nullptr;   // shared-memory copies (stage_tables), else the blob's
  __device__ const int32_t* pairs() const { return pairs_s ? pairs_s : i(OFF_PAIR); }
  __device__ const int32_t* lm_start() const { return lmst_s ? lmst_s : i(OFF_LM_START); }
  __device__ const int32_t* lm_feat() const { return lmft_s ? lmft_s : i(OFF_LM_FEAT); }
  __device__ const double* d(int which) const { return reinterpret_cast<const double*>(base + h->off[which]); }
  __device__ const int32_t* i(int which) const { return reinterpret_cast<const int32_t*>(base + h->off[which]); }
  __device__ const uint8_t* u(int which) const { return base + h->off[which]; }
};

__device__ __forceinline__ Win decode(const SolveParams& P, int slot) {
  Win W; W.base = P.blobs + (size_t)slot * P.blob_stride; W.h = reinterpret_cast<const WinHdr*>(W.base);
  W.N = W.h->n_kf; W.M = W.h->n_feat; W.D = 15 * W.N + 7; W.Dv = 6 * W.N + 7; W.Dvp = P.sl.Dv_pad; W.nb = (W.D + TB - 1) / TB;
  return W;
}

// Header + index tables -> shared memory; W then reads them from there.  All threads; ends with a block barrier.
__device__ __forceinline__ void stage_tables(Win& W, const Smem& L, double* sm) {
  int32_t* hs = reinterpret_cast<int32_t*>(sm + L.hdr); int32_t* ps = reinterpret_cast<int32_t*>(sm + L.ptab);
  int32_t* ls = reinterpret_cast<int32_t*>(sm + L.lmst); int32_t* lf = reinterpret_cast<int32_t*>(sm + L.lmft);
  const int32_t* hg = reinterpret_cast<const int32_t*>(W.h);
  const int npair = W.h->n_pair, nlm = W.h->n_lm;
  const int32_t* pg = W.i(OFF_PAIR); const int32_t* lsg = W.i(OFF_LM_START); const int32_t* lfg = W.i(OFF_LM_FEAT);
  for (int k = threadIdx.x; k < (int)(sizeof(WinHdr) / 4); k += blockDim.x) hs[k] = hg[k];
  for (int k = threadIdx.x; k < 4 * npair; k += blockDim.x) ps[k] = pg[k];
  for (int k = threadIdx.x; k <= nlm; k += blockDim.x) ls[k] = lsg[k];
  for (int k = threadIdx.x; k < nlm; k += blockDim.x) lf[k] = lfg[k];
  __syncthreads();
  W.h = reinterpret_cast<const WinHdr*>(hs); W.pairs_s = ps; W.lmst_s = ls; W.lmft_s = lf;
}

__device__ __forceinline__ int vis2cam(int v, int N) { return v < 6 * N ? 15 * (v / 6) + v % 6 : 15 * N + (v - 6 * N); }

// Block-wide sum of one double per thread, result broadcast. red: >= SOLVE_WARPS doubles of scratch.
__device__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += red[w];
  return s;
}

// =================================================================================================================
// Projection residual only (cost evaluation): returns 1/2 rho
// =================================================================================================================
__device__ double proj_cost(const SolveParams& P, const Win& W, const double* x, int f) {
  const int np = W.h->n_proj;
  const double* c0 = W.d(OFF_PROJ); const int32_t* ix = W.i(OFF_PROJ_IDX);
  double c[14];
#pragma unroll
  for (int k = 0; k < 14; k++) c[k] = c0[(size_t)k * np + f];
  const int i = ix[f], j = ix[np + f], rank = ix[2 * np + f];
  const int feat = W.lm_feat()[rank];
  double r[2];
  vf::proj_eval(P.cfg, c, x + XP(i), x + XP(j), x + XE(W.N), x[XL(W.N) + feat], x[XT(W.N)], r, nullptr);
  double rho, w; vf::cauchy(P.cfg.cauchy_a, r[0] * r[0] + r[1] * r[1], rho, w);
  return 0.5 * rho;
}

// IMU Jacobian assembly table (vf::imu_build_table), uploaded once per device by vils_ba_create.
__device__ uint16_t g_imu_tbl[450];

// Warp-cooperative IMUFactor::Evaluate: lane 0 does the quaternion algebra (vf::imu_core) while the other lanes fetch the
// bias-Jacobian blocks of the pre-integration; all lanes then assemble the 15 x 30 Jacobian through the table and whiten it
// with sqrt_info (W, upper triangular) staged in shared memory.  The independent global loads (W, table) are issued first.
// stage (shared memory, private to the warp): J 450 | r 15 | pad | core IMU_CORE_LD | W 226.   On exit J = W J_raw, r = W r_raw.
constexpr int IMU_SLOT = 466 + vf::IMU_CORE_LD + 226;
constexpr int IMU_PROD_LD = 512;                  // per-factor products in the scratch: lower J^T J (465) | J^T r (30)
constexpr int PAIR_WARPS = (PAIR_CHUNK + 31) / 32;       // warps that evaluate projection factors in a pair-pass round
constexpr int IMU_IDLE = SOLVE_WARPS - PAIR_WARPS;   // warps free for IMU stages in every round
__device__ __forceinline__ void warp_imu_whitened(const uint16_t* tbl /* shared memory copy of g_imu_tbl */, const double* pre, const double* Wk, const double* G,
                                                  const double* pi, const double* sbi, const double* pj, const double* sbj, double* stage, bool want_J,
                                                  long long* prof = nullptr) {
  const int lane = threadIdx.x & 31;
  long long wt_ = prof ? clock64() : 0;
#define WPROF(i) do { if (prof) { const long long n_ = prof_clock(); prof[i] += n_ - wt_; wt_ = n_; } } while (0)
  double* J = stage; double* r = stage + 450; double* core = stage + 466; double* Wsm = core + vf::IMU_CORE_LD;
  // one coalesced pass brings everything the single-lane algebra reads from HBM/L2 into shared memory (the J area doubles
  // as the landing zone of the pre-integration record minus its covariance): one exposed latency instead of one per use
#pragma unroll
  for (int q = 0; q < 8; q++) { const int e = lane + 32 * q; if (e < 242) J[e] = pre[e]; }
#pragma unroll
  for (int q = 0; q < 8; q++) { const int e = lane + 32 * q; if (e < 225) Wsm[e] = Wk[e]; }
  __syncwarp();
  if (lane == 0) vf::imu_core(J, G, pi, sbi, pj, sbj, core);
  else for (int e = lane - 1; e < 36; e += 31) core[vf::IC_JAC + e] = vf::imu_core_jac(J, e);
  __syncwarp();
  WPROF(20);
  double rw = 0;
  if (lane < 15) {
#pragma unroll
    for (int m = 0; m < 15; m++) if (m >= lane) rw = fma(Wsm[lane * 15 + m], core[vf::IC_R + m], rw);
    r[lane] = rw;
  }
  if (want_J) {
#pragma unroll 1
    for (int e = lane; e < 450; e += 32) J[e] = vf::imu_tbl_value(tbl[e], core);     // rolled on purpose: a lone warp is instruction-fetch bound
    __syncwarp();
    WPROF(21);
    if (lane < 30) {                                            // J <- W J: lane owns column `lane` (no hazards)
      double col[15];
#pragma unroll
      for (int m = 0; m < 15; m++) col[m] = J[m * 30 + lane];
#pragma unroll
      for (int a = 0; a < 15; a++) { double v = 0;
#pragma unroll
        for (int m = 0; m < 15; m++) if (m >= a) v = fma(Wsm[a * 15 + m], col[m], v);
        J[a * 30 + lane] = v; }
    }
  }
  __syncwarp();
  WPROF(22);
#undef WPROF
}

// The IMU factor of the solve kernel in two warp-level stages, so that each fits inside one round of the pair pass (a lone warp next
// to FP64-busy ones is slow: the whole chain takes ~1.6 rounds).  slot (shared memory, one per FACTOR, kept between the stages):
// J 450 | r 15 | pad | work 226 (stage R: core, stage P: sqrt_info).
constexpr int IMU_SLOT2 = 466 + 226;
// Stage R: stage the pre-integration record, quaternion algebra by lane 0, assemble the raw 15 x 30 Jacobian.
__device__ void imu_stage_R(const SolveParams& P, const Win& W, const uint16_t* tbl, const double* x, int k, double* slot) {
  const int lane = threadIdx.x & 31;
  const double* pre = W.d(OFF_IMU) + (size_t)k * 467;
  double* J = slot; double* r = slot + 450; double* core = slot + 466;
  if (pre[16] > 10.0) {                                      // estimator.cpp:1182 skip if sum_dt > 10: the factor contributes nothing
    for (int e = lane; e < 466; e += 32) slot[e] = 0.0;
    __syncwarp();
    return;
  }
  const int i = W.i(OFF_IMU_KF)[k];
#pragma unroll
  for (int q = 0; q < 8; q++) { const int e = lane + 32 * q; if (e < 242) J[e] = pre[e]; }   // landing zone of the record (no covariance)
  __syncwarp();
  if (lane == 0) vf::imu_core(J, P.cfg.G, x + XP(i), x + XS(W.N, i), x + XP(i + 1), x + XS(W.N, i + 1), core);
  else for (int e = lane - 1; e < 36; e += 31) core[vf::IC_JAC + e] = vf::imu_core_jac(J, e);
  __syncwarp();
  if (lane < 15) r[lane] = core[vf::IC_R + lane];
#pragma unroll 1
  for (int e = lane; e < 450; e += 32) J[e] = vf::imu_tbl_value(tbl[e], core);     // rolled on purpose: a lone warp is instruction-fetch bound
  __syncwarp();
}
// Stage P: whiten with sqrt_info (staged over the dead core area), then the 30x30 lower product J^T J and J^T r into the per-window
// scratch (added into H later by imu_add).  Returns this lane's share of the cost.
__device__ double imu_stage_P(const SolveParams& P, const Win& W, int k, double* slot, double* scr) {
  const int lane = threadIdx.x & 31;
  double* J = slot; double* r = slot + 450; double* Wsm = slot + 466;
  const double* Wk = scr + P.sl.w_imu + (size_t)k * 225;
  double* prod = scr + P.sl.imuprod + (size_t)k * IMU_PROD_LD;
#pragma unroll
  for (int q = 0; q < 8; q++) { const int e = lane + 32 * q; if (e < 225) Wsm[e] = Wk[e]; }
  __syncwarp();
  double rw = 0;
  if (lane < 15) {
#pragma unroll
    for (int m = 0; m < 15; m++) if (m >= lane) rw = fma(Wsm[lane * 15 + m], r[m], rw);
  }
  __syncwarp();
  if (lane < 15) r[lane] = rw;
  if (lane < 30) {                                            // J <- W J: lane owns column `lane` (no hazards)
    double col[15];
#pragma unroll
    for (int m = 0; m < 15; m++) col[m] = J[m * 30 + lane];
#pragma unroll
    for (int a = 0; a < 15; a++) { double v = 0;
#pragma unroll
      for (int m = 0; m < 15; m++) if (m >= a) v = fma(Wsm[a * 15 + m], col[m], v);
      J[a * 30 + lane] = v; }
  }
  __syncwarp();
  const double cost = lane < 15 ? 0.5 * rw * rw : 0.0;
#pragma unroll 4
  for (int q = 0; q < 16; q++) {
    const int e = lane + 32 * q;
    if (e < 465) {
      int a = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
      while (a * (a + 1) / 2 > e) a--;
      while ((a + 1) * (a + 2) / 2 <= e) a++;
      const int b = e - a * (a + 1) / 2;
      double v = 0;
#pragma unroll
      for (int m = 0; m < 15; m++) v = fma(J[m * 30 + a], J[m * 30 + b], v);
      prod[e] = v;
    } else if (e < 495) {
      const int a = e - 465; double v = 0;
#pragma unroll
      for (int m = 0; m < 15; m++) v = fma(J[m * 30 + a], r[m], v);
      prod[e] = v;
    }
  }
  __syncwarp();
  return cost;
}

// H += the IMU products of imu_stage_P: one warp per factor, even keyframes first, then odd (adjacent factors share
// a 15x15 block).  Ends with a block barrier.
__device__ void imu_add(const SolveParams& P, const Win& W, double* H, double* g, double* hd, const double* scr) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nimu = W.h->n_imu;
  const int32_t* kfs = W.i(OFF_IMU_KF);
  // Warp w owns the factors 2w, 2w + 1 (+ 2 SOLVE_WARPS ...): consecutive factors have keyframes of opposite parity, so every warp has one
  // factor per parity phase (dealt k = w, w + 12, .. a warp got two factors of the SAME parity and half the warps none: the 19 factors of a
  // 20-keyframe window took two serial factors per phase).
  for (int parity = 0; parity < 2; parity++) {
#pragma unroll 1
    for (int m = 0; 2 * (warp + (m >> 1) * SOLVE_WARPS) < nimu; m++) {
      const int k = 2 * (warp + (m >> 1) * SOLVE_WARPS) + (m & 1);
      if (k >= nimu) continue;
      const int i = kfs[k];
      if ((i & 1) != parity) continue;
      const double* prod = scr + P.sl.imuprod + (size_t)k * IMU_PROD_LD;
      const int base = 15 * i;
      double pv[16];
#pragma unroll
      for (int q = 0; q < 16; q++) { const int e = lane + 32 * q; pv[q] = e < 495 ? prod[e] : 0.0; }
      if (P.h_in_smem) {
#pragma unroll
        for (int q = 0; q < 16; q++) {
          const int e = lane + 32 * q;
          if (e < 465) {
            int a = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
            while (a * (a + 1) / 2 > e) a--;
            while ((a + 1) * (a + 2) / 2 <= e) a++;
            const int b = e - a * (a + 1) / 2;
            H[tidx(base + a, base + b)] += pv[q];
            if (a == b) hd[base + a] += pv[q];
          } else if (e < 495) g[base + e - 465] += pv[q];
        }
        continue;
      }
      // H in global memory: its 15 entries per lane are read together, then written (one L2 round trip, not fifteen)
      int hi[15]; double hv[15];
#pragma unroll
      for (int q = 0; q < 15; q++) {
        const int e = lane + 32 * q;
        hi[q] = -1; hv[q] = 0.0;
        if (e < 465) {
          int a = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
          while (a * (a + 1) / 2 > e) a--;
          while ((a + 1) * (a + 2) / 2 <= e) a++;
          const int b = e - a * (a + 1) / 2;
          hi[q] = tidx(base + a, base + b); hv[q] = H[hi[q]];
          if (a == b) hd[base + a] += pv[q];
        }
      }
#pragma unroll
      for (int q = 0; q < 15; q++) if (hi[q] >= 0) H[hi[q]] = hv[q] + pv[q];
#pragma unroll
      for (int q = 14; q < 16; q++) { const int e = lane + 32 * q; if (e >= 465 && e < 495) g[base + e - 465] += pv[q]; }
    }
    __syncthreads();
  }
}

// =================================================================================================================
// V: pair pass
// =================================================================================================================
// imu_stage != nullptr: the warps that have no projection factor in a round (PAIR_WARPS..SOLVE_WARPS-1) process the IMU
// factors meanwhile, one stage (imu_stage_R / imu_stage_P) per warp per round; imu_stage holds n_imu slots of IMU_SLOT2 doubles.
__device__ double pair_pass(const SolveParams& P, const Win& W, const double* x, double* stage, double* scr, const int* pid, bool need_cost, double* imu_stage, const uint16_t* tbl, double* rot /* (N + 1) x 9 doubles of shared memory */) {
  // Rounds of PAIR_CHUNK factors in PAIR order: (1) one thread per factor evaluates ProjectionTdFactor + corrector and
  // stages the weighted rows [J(19) | r] in shared memory, writes the landmark partials / E row to the scratch;
  // (2) one warp per keyframe pair accumulates the pair-local 20x20 block A^T A over the staged rows of that pair
  // (5x3 register tile per lane).  A pair that straddles two rounds is finished by the same warp-slot with +=.
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int np = W.h->n_proj, npair = W.h->n_pair;
  const double* c0 = W.d(OFF_PROJ); const int32_t* ix = W.i(OFF_PROJ_IDX);
  const int32_t* pairs = W.pairs(); const int32_t* perm = W.i(OFF_PAIR_PERM);
  const int32_t* lm_feat = W.lm_feat();
  const uint8_t* dfix = W.u(OFF_FIXED);
  double* part = scr + P.sl.part; double* E = scr + P.sl.E; double* pairpart = scr + P.sl.pairpart;
  const int ra = 5 * (lane >> 3), cb = 3 * (lane & 7);
  double cost = 0;
  int p_lo = 0;                                   // first pair that may intersect the current round
  // per-pair context (everything that only depends on the two keyframe poses and the extrinsic), once per linearisation
  // rotation table (shared memory): R_k of every keyframe, then ric — all a projection factor needs besides the positions in x
  for (int k = threadIdx.x; k <= W.N; k += blockDim.x) {
    const vm::m3 R = vm::q2R(vm::ldq((k < W.N ? x + XP(k) : x + XE(W.N)) + 3));
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int b = 0; b < 3; b++) rot[9 * k + 3 * a + b] = R.m[a][b];
  }
  __syncthreads();
  int imu_R = 0, imu_P = 0;                      // raw stages started / product stages started (block-uniform)
  const int nimu = imu_stage ? W.h->n_imu : 0;
  long long pt_ = (P.prof && blockIdx.x == 0 && threadIdx.x == 0) ? clock64() : 0;
#define PPROF(i) do { if (P.prof && blockIdx.x == 0 && threadIdx.x == 0) { const long long n_ = prof_clock(); P.prof[i] += n_ - pt_; pt_ = n_; } } while (0)
  for (int base = 0; base < np; base += PAIR_CHUNK) {
    const int cnt_round = min(PAIR_CHUNK, np - base);
    const int t = threadIdx.x;
    if (t < cnt_round) {
      const int f = perm[base + t];
      double c[14];
#pragma unroll
      for (int k = 0; k < 14; k++) c[k] = c0[(size_t)k * np + f];
      const int kfi = ix[f], kfj = ix[np + f], rank = ix[2 * np + f];
      const int feat = lm_feat[rank];
      double* row0 = stage + (2 * t) * STAGE_LD; double* row1 = row0 + STAGE_LD;
      double r[2], J[40];
      vf::proj_eval_rows(P.cfg, c, vf::ldm(rot + 9 * kfi), vf::ldm(rot + 9 * kfj), vf::ldm(rot + 9 * W.N), vm::ld3(x + XP(kfi)), vm::ld3(x + XP(kfj)), vm::ld3(x + XE(W.N)),
                         x[XL(W.N) + feat], x[XT(W.N)], r, J);
      double rho, w; const double s2 = r[0] * r[0] + r[1] * r[1];
      if (need_cost) vf::cauchy(P.cfg.cauchy_a, s2, rho, w); else { w = vf::cauchy_w(P.cfg.cauchy_a, s2); rho = s2; }   // the log is only paid when the cost is reported
      cost += 0.5 * rho;
      r[0] *= w; r[1] *= w;
#pragma unroll
      for (int k = 0; k < 40; k++) J[k] *= w;
      const bool fixed = dfix[feat] != 0;
      const double jl0 = fixed ? 0.0 : J[18], jl1 = fixed ? 0.0 : J[38];
#pragma unroll
      for (int k = 0; k < 18; k++) { row0[k] = J[k]; row1[k] = J[20 + k]; }
      row0[18] = J[19]; row0[19] = r[0]; row1[18] = J[39]; row1[19] = r[1];
#pragma unroll
      for (int k = 20; k < STAGE_LD; k++) { row0[k] = 0; row1[k] = 0; }
      double* pt = part + (size_t)f * PART_LD;
      pt[0] = jl0 * jl0 + jl1 * jl1;
      pt[1] = jl0 * r[0] + jl1 * r[1];
#pragma unroll
      for (int k = 0; k < 6; k++) {
        pt[2 + k] = J[k] * jl0 + J[20 + k] * jl1;              // e_i
        pt[8 + k] = J[12 + k] * jl0 + J[32 + k] * jl1;         // e_ex
        E[(size_t)rank * W.Dvp + 6 * kfj + k] = J[6 + k] * jl0 + J[26 + k] * jl1;   // e_j: this factor only
      }
      pt[14] = J[19] * jl0 + J[39] * jl1;                      // e_td
    } else if (warp >= PAIR_WARPS) {
      // idle warps: new raw stages first (so that no stage is left for after the last round), products of completed raws next
      const int w = warp - PAIR_WARPS, nR = min(nimu - imu_R, IMU_IDLE), nP = min(IMU_IDLE - nR, imu_R - imu_P);
      if (w < nR) imu_stage_R(P, W, tbl, x, imu_R + w, imu_stage + (imu_R + w) * IMU_SLOT2);
      else if (w - nR < nP) cost += imu_stage_P(P, W, imu_P + w - nR, imu_stage + (imu_P + w - nR) * IMU_SLOT2, scr);
    }
    { const int nR = min(nimu - imu_R, IMU_IDLE), nP = min(IMU_IDLE - nR, imu_R - imu_P); imu_P += nP; imu_R += nR; }
    PPROF(16);
    __syncthreads();
    PPROF(17);
    // pairs intersecting [base, base + cnt_round)
    while (p_lo < npair && pairs[4 * p_lo] + pairs[4 * p_lo + 1] <= base) p_lo++;
    for (int p = p_lo + warp; p < npair; p += SOLVE_WARPS) {
      const int start = pairs[4 * p], cnt = pairs[4 * p + 1];
      if (start >= base + cnt_round) break;
      const int s0 = max(start, base), s1 = min(start + cnt, base + cnt_round);
      double acc[5][3];
#pragma unroll
      for (int a = 0; a < 5; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) acc[a][b] = 0;
      const double* sp = stage + (size_t)(2 * (s0 - base)) * STAGE_LD;
#pragma unroll 4
      for (int row = 0; row < 2 * (s1 - s0); row++, sp += STAGE_LD) {
        double av[5], bv[3];
#pragma unroll
        for (int a = 0; a < 5; a++) av[a] = sp[ra + a];
#pragma unroll
        for (int b = 0; b < 3; b++) bv[b] = sp[cb + b];
#pragma unroll
        for (int a = 0; a < 5; a++)
#pragma unroll
          for (int b = 0; b < 3; b++) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
      }
      double* out = pairpart + (size_t)p * PAIR_LD * PAIR_LD;
      const bool first = s0 == start;
#pragma unroll
      for (int a = 0; a < 5; a++)
#pragma unroll
        for (int b = 0; b < 3; b++)
          if (cb + b < PAIR_LD) { double* o = out + (ra + a) * PAIR_LD + cb + b; *o = first ? acc[a][b] : *o + acc[a][b]; }
    }
    PPROF(18);
    __syncthreads();
    PPROF(19);
  }
#undef PPROF
  // stages that did not fit into the projection rounds (few projection factors, or more than two rounds' worth of IMU factors)
  while (imu_P < nimu) {
    const int nR = min(nimu - imu_R, IMU_IDLE), nP = min(IMU_IDLE - nR, imu_R - imu_P);
    if (warp >= PAIR_WARPS) {
      const int w = warp - PAIR_WARPS;
      if (w < nR) imu_stage_R(P, W, tbl, x, imu_R + w, imu_stage + (imu_R + w) * IMU_SLOT2);
      else if (w - nR < nP) cost += imu_stage_P(P, W, imu_P + w - nR, imu_stage + (imu_P + w - nR) * IMU_SLOT2, scr);
    }
    imu_P += nP; imu_R += nR;
    __syncthreads();
  }
  return cost;
}

// L: per-landmark reduction of the factor partials (fixed order) -> cinv, glam, E rows (anchor / ex / td parts)
// jac_mode (dogleg): 0 = plain clamp(C) damping; 1 = first linearisation: also fix the landmark Jacobi scale s = 1 / (1 + sqrt(C)); 2 = use the
// stored scale.  With a scale the damping diagonal is clamp(C s^2, 1e-6, 1e32) / s^2 and the raw C is kept in craw.
__device__ void landmark_reduce(const SolveParams& P, const Win& W, double* cinv, double* glam, double* scr, double mu, int jac_mode = 0, double* craw = nullptr, double* scl = nullptr) {
  const int nlm = W.h->n_lm, np = W.h->n_proj;
  const int32_t* lm_start = W.lm_start(); const int32_t* ix = W.i(OFF_PROJ_IDX);
  const double* part = scr + P.sl.part; double* E = scr + P.sl.E;
  for (int rnk = threadIdx.x; rnk < nlm; rnk += blockDim.x) {
    double s[15];
#pragma unroll
    for (int k = 0; k < 15; k++) s[k] = 0;
    const int f0 = lm_start[rnk], f1 = lm_start[rnk + 1];
    for (int f = f0; f < f1; f++) {
      const double* pt = part + (size_t)f * PART_LD;
#pragma unroll
      for (int k = 0; k < 15; k++) s[k] += pt[k];
    }
    const int kfi = ix[f0];
    double* e = E + (size_t)rnk * W.Dvp;
#pragma unroll
    for (int k = 0; k < 6; k++) { e[6 * kfi + k] = s[2 + k]; e[6 * W.N + k] = s[8 + k]; }
    e[6 * W.N + 6] = s[14];
    const double C = s[0];
    if (C > 0.0) {   // free landmark: (C + mu clamp(C))^-1   [ceres min/max_lm_diagonal 1e-6 / 1e32 on the squared column norm]
      double dd = fmin(fmax(C, 1e-6), 1e32);
      if (jac_mode) {
        const double sj = jac_mode == 1 ? 1.0 / (1.0 + sqrt(C)) : scl[rnk];
        if (jac_mode == 1) scl[rnk] = sj;
        dd = fmin(fmax(C * sj * sj, 1e-6), 1e32) / (sj * sj);
        craw[rnk] = C;
      }
      cinv[rnk] = 1.0 / (C + mu * dd);
      glam[rnk] = s[1];
    } else { cinv[rnk] = 0.0; glam[rnk] = 0.0; if (jac_mode) craw[rnk] = 0.0; }
  }
}

// S: Hv(lower) = -sum_f cinv_f e_f e_f^T ; gv = -sum_f cinv_f glam_f e_f.  3x3 register tiles over the lower triangle.
__device__ void schur_syrk(const SolveParams& P, const Win& W, const double* cinv, const double* glam, double* Hv, double* gv,
                           double* chunk, const double* scr, int chunk_cap /* doubles available at `chunk` */) {
  const int ECH = max(ECHUNK, chunk_cap / W.Dvp);     // as many landmarks per staging pass as fit (all of them for a 10-keyframe window)
  const int Dv = W.Dv, Dvp = W.Dvp, nlm = W.h->n_lm;
  const int nt = (Dv + 2) / 3;                 // tiles per side
  const int ntiles = nt * (nt + 1) / 2;
  const double* E = scr + P.sl.E;
  for (int tbase = 0; tbase < ntiles; tbase += blockDim.x) {
  // tile id -> (ti >= tj)
  int ti = -1, tj = 0;
  double acc[3][3], gacc[3];
  const int t = tbase + threadIdx.x;
  if (t < ntiles) {
    ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while (ti * (ti + 1) / 2 > t) ti--;
    while ((ti + 1) * (ti + 2) / 2 <= t) ti++;
    tj = t - ti * (ti + 1) / 2;
  }
#pragma unroll
  for (int a = 0; a < 3; a++) { gacc[a] = 0;
#pragma unroll
    for (int b = 0; b < 3; b++) acc[a][b] = 0; }
  for (int f0 = 0; f0 < nlm; f0 += ECH) {
    const int n = min(ECH, nlm - f0);
    __syncthreads();
    for (int k = threadIdx.x; k < n * Dvp; k += blockDim.x) cp_async_f64(chunk + k, E + (size_t)f0 * Dvp + k);
    cp_async_wait_all();
    __syncthreads();
    if (ti >= 0) {
      for (int f = 0; f < n; f++) {
        const double ci = cinv[f0 + f];
        const double* e = chunk + f * Dvp;
        double av[3], bv[3];
#pragma unroll
        for (int a = 0; a < 3; a++) { const int ia = 3 * ti + a; av[a] = ia < Dv ? e[ia] : 0.0; }
#pragma unroll
        for (int b = 0; b < 3; b++) { const int ib = 3 * tj + b; bv[b] = ib < Dv ? e[ib] * ci : 0.0; }
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
          for (int b = 0; b < 3; b++) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
        if (tj == 0) {
          const double gl = glam[f0 + f] * ci;
#pragma unroll
          for (int a = 0; a < 3; a++) gacc[a] = fma(av[a], gl, gacc[a]);
        }
      }
    }
  }
  if (ti >= 0) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const int ia = 3 * ti + a;
      if (ia >= Dv) continue;
#pragma unroll
      for (int b = 0; b < 3; b++) { const int ib = 3 * tj + b; if (ib < Dv && ib <= ia) Hv[ia * Dvp + ib] = -acc[a][b]; }
      if (tj == 0) gv[ia] = -gacc[a];
    }
  }
  }
}

// G: Hv (Schur part) + pair blocks (direct part) -> H tiles (lower), g, hd.  One owner per entry.
// pid: (i,j) -> pair index table in shared memory, -1 entries redirected to an all-zero block (index zblk) so the accumulation loops are
// branch-free.  Every load of a pair block is an L2 round trip, so each family of entries is written to keep many of them in flight:
//   (1) pose block q (its own lower triangle) and the extrinsic / td rows against pose block q: one term per OTHER keyframe j (the pair (q, j)
//       with q as the anchor if j > q, as the observer if j < q), loaded ten at a time with no branch in between;
//   (2) the gradient / diagonal of the pose rows: the same walk, two loads per keyframe;
//   (3) pose block p against pose block q < p: exactly one pair block, a block-wise copy;
//   (4) extrinsic / td rows against themselves: every pair contributes, one warp per entry with the lanes over the pairs.
// (As one task loop with a serial sum per entry the gather took 45 k cycles per linearisation of a config-2 window, the L2 latency of
// family (1) and (4) paid term by term.)
struct GatherTask { int a, b, q, off_anchor, off_obs; };

// entry e of family (1): N x 21 in-block lower entries, then nsh x 6N entries of the shared rows
__device__ __forceinline__ GatherTask gather_task1(int e, int N) {
  GatherTask T;
  if (e < 21 * N) {
    const int blk = e / 21, w = e - blk * 21;
    int ao = 0; while ((ao + 1) * (ao + 2) / 2 <= w) ao++;
    const int bo = w - ao * (ao + 1) / 2;
    T.a = 6 * blk + ao; T.b = 6 * blk + bo; T.q = blk;
    T.off_anchor = ao * PAIR_LD + bo; T.off_obs = (6 + ao) * PAIR_LD + 6 + bo;
  } else {
    const int e2 = e - 21 * N, ia = e2 / (6 * N), b = e2 - ia * 6 * N, bo = b % 6, la = ia < 6 ? 12 + ia : 18;
    T.a = 6 * N + ia; T.b = b; T.q = b / 6;
    T.off_anchor = la * PAIR_LD + bo; T.off_obs = la * PAIR_LD + 6 + bo;
  }
  return T;
}

// sum over the keyframes j != q of blk(q, j)[j > q ? off_anchor : off_obs]
__device__ __forceinline__ double gather_walk(const double* __restrict__ pp, const int* __restrict__ pid, int N, int q, int off_anchor, int off_obs, int zblk) {
  double s = 0;
  for (int j0 = 0; j0 < N; j0 += 10) {
    double v[10];
#pragma unroll
    for (int jj = 0; jj < 10; jj++) {
      const int j = j0 + jj;
      const int pr = (j < N && j != q) ? (j > q ? pid[q * N + j] : pid[j * N + q]) : -1;
      v[jj] = pp[(size_t)(pr < 0 ? zblk : pr) * (PAIR_LD * PAIR_LD) + (j > q ? off_anchor : off_obs)];
    }
#pragma unroll
    for (int jj = 0; jj < 10; jj++) s += v[jj];
  }
  return s;
}

__device__ __forceinline__ void gather_pair_of(int pi, int& pb, int& qb) {   // pi-th pose pair (pb > qb) in row-major lower-triangular order
  pb = (int)((1.0f + sqrtf(8.0f * pi + 1.0f)) * 0.5f);
  while (pb * (pb - 1) / 2 > pi) pb--;
  while ((pb + 1) * pb / 2 <= pi) pb++;
  qb = pi - pb * (pb - 1) / 2;
}

__device__ void gather_visual(const SolveParams& P, const Win& W, const double* Hv, const double* gv, double* H, double* g, double* hd,
                              const double* scr, const int* pid, int zblk) {
  const int N = W.N, Dv = W.Dv, Dvp = W.Dvp, nsh = Dv - 6 * N, npair = W.h->n_pair, lane = threadIdx.x & 31;
  const double* __restrict__ pp = scr + P.sl.pairpart;
  // (1)
  for (int e = threadIdx.x; e < 21 * N + nsh * 6 * N; e += blockDim.x) {
    const GatherTask T = gather_task1(e, N);
    const double s = gather_walk(pp, pid, N, T.q, T.off_anchor, T.off_obs, zblk);
    H[tidx(vis2cam(T.a, N), vis2cam(T.b, N))] = Hv[T.a * Dvp + T.b] + s;
  }
  // (2)
  for (int e = threadIdx.x; e < 12 * N; e += blockDim.x) {
    const int a = e >> 1, q = a / 6, ao = a - 6 * q, which = e & 1;      // which: 0 gradient (column 19), 1 diagonal
    const double s = gather_walk(pp, pid, N, q, ao * PAIR_LD + (which ? ao : 19), (6 + ao) * PAIR_LD + (which ? 6 + ao : 19), zblk);
    const int ca = vis2cam(a, N);
    if (which) hd[ca] = s; else g[ca] = gv[a] + s;
  }
  // (3)
  {
    const double* __restrict__ Hvr = Hv; double* __restrict__ Hr = H;
    const int nent = N * (N - 1) / 2 * 36;
#pragma unroll 4
    for (int e = threadIdx.x; e < nent; e += blockDim.x) {
      const int pi = e / 36, o = e - pi * 36, ao = o / 6, bo = o - ao * 6;
      int pb, qb; gather_pair_of(pi, pb, qb);
      const int pr = pid[qb * N + pb];
      const double v = pp[(size_t)(pr < 0 ? zblk : pr) * (PAIR_LD * PAIR_LD) + (6 + ao) * PAIR_LD + bo];
      const int a = 6 * pb + ao, b = 6 * qb + bo;
      Hr[tidx(vis2cam(a, N), vis2cam(b, N))] = Hvr[a * Dvp + b] + v;
    }
  }
  // (4)
  for (int wt = threadIdx.x >> 5; wt < nsh * (nsh + 1) / 2 + nsh; wt += SOLVE_WARPS) {
    int ia, ib; const bool grad = wt < nsh;
    if (grad) { ia = wt; ib = wt; }
    else { const int u = wt - nsh; ia = 0; while ((ia + 1) * (ia + 2) / 2 <= u) ia++; ib = u - ia * (ia + 1) / 2; }
    const int la = ia < 6 ? 12 + ia : 18, lb = ib < 6 ? 12 + ib : 18;
    double s0 = 0, s1 = 0;
    for (int pr = lane; pr < npair; pr += 32) {
      const double* blk = pp + (size_t)pr * (PAIR_LD * PAIR_LD);
      if (grad) { s0 += blk[la * PAIR_LD + 19]; s1 += blk[la * PAIR_LD + la]; }
      else s0 += blk[la * PAIR_LD + lb];
    }
    s0 = warp_sum(s0); if (grad) s1 = warp_sum(s1);
    if (lane == 0) {
      const int a = 6 * N + ia, b = 6 * N + ib, ca = vis2cam(a, N), cbm = vis2cam(b, N);
      if (grad) { g[ca] = gv[a] + s0; hd[ca] = s1; }
      else H[tidx(ca, cbm)] = Hv[a * Dvp + b] + s0;
    }
  }
}

// =================================================================================================================
// I: IMU factors. One warp per factor; even keyframe index first, then odd (adjacent factors share a 15x15 block).
// stage: 466 doubles per warp slot (J 15x30 | r 15).
// =================================================================================================================
__device__ double imu_pass(const SolveParams& P, const Win& W, const double* x, double* H, double* g, double* hd, double* imu_stage,
                           const double* scr, bool want_J, bool all_slots) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nimu = W.h->n_imu;
  const double* pre_all = W.d(OFF_IMU); const int32_t* kfs = W.i(OFF_IMU_KF);
  const double* Wall = scr + P.sl.w_imu;
  double cost = 0;
  if (all_slots && nimu <= SOLVE_WARPS) {
    // Fast path: every factor has its own warp and staging slot; the 30x30 products are formed in registers by all
    // factors concurrently and added to H in two barrier-separated phases (adjacent factors share a 15x15 block).
    const int k = warp;
    const bool mine = k < nimu && pre_all[(size_t)(k < nimu ? k : 0) * 467 + 16] <= 10.0;   // estimator.cpp:1182 skip if sum_dt > 10
    double hv[16]; int i = 0;
    long long it_ = (P.prof && blockIdx.x == 0 && threadIdx.x == 0) ? clock64() : 0;
#define IPROF(i) do { if (P.prof && blockIdx.x == 0 && threadIdx.x == 0) { const long long n_ = prof_clock(); P.prof[i] += n_ - it_; it_ = n_; } } while (0)
#pragma unroll
    for (int e = 0; e < 16; e++) hv[e] = 0;
    if (mine) {
      i = kfs[k];
      double* J = imu_stage + k * 466; double* r = J + 450;
      const double* pre = pre_all + (size_t)k * 467; const double* Wk = Wall + (size_t)k * 225;
      for (int e = lane; e < 450; e += 32) J[e] = 0;
      __syncwarp();
      // one lane, straight-line: measured 4-5 k cycles; splitting the blocks over lanes with a switch DIVERGES and costs 12 k
      if (lane == 0) vf::imu_eval_raw(pre, P.cfg.G, x + XP(i), x + XS(W.N, i), x + XP(i + 1), x + XS(W.N, i + 1), r, want_J ? J : nullptr);
      __syncwarp();
      IPROF(20);
      double rw = 0;
      if (lane < 15) { for (int m = lane; m < 15; m++) rw = fma(Wk[lane * 15 + m], r[m], rw); }
      __syncwarp();
      if (lane < 15) { r[lane] = rw; cost += 0.5 * rw * rw; }
      if (want_J) {
        if (lane < 30) {                                    // J <- W J: lane owns column `lane` (no hazards)
          double col[15];
#pragma unroll
          for (int m = 0; m < 15; m++) col[m] = J[m * 30 + lane];
#pragma unroll
          for (int a = 0; a < 15; a++) { double v = 0;
#pragma unroll
            for (int m = 0; m < 15; m++) if (m >= a) v = fma(Wk[a * 15 + m], col[m], v);
            J[a * 30 + lane] = v; }
        }
        __syncwarp();
        IPROF(21);
#pragma unroll
        for (int q = 0; q < 16; q++) {
          const int e = lane + 32 * q;
          if (e < 465) {
            int a = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
            while (a * (a + 1) / 2 > e) a--;
            while ((a + 1) * (a + 2) / 2 <= e) a++;
            const int b = e - a * (a + 1) / 2;
            double v = 0;
#pragma unroll
            for (int m = 0; m < 15; m++) v = fma(J[m * 30 + a], J[m * 30 + b], v);
            hv[q] = v;
          } else if (e < 495) {
            const int a = e - 465; double v = 0;
#pragma unroll
            for (int m = 0; m < 15; m++) v = fma(J[m * 30 + a], r[m], v);
            hv[q] = v;
          }
        }
      }
    }
    IPROF(22);
    if (want_J) {
      for (int parity = 0; parity < 2; parity++) {
        if (mine && (i & 1) == parity) {
          const int base = 15 * i;
#pragma unroll
          for (int q = 0; q < 16; q++) {
            const int e = lane + 32 * q;
            if (e < 465) {
              int a = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
              while (a * (a + 1) / 2 > e) a--;
              while ((a + 1) * (a + 2) / 2 <= e) a++;
              const int b = e - a * (a + 1) / 2;
              H[tidx(base + a, base + b)] += hv[q];
              if (a == b) hd[base + a] += hv[q];
            } else if (e < 495) g[base + e - 465] += hv[q];
          }
        }
        __syncthreads();
      }
    }
    IPROF(23);
#undef IPROF
    return cost;
  }
  const int nslot = (P.Ncap + 1) / 2;
  for (int parity = 0; parity < 2; parity++) {
    // factors of this parity, in index order: slot s handles the s-th one (s strided by available warps)
    int seen = 0;
    for (int k = 0; k < nimu; k++) {
      const int i = kfs[k];
      if ((i & 1) != parity) continue;
      const int s = seen++;
      if ((s % SOLVE_WARPS) != warp) continue;     // same-parity factors <= nslot, so stage slot s is private
      double* J = imu_stage + (s % nslot) * 466; double* r = J + 450;
      const double* pre = pre_all + (size_t)k * 467; const double* Wk = Wall + (size_t)k * 225;
      if (pre[16] > 10.0) continue;                                  // estimator.cpp:1182 skip if sum_dt > 10
      for (int e = lane; e < 450; e += 32) J[e] = 0;
      __syncwarp();
      // one lane, straight-line: measured 4-5 k cycles; splitting the blocks over lanes with a switch DIVERGES and costs 12 k
      if (lane == 0) vf::imu_eval_raw(pre, P.cfg.G, x + XP(i), x + XS(W.N, i), x + XP(i + 1), x + XS(W.N, i + 1), r, want_J ? J : nullptr);
      __syncwarp();
      // r <- W r (W upper triangular), in place top-down
      double rw = 0;
      if (lane < 15) { for (int m = lane; m < 15; m++) rw = fma(Wk[lane * 15 + m], r[m], rw); }
      __syncwarp();
      if (lane < 15) { r[lane] = rw; cost += 0.5 * rw * rw; }
      if (want_J) {
        for (int row = 0; row < 15; row++) {                          // J <- W J in place: row `row` needs rows >= row only
          double v = 0;
          if (lane < 30) for (int m = row; m < 15; m++) v = fma(Wk[row * 15 + m], J[m * 30 + lane], v);
          __syncwarp();
          if (lane < 30) J[row * 30 + lane] = v;
        }
        __syncwarp();
        const int base = 15 * i;
        for (int e = lane; e < 465 + 30; e += 32) {
          if (e < 465) {
            int a = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
            while (a * (a + 1) / 2 > e) a--;
            while ((a + 1) * (a + 2) / 2 <= e) a++;
            const int b = e - a * (a + 1) / 2;
            double v = 0;
#pragma unroll
            for (int m = 0; m < 15; m++) v = fma(J[m * 30 + a], J[m * 30 + b], v);
            H[tidx(base + a, base + b)] += v;
            if (a == b) hd[base + a] += v;
          } else {
            const int a = e - 465;
            double v = 0;
#pragma unroll
            for (int m = 0; m < 15; m++) v = fma(J[m * 30 + a], r[m], v);
            g[base + a] += v;
          }
        }
      }
      __syncwarp();
    }
    if (want_J) __syncthreads();
  }
  return cost;
}

// LiDAR plane + edge factors of one keyframe reduced by one warp (HuberLoss corrector), into the pose diagonal block.
__device__ double lidar_pass(const SolveParams& P, const Win& W, const double* x, double* H, double* g, double* hd, bool want_J) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int npl = W.h->n_plane, ned = W.h->n_edge;
  const double* pl = W.d(OFF_PLANE); const double* ed = W.d(OFF_EDGE);
  const int32_t* pls = W.i(OFF_PLANE_START); const int32_t* eds = W.i(OFF_EDGE_START);
  double cost = 0;
  for (int k = warp; k < W.N; k += SOLVE_WARPS) {
    const double* pose = x + XP(k);
    double A[21], gg[6];
#pragma unroll
    for (int e = 0; e < 21; e++) A[e] = 0;
#pragma unroll
    for (int e = 0; e < 6; e++) gg[e] = 0;
    if (npl) for (int f = pls[k] + lane; f < pls[k + 1]; f += 32) {
      const vm::v3 pb = vm::mk(pl[f], pl[(size_t)npl + f], pl[(size_t)2 * npl + f]);
      const vm::v3 n = vm::mk(pl[(size_t)3 * npl + f], pl[(size_t)4 * npl + f], pl[(size_t)5 * npl + f]);
      double J[6];
      double r = vf::plane_eval(pose, pb, n, pl[(size_t)6 * npl + f], want_J ? J : nullptr);
      double rho, w; vf::huber(P.cfg.huber_a, r * r, rho, w);
      cost += 0.5 * rho;
      if (want_J) {
        r *= w;
        int e = 0;
#pragma unroll
        for (int a = 0; a < 6; a++) { J[a] *= w; }
#pragma unroll
        for (int a = 0; a < 6; a++) { gg[a] = fma(J[a], r, gg[a]);
#pragma unroll
          for (int b = 0; b <= a; b++) { A[e] = fma(J[a], J[b], A[e]); e++; } }
      }
    }
    if (ned) for (int f = eds[k] + lane; f < eds[k + 1]; f += 32) {
      const vm::v3 pb = vm::mk(ed[f], ed[(size_t)ned + f], ed[(size_t)2 * ned + f]);
      const vm::v3 a_ = vm::mk(ed[(size_t)3 * ned + f], ed[(size_t)4 * ned + f], ed[(size_t)5 * ned + f]);
      const vm::v3 b_ = vm::mk(ed[(size_t)6 * ned + f], ed[(size_t)7 * ned + f], ed[(size_t)8 * ned + f]);
      double r[3], J[18];
      vf::edge_eval(pose, pb, a_, b_, r, want_J ? J : nullptr);
      double rho, w; vf::huber(P.cfg.huber_a, r[0] * r[0] + r[1] * r[1] + r[2] * r[2], rho, w);
      cost += 0.5 * rho;
      if (want_J) {
#pragma unroll
        for (int m = 0; m < 3; m++) {
          const double rm = r[m] * w;
          int e = 0;
#pragma unroll
          for (int a = 0; a < 6; a++) { const double ja = J[m * 6 + a] * w; gg[a] = fma(ja, rm, gg[a]);
#pragma unroll
            for (int b = 0; b <= a; b++) { A[e] = fma(ja, J[m * 6 + b] * w, A[e]); e++; } }
        }
      }
    }
    if (want_J) {
#pragma unroll
      for (int e = 0; e < 21; e++) A[e] = warp_sum(A[e]);
#pragma unroll
      for (int e = 0; e < 6; e++) gg[e] = warp_sum(gg[e]);
      if (lane == 0) {
        int e = 0;
        for (int a = 0; a < 6; a++) { g[15 * k + a] += gg[a];
          for (int b = 0; b <= a; b++) { H[tidx(15 * k + a, 15 * k + b)] += A[e]; if (a == b) hd[15 * k + a] += A[e]; e++; } }
      }
    }
  }
  return cost;
}

// ICP (<=5) and LPS (<=7) autodiff constraints: thread 0..n evaluates into smem, then the whole block adds one factor at a time.
__device__ double icp_lps_pass(const SolveParams& P, const Win& W, const double* x, double* H, double* g, double* hd, double* stage, bool want_J) {
  const int nicp = W.h->n_icp, nlps = W.h->n_lps;
  if (nicp + nlps == 0) return 0.0;
  const double* icp = W.d(OFF_ICP); const double* lps = W.d(OFF_LPS);
  double cost = 0;
  // stage per factor: 76 doubles (r 3 | J 3x24 | pad)
  const int t = threadIdx.x;
  if (t < nicp) {
    const double* c = icp + 14 * t; double* s = stage + 76 * t;
    vf::icp_eval(c, x + XP((int)c[10]), x + XP((int)c[11]), x + XP((int)c[12]), x + XP((int)c[13]), s, s + 3);
    double rho, w; vf::cauchy(P.cfg.cauchy_a, s[0] * s[0] + s[1] * s[1] + s[2] * s[2], rho, w);
    cost += 0.5 * rho;
    for (int k = 0; k < 75; k++) s[k] *= w;
  } else if (t >= 32 && t - 32 < nlps) {            // the LPS constraints on the second warp: next to the ICP ones, not after them
    const int f = nicp + t - 32;
    const double* c = lps + 9 * (f - nicp); double* s = stage + 76 * f;
    vf::lps_eval(c, x + XP((int)c[7]), x + XP((int)c[8]), s, s + 3);
    double rho, w; vf::cauchy(P.cfg.cauchy_a, s[0] * s[0] + s[1] * s[1] + s[2] * s[2], rho, w);
    cost += 0.5 * rho;
    for (int k = 0; k < 39; k++) s[k] *= w;
  }
  if (!want_J) return cost;
  __syncthreads();
  for (int f = 0; f < nicp + nlps; f++) {
    const bool is_icp = f < nicp;
    const int nbk = is_icp ? 4 : 2, wdt = 6 * nbk;
    const double* c = is_icp ? icp + 14 * f : lps + 9 * (f - nicp);
    const double* s = stage + 76 * f;
    int kf[4];
    for (int b = 0; b < nbk; b++) kf[b] = (int)(is_icp ? c[10 + b] : c[7 + b]);
    for (int e = t; e < wdt * wdt + wdt; e += blockDim.x) {
      if (e < wdt * wdt) {
        const int a = e / wdt, b = e % wdt;
        const int ca = 15 * kf[a / 6] + a % 6, cb2 = 15 * kf[b / 6] + b % 6;
        // the pose blocks of one constraint are distinct (ceres rejects duplicates) but not ordered: keep ca >= cb
        if (ca < cb2) continue;
        const double v = s[3 + a] * s[3 + b] + s[3 + wdt + a] * s[3 + wdt + b] + s[3 + 2 * wdt + a] * s[3 + 2 * wdt + b];
        H[tidx(ca, cb2)] += v; if (ca == cb2) hd[ca] += v;
      } else {
        const int a = e - wdt * wdt;
        g[15 * kf[a / 6] + a % 6] += s[3 + a] * s[0] + s[3 + wdt + a] * s[1] + s[3 + 2 * wdt + a] * s[2];
      }
    }
    __syncthreads();
  }
  return cost;
}

// MarginalizationFactor (factor/marginalization_factor.cpp:352-400): r = r_lin + J_lin dx. A = J^T J and b0 = J^T r_lin are
// constant across iterations and precomputed by prep_kernel, so g += b0 + A dx, H += A, cost = 1/2 |r|^2.
__device__ double prior_pass(const SolveParams& P, const Win& W, const double* x, double* H, double* g, double* hd, double* dxp /* n doubles smem */,
                             const double* scr, bool want_J, bool add_H = true /* false: H += A was done by prior_add_H_cluster */) {
  const int n = W.h->prior_n;
  if (n == 0) return 0.0;
  const int nblk = W.h->prior_nblk;
  const int32_t* blk = W.i(OFF_PRIOR_BLK); const int32_t* col = W.i(OFF_PRIOR_COL);
  const double* x0 = W.d(OFF_PRIOR_X0); const double* Jl = W.d(OFF_PRIOR_J); const double* rl = W.d(OFF_PRIOR_R);
  const double* A = scr + P.sl.priorA; const double* b0 = scr + P.sl.priorb0;
  __syncthreads();
  for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
    const int type = blk[4 * b], idx = blk[4 * b + 1], xo = blk[4 * b + 2], c0 = blk[4 * b + 3];
    const double* xb = type == VILS_BLK_POSE ? x + XP(idx) : type == VILS_BLK_SPEEDBIAS ? x + XS(W.N, idx) : type == VILS_BLK_EXPOSE ? x + XE(W.N) : x + XT(W.N);
    if (type == VILS_BLK_POSE || type == VILS_BLK_EXPOSE) vf::prior_dx_pose(xb, x0 + xo, dxp + c0);
    else { const int sz = type == VILS_BLK_SPEEDBIAS ? 9 : 1; for (int k = 0; k < sz; k++) dxp[c0 + k] = xb[k] - x0[xo + k]; }
  }
  __syncthreads();
  double cost = 0;
  // (operands from L2 / HBM: eight loads in flight per thread, the sums keep their order)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double r = rl[i];
#pragma unroll 8
    for (int j = 0; j < n; j++) r = fma(Jl[(size_t)j * n + i], dxp[j], r);
    cost += 0.5 * r * r;
  }
  if (want_J) {
    for (int e = threadIdx.x; e < n * n + n; e += blockDim.x) {
      if (e < n * n) {
        if (!add_H) continue;
        const int a = e / n, b = e % n;
        const int ca = col[a], cb2 = col[b];
        if (ca < cb2) continue;
        H[tidx(ca, cb2)] += A[(size_t)a * n + b];
        if (ca == cb2) hd[ca] += A[(size_t)a * n + b];
      } else {
        const int a = e - n * n;
        double v = b0[a];
#pragma unroll 8
        for (int j = 0; j < n; j++) v = fma(A[(size_t)j * n + a], dxp[j], v);   // A is symmetric bit for bit (prep_window): column a read as row a, coalesced
        g[col[a]] += v;
      }
    }
  }
  return cost;
}

// 16x16 diagonal tile factorisation by ONE warp.  Lane (r = lane & 15) owns row r in a 16-register window that SLIDES:
// a[k] always holds A[r][j + k], so the column loop stays ROLLED with static register indices (the fully unrolled form
// was 2000 instructions = 32 KB and instruction-fetch bound for a single warp).  1/sqrt by rsqrt, no FP64 divide.
// The pivot chain never touches shared memory: column j + 1 (the next pivot column) receives the update of column j through ONE
// register broadcast (L[j+1][j] from lane j + 1), the other columns receive it one iteration later from the copy of L[:, j] that
// went to the tile in shared memory meanwhile.  Chain per pivot: SHFL + rsqrt + DMUL + SHFL + DFMA ~ 90 cycles; with the
// STS -> LDS round trip in it the chain measured 430 cycles per pivot inside the solve kernel, because the trailing-update warps
// of the other sub-partitions keep the shared-memory pipe busy.
// Entries above the diagonal are don't-care.  dinv[16] receives 1/L_jj.  Lanes 16..31 mirror 0..15.
#ifndef VILS_DIAG_PIPELINED
#define VILS_DIAG_PIPELINED 1
#endif
// The tile is addressed through TileMem: explicit ld/st.shared when the reduced system lives in shared memory (a generic access costs two
// R2UR descriptor moves per load, and the lone warp on this chain is issue-bound: 5.8 k -> see tools/diag_micro.cu), plain pointers otherwise.
template <bool SH> struct TileMem;
template <> struct TileMem<false> {
  double* p;
  __device__ __forceinline__ explicit TileMem(double* q) : p(q) {}
  __device__ __forceinline__ double ld(int i) const { return p[i]; }
  __device__ __forceinline__ void st(int i, double v) const { p[i] = v; }
  __device__ __forceinline__ void advance(int i) { p += i; }
};
template <> struct TileMem<true> {
  uint32_t a;
  __device__ __forceinline__ explicit TileMem(double* q) : a((uint32_t)__cvta_generic_to_shared(q)) {}
  __device__ __forceinline__ double ld(int i) const { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a + 8u * (uint32_t)i) : "memory"); return v; }
  __device__ __forceinline__ void st(int i, double v) const { asm volatile("st.shared.f64 [%0], %1;" :: "r"(a + 8u * (uint32_t)i), "d"(v) : "memory"); }
  __device__ __forceinline__ void advance(int i) { a += 8u * (uint32_t)i; }
};

template <bool SH>
__device__ __noinline__ void chol_diag_factor(double* Akk, double* dinv_, int* flag, const int lane) {
  // (lane comes in as a register and 1/L_jj leaves after the loop: a re-read of %tid and the divergent one-lane store of the rolled loop,
  // with its shared-window S2UR, sat on the pivot chain)
  const int r = lane & 15;
  const bool lo = lane < 16;
  double a[16], mydi = 0.0;
  TileMem<SH> dst(Akk + r * TLD), col(Akk), dinv(dinv_);
#pragma unroll
  for (int c = 0; c < 16; c++) a[c] = dst.ld(c);
  __syncwarp();                                      // every lane has its row before the first column is written back
  bool bad = false;
#if VILS_DIAG_PIPELINED
  // A warp issues in order, so the statement order below is the schedule: the loads and FMAs of the pending column sit in the shadow of
  // the rsqrt chain; the chain itself is SHFL -> rsqrt -> DMUL -> DFMA (lane j + 1 forms its own next pivot A[j+1][j+1] - L[j+1][j]^2
  // locally) -> SHFL.
  auto pivot = [&](const double c0, const double ajj, const int j, double& a0, double& nxt_ajj) {
    bad |= !(ajj > 0.0);                             // also catches NaN
    const double di = rsqrt(ajj);
    a0 = c0 * di;                                    // L[r][j] (rows r < j: don't-care)
    const double dn = fma(-a0, a0, a[1]);            // lane j + 1: its next pivot
    nxt_ajj = __shfl_sync(0xffffffffu, dn, (j + 1) & 15);
    mydi = (lane == j) ? di : mydi;
  };
  double c0 = a[0], a0, ajj = __shfl_sync(0xffffffffu, a[0], 0), nxt;
  pivot(c0, ajj, 0, a0, nxt);
  {
    const double l1 = __shfl_sync(0xffffffffu, a0, 1);
    if (lo) dst.st(0, a0);                           // lanes 16..31 hold the same value: one writer per address (racecheck-clean)
    c0 = fma(-a0, l1, a[1]);                         // column 1, complete
  }
  // col = &A[j-1][j-1]; a0 = L[r][j-1] is the column whose update of columns > j is still pending
#pragma unroll 1
  for (int j = 1; j < 16; j++) {
    __syncwarp();                                    // column j - 1 of the tile is visible
    double tcol[16];
#pragma unroll
    for (int k = 2; k < 16; k++) tcol[k] = col.ld(k * TLD);
    col.advance(TLD + 1);
    const double p0 = a0;
    ajj = nxt;
    // pending update of column j - 1 on columns j + 1 .. (column j already has it), fused with the slide of the window:
    // A[r][j-1+k] -= L[r][j-1] L[j-1+k][j-1], afterwards a[m] = A[r][j + m]
#pragma unroll
    for (int k = 2; k < 16; k++) a[k - 1] = fma(-p0, tcol[k], a[k]);
    a[15] = 0.0;
    pivot(c0, ajj, j, a0, nxt);
    const double l1 = __shfl_sync(0xffffffffu, a0, (j + 1) & 15);   // L[j+1][j]
    if (lo) dst.st(j, a0);
    c0 = fma(-a0, l1, a[1]);                         // the next pivot column, complete
  }
#else
  // col = &A[j][j]
#pragma unroll 1
  for (int j = 0; j < 16; j++) {
    const double ajj = __shfl_sync(0xffffffffu, a[0], j);
    bad |= !(ajj > 0.0);                             // also catches NaN
    const double di = rsqrt(ajj);
    const double a0 = a[0] * di;                     // L[r][j] (rows r < j: don't-care)
    if (lo) dst.st(j, a0);                           // lanes 16..31 hold the same value: one writer per address (racecheck-clean)
    mydi = (lane == j) ? di : mydi;
    __syncwarp();
#pragma unroll
    for (int k = 1; k < 16; k++) a[k - 1] = fma(-a0, col.ld(k * TLD), a[k]);   // A[r][j+k] -= L[r][j] L[j+k][j]; window slides by one
    a[15] = 0.0;
    col.advance(TLD + 1);
  }
#endif
  if (lo) dinv.st(lane, mydi);
  if (bad && lane == 0) *flag = 1;
}

// Inverse of a factored diagonal tile by one warp: lane c < 16 owns column c of X = L^-1 (forward substitution).
__device__ __noinline__ void chol_diag_inverse(const double* Akk, const double* dinv, double* Li) {
  const int lane = threadIdx.x & 31;
  if (lane < 16) {
    double z[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
      double sacc = (i == lane) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < 16; k++) if (k < i) sacc = fma(-Akk[i * TLD + k], z[k], sacc);
      z[i] = (i >= lane) ? sacc * dinv[i] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 16; i++) Li[i * 16 + lane] = z[i];
  }
}

// =================================================================================================================
// C: blocked Cholesky of the tile-packed lower matrix, with b (= -g) carried along as an extra row.
// On exit: H holds L, linv[kb] the inverse of each diagonal tile, y = L^-1 b in `b`.  Returns false on breakdown.
// =================================================================================================================
// The same factorisation for a reduced system that does NOT fit shared memory (20-keyframe windows, D = 307: 210 tiles = 460 KB, kept in the
// window's L2 scratch).  Run straight on global memory every operand of the trailing update was an L2 round trip (128 loads per 4x4 register
// tile) and the pivot chain of the diagonal tile an L2 round trip per column: 45 k cycles per tile row.  Here shared memory holds what a tile
// row re-reads: the diagonal tile being factored (two buffers: this row's and the next one's) and the panel column L(kb+1.., kb) once it is
// computed; global memory sees every tile once per row (panel rows in, L rows out, one read-modify-write of each trailing tile).
// stage: shared memory, (nb + 1) tiles.
__device__ bool cholesky_tiles_staged(double* H, double* b, double* linv, double* dinv, int nb, int* flag, double* stage) {
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int dw = SOLVE_WARPS - 1;
  double* dcur = stage; double* dnext = stage + TSZ; double* psm = stage + 2 * TSZ;   // psm tile j = L(kb + 1 + j, kb)
  for (int e = t; e < 16 * TLD; e += blockDim.x) dcur[e] = H[e];
  __syncthreads();
  if (warp == dw) chol_diag_factor<true>(dcur, dinv, flag, lane);
  __syncthreads();
  for (int kb = 0; kb < nb; kb++) {
    double* Akk_g = H + (size_t)(tri(kb) + kb) * TSZ;
    for (int e = t; e < 16 * TLD; e += blockDim.x) Akk_g[e] = dcur[e];          // the factored diagonal tile goes back (back-substitution, dogleg)
    // panel: rows of tiles (ib, kb), ib > kb, and the b row, times Lkk^-T (Lkk in shared memory); L rows to global memory AND to psm
    const int nrows = (nb - kb - 1) * 16 + 1;
    for (int rr = t; rr < nrows; rr += blockDim.x) {
      const bool brow = rr == nrows - 1;
      double* row = brow ? b + kb * 16 : H + (size_t)(tri(kb + 1 + rr / 16) + kb) * TSZ + (rr & 15) * TLD;
      const double* dk = dinv + kb * 16;
      double v[16];
#pragma unroll
      for (int m = 0; m < 16; m++) v[m] = row[m];
#pragma unroll
      for (int m = 0; m < 16; m++) {
        v[m] *= dk[m];
#pragma unroll
        for (int c = m + 1; c < 16; c++) v[c] = fma(-v[m], dcur[c * TLD + m], v[c]);
      }
#pragma unroll
      for (int c = 0; c < 16; c++) row[c] = v[c];
      if (!brow) {
        double* ps = psm + (size_t)(rr / 16) * TSZ + (rr & 15) * TLD;
#pragma unroll
        for (int c = 0; c < 16; c++) ps[c] = v[c];
      }
    }
    __syncthreads();
    const int rem = nb - kb - 1;
    if (rem == 0) break;
    // phase A: the next diagonal tile, updated, into its shared-memory buffer
    if (t < 256) {
      const int i = t >> 4, j = t & 15;
      double val = 0.0;
      if (i >= j) {
        const double* Li = psm + i * TLD; const double* Lj = psm + j * TLD;
        double s0 = 0, s1 = 0;
#pragma unroll
        for (int m = 0; m < 16; m += 2) { s0 = fma(Li[m], Lj[m], s0); s1 = fma(Li[m + 1], Lj[m + 1], s1); }
        val = H[(size_t)(tri(kb + 1) + kb + 1) * TSZ + i * TLD + j] - (s0 + s1);
      }
      dnext[i * TLD + j] = val;
    }
    __syncthreads();
    // phase B: diagonal warp | trailing tiles and b row | inverse of this row's diagonal tile
    if (warp == dw) {
      chol_diag_factor<true>(dnext, dinv + (kb + 1) * 16, flag, lane);
    } else if ((warp & 3) != (dw & 3)) {
      const int ntile = tri(rem) - 1, ntl = ntile * 16 + rem * 16, tu = (warp - (warp >> 2)) * 32 + lane;
      for (int it = tu; it < ntl; it += (SOLVE_WARPS - SOLVE_WARPS / 4) * 32) {
        if (it < ntile * 16) {
          const int tl = (it >> 4) + 1, sub = it & 15;
          int bi = (int)((sqrtf(8.0f * tl + 1.0f) - 1.0f) * 0.5f);
          while (bi * (bi + 1) / 2 > tl) bi--;
          while ((bi + 1) * (bi + 2) / 2 <= tl) bi++;
          const int bj = tl - bi * (bi + 1) / 2;
          const double* Lik = psm + (size_t)bi * TSZ;
          const double* Ljk = psm + (size_t)bj * TSZ;
          double* Aij = H + (size_t)(tri(kb + 1 + bi) + kb + 1 + bj) * TSZ;
          const int r0 = (sub >> 2) * 4, c0 = (sub & 3) * 4;
          double acc[4][4], old[4][4];
#pragma unroll
          for (int a = 0; a < 4; a++)
#pragma unroll
            for (int c = 0; c < 4; c++) { acc[a][c] = 0; old[a][c] = Aij[(r0 + a) * TLD + c0 + c]; }    // in flight during the products
#pragma unroll 4
          for (int m = 0; m < 16; m++) {
            double av[4], bv[4];
#pragma unroll
            for (int a = 0; a < 4; a++) av[a] = Lik[(r0 + a) * TLD + m];
#pragma unroll
            for (int c = 0; c < 4; c++) bv[c] = Ljk[(c0 + c) * TLD + m];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
              for (int c = 0; c < 4; c++) acc[a][c] = fma(av[a], bv[c], acc[a][c]);
          }
#pragma unroll
          for (int a = 0; a < 4; a++)
#pragma unroll
            for (int c = 0; c < 4; c++) Aij[(r0 + a) * TLD + c0 + c] = old[a][c] - acc[a][c];
        } else {
          const int q = it - ntile * 16, j = q >> 4, c = q & 15;
          const double* Ljk = psm + (size_t)j * TSZ;
          const double* yk = b + kb * 16;
          double s0 = 0, s1 = 0;
#pragma unroll
          for (int m = 0; m < 16; m += 2) { s0 = fma(yk[m], Ljk[c * TLD + m], s0); s1 = fma(yk[m + 1], Ljk[c * TLD + m + 1], s1); }
          b[(kb + 1 + j) * 16 + c] -= s0 + s1;
        }
      }
    } else if (warp == (dw & 3)) {
      chol_diag_inverse(dcur, dinv + kb * 16, linv + kb * 256);
    }
    __syncthreads();
    double* sw = dcur; dcur = dnext; dnext = sw;
  }
  if (warp == 0) chol_diag_inverse(dcur, dinv + (nb - 1) * 16, linv + (nb - 1) * 256);
  __syncthreads();
  return *flag == 0;
}

template <bool SH>
__device__ bool cholesky_tiles(double* H, double* b, double* linv, double* dinv, int nb, int* flag, long long* prof = nullptr, double* stage = nullptr) {
  if (!SH && stage) return cholesky_tiles_staged(H, b, linv, dinv, nb, flag, stage);   // H in the L2 scratch: operands staged in shared memory
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  long long pt = (prof && blockIdx.x == 0 && t == 0) ? clock64() : 0;
#define CPROF(i) do { if (prof && blockIdx.x == 0 && t == 0) { const long long n_ = prof_clock(); prof[i] += n_ - pt; pt = n_; } } while (0)
  // Schedule of one tile row kb (tools/chol_micro.cu measures it in isolation):
  //   panel    all threads: rows of the tiles below the diagonal tile (and the b row) times Lkk^-T
  //   phase A  256 threads: ONLY the next diagonal tile takes its update (one entry per thread, 16 FMAs)
  //   phase B  the highest warp factors the next diagonal tile (the serial chain of the whole factorisation) WHILE the warps of the other three
  //            SM sub-partitions update every other trailing tile and the b row, and the lowest warp of the diagonal warp's own sub-partition
  //            inverts the diagonal tile of this row (only the back-substitution needs it).
  // The serial diagonal tile runs on the HIGHEST warp id: the SMSP arbiter favours high warp ids (B300_MICROARCH.md), and the DFMA-saturating
  // update warps stay off its sub-partition (next to them the chain runs twice as long: tools/diag_micro.cu).
  const int dw = SOLVE_WARPS - 1;
  if (warp == dw) chol_diag_factor<SH>(H, dinv, flag, lane);
  CPROF(12);
  __syncthreads();
  for (int kb = 0; kb < nb; kb++) {
    double* Akk = H + (size_t)(tri(kb) + kb) * TSZ;
    CPROF(13);
    // panel: rows of tiles (ib, kb), ib > kb, and the b row: solve X Lkk^T = A by column-oriented substitution
    // (one thread per row; after x[m] is known every later entry is updated independently -> depth 16 x (mul + fma)).
    const int nrows = (nb - kb - 1) * 16 + 1;
    for (int rr = t; rr < nrows; rr += blockDim.x) {
      double* row = (rr == nrows - 1) ? b + kb * 16 : H + (size_t)(tri(kb + 1 + rr / 16) + kb) * TSZ + (rr & 15) * TLD;
      const double* dk = dinv + kb * 16;
      double v[16];
#pragma unroll
      for (int m = 0; m < 16; m++) v[m] = row[m];
#pragma unroll
      for (int m = 0; m < 16; m++) {
        v[m] *= dk[m];
#pragma unroll
        for (int c = m + 1; c < 16; c++) v[c] = fma(-v[m], Akk[c * TLD + m], v[c]);
      }
#pragma unroll
      for (int c = 0; c < 16; c++) row[c] = v[c];
    }
    __syncthreads();
    CPROF(14);
    const int rem = nb - kb - 1;
    if (rem == 0) break;
    // phase A: A(kb+1, kb+1) -= L(kb+1, kb) L(kb+1, kb)^T, lower triangle, one entry per thread
    if (t < 256) {
      const int i = t >> 4, j = t & 15;
      if (i >= j) {
        const double* Li = H + (size_t)(tri(kb + 1) + kb) * TSZ + i * TLD;
        const double* Lj = H + (size_t)(tri(kb + 1) + kb) * TSZ + j * TLD;
        double s0 = 0, s1 = 0;
#pragma unroll
        for (int m = 0; m < 16; m += 2) { s0 = fma(Li[m], Lj[m], s0); s1 = fma(Li[m + 1], Lj[m + 1], s1); }
        H[(size_t)(tri(kb + 1) + kb + 1) * TSZ + i * TLD + j] -= s0 + s1;
      }
    }
    __syncthreads();
    CPROF(15);
    // phase B
    if (warp == dw) {
      chol_diag_factor<SH>(H + (size_t)(tri(kb + 1) + kb + 1) * TSZ, dinv + (kb + 1) * 16, flag, lane);
    } else if ((warp & 3) != (dw & 3)) {
      // trailing update A(ib,jb) -= L(ib,kb) L(jb,kb)^T of every tile but the next diagonal one, 4x4 register tiles (16 threads per tile),
      // then the b row: b[jb] -= L(jb,kb) y_kb
      const int ntile = tri(rem) - 1, ntl = ntile * 16 + rem * 16, tu = (warp - (warp >> 2)) * 32 + lane;
      for (int it = tu; it < ntl; it += (SOLVE_WARPS - SOLVE_WARPS / 4) * 32) {
        if (it < ntile * 16) {
          const int tl = (it >> 4) + 1, sub = it & 15;
          int bi = (int)((sqrtf(8.0f * tl + 1.0f) - 1.0f) * 0.5f);
          while (bi * (bi + 1) / 2 > tl) bi--;
          while ((bi + 1) * (bi + 2) / 2 <= tl) bi++;
          const int bj = tl - bi * (bi + 1) / 2;
          const int ib = kb + 1 + bi, jb = kb + 1 + bj;
          const double* Lik = H + (size_t)(tri(ib) + kb) * TSZ;
          const double* Ljk = H + (size_t)(tri(jb) + kb) * TSZ;
          double* Aij = H + (size_t)(tri(ib) + jb) * TSZ;
          const int r0 = (sub >> 2) * 4, c0 = (sub & 3) * 4;
          double acc[4][4];
#pragma unroll
          for (int a = 0; a < 4; a++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[a][c] = 0;
#pragma unroll 4
          for (int m = 0; m < 16; m++) {
            double av[4], bv[4];
#pragma unroll
            for (int a = 0; a < 4; a++) av[a] = Lik[(r0 + a) * TLD + m];
#pragma unroll
            for (int c = 0; c < 4; c++) bv[c] = Ljk[(c0 + c) * TLD + m];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
              for (int c = 0; c < 4; c++) acc[a][c] = fma(av[a], bv[c], acc[a][c]);
          }
#pragma unroll
          for (int a = 0; a < 4; a++)
#pragma unroll
            for (int c = 0; c < 4; c++) Aij[(r0 + a) * TLD + c0 + c] -= acc[a][c];
        } else {
          const int q = it - ntile * 16, jb = kb + 1 + (q >> 4), c = q & 15;
          const double* Ljk = H + (size_t)(tri(jb) + kb) * TSZ;
          const double* yk = b + kb * 16;
          double s0 = 0, s1 = 0;
#pragma unroll
          for (int m = 0; m < 16; m += 2) { s0 = fma(yk[m], Ljk[c * TLD + m], s0); s1 = fma(yk[m + 1], Ljk[c * TLD + m + 1], s1); }
          b[jb * 16 + c] -= s0 + s1;
        }
      }
    } else if (warp == (dw & 3)) {
      chol_diag_inverse(Akk, dinv + kb * 16, linv + kb * 256);
    }
    CPROF(12);
    __syncthreads();
  }
  if (warp == 0) chol_diag_inverse(H + (size_t)(tri(nb - 1) + nb - 1) * TSZ, dinv + (nb - 1) * 16, linv + (nb - 1) * 256);
  __syncthreads();
  CPROF(12);
#undef CPROF
  return *flag == 0;
}

// Back-substitution L^T x = y (y in `b`, x written to `dx`), blocked with the stored tile inverses, right-looking and in place in dx:
// x_kb = Lkk^-T dx_kb by 16 threads, then every remaining entry takes its update dx[jb*16 + c] -= sum_r L(kb,jb)[r][c] x_kb[r] on its own thread
// (16 independent FMAs per thread, no warp reductions): two barriers and two ~16-deep FMA chains per tile row.  The left-looking form it
// replaces (a warp reduction per column over all rows below) took 23.6 k cycles per solve of the 157-dimensional system, this one see
// tools/chol_micro.cu.
__device__ void backsub_tiles(const double* H, const double* b, const double* linv, double* dx, int nb, double* /*tmp*/) {
  const int t = threadIdx.x;
  for (int e = t; e < nb * 16; e += blockDim.x) dx[e] = b[e];
  __syncthreads();
  for (int kb = nb - 1; kb >= 0; kb--) {
    if (t < 32) {   // x[c] = sum_{m>=c} Linv[m][c] y[m]; lanes 16..31 mirror and do not write
      const int c = t & 15;
      const double* Li = linv + kb * 256;
      double s0 = 0, s1 = 0;
#pragma unroll
      for (int m = 0; m < 16; m += 2) {
        s0 = fma(m >= c ? Li[m * 16 + c] : 0.0, dx[kb * 16 + m], s0);
        s1 = fma(m + 1 >= c ? Li[(m + 1) * 16 + c] : 0.0, dx[kb * 16 + m + 1], s1);
      }
      __syncwarp();
      if (t < 16) dx[kb * 16 + c] = s0 + s1;
    }
    __syncthreads();
    for (int o = t; o < kb * 16; o += blockDim.x) {
      const int jb = o >> 4, c = o & 15;
      const double* T = H + (size_t)(tri(kb) + jb) * TSZ + c;
      const double* xk = dx + kb * 16;
      double s0 = 0, s1 = 0;
#pragma unroll
      for (int r = 0; r < 16; r += 2) { s0 = fma(T[r * TLD], xk[r], s0); s1 = fma(T[(r + 1) * TLD], xk[r + 1], s1); }
      dx[o] -= s0 + s1;
    }
    __syncthreads();
  }
}

}  // namespace vb
