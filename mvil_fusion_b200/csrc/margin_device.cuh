// margin_device.cuh — marginalization of one window on the device:
// Estimator::optimization() tail (vils_estimator/src/estimator.cpp:1483-1684) + MarginalizationInfo::{preMarginalize,
// marginalize, getParameterBlocks} (factor/marginalization_factor.cpp:110-129,176-338).
//
//   MARGIN_OLD       : prior, last ICP / LPS constraint starting at frame 0, IMU(0,1), every projection factor anchored at
//                      frame 0; drops pose0, speed-bias0 and those landmarks.
//   MARGIN_SECOND_NEW: prior only, drops pose[N-2].
// A = sum J^T J, b = sum J^T r with the robust corrector applied (ResidualBlockInfo::Evaluate, :37-68); Amm pseudo-inverse
// and the final factorisation A' = V S V^T use a parallel cyclic Jacobi eigen-solver (stands in for
// Eigen::SelfAdjointEigenSolver, :275,301) with the reference's eps = 1e-8 cut (:70).
// Block order is canonical (the reference's unordered_map order is not reproducible): dropped = pose0, sb0 / pose[N-2],
// landmarks by feature index; kept = poses, speed-biases, ex, td — each only if a participating factor touches it.
// One CTA; matrices live in a global workspace (L2 resident).
#pragma once
#include "ba_device.cuh"

namespace vb {

struct MargParams {
  int32_t slot, flag;
  double* ws;            // workspace (doubles)
  int64_t oH, oG, oA, oB, oV, oW, oAinv, oArm, oAr, oBr, oV2, oS, oStage, oJout, oRout, oX0;   // offsets into ws
  int32_t* iws;          // int workspace: [0]=n [1]=m [2]=nblk [3]=pos [4]=Tp ; touched[Tcap] drop[Tcap] order[Tcap] lmloc[Mcap] blk[64] idxA[..] idxB[..]
  int32_t Tcap, Mcap;
};

// Parallel cyclic Jacobi for a symmetric n x n matrix A (row-major, ld = n, destroyed); eigenvalues -> w, eigenvectors -> columns of V.
__device__ void jacobi_eigh(double* A, int n, double* w, double* V, int* top, int* bot, double* cs, double* red) {
  const int t = threadIdx.x, T = blockDim.x;
  const int np = (n + 1) & ~1, half = np / 2;
  for (int e = t; e < n * n; e += T) V[e] = (e / n == e % n) ? 1.0 : 0.0;
  for (int k = t; k < half; k += T) { top[k] = k; bot[k] = half + k; }
  __syncthreads();
  for (int sweep = 0; sweep < 40; sweep++) {
    double off = 0, dg = 0;
    for (int e = t; e < n * n; e += T) { const int i = e / n, j = e % n; const double v = A[e]; if (i == j) dg += v * v; else off += v * v; }
    off = block_sum(off, red); dg = block_sum(dg, red);
    if (off <= 1e-40 * (dg + 1e-300) || off == 0.0) break;
    for (int step = 0; step < np - 1; step++) {
      for (int k = t; k < half; k += T) {
        int p = min(top[k], bot[k]), q = max(top[k], bot[k]);
        double c = 1.0, s = 0.0;
        if (q < n) {
          const double apq = A[p * n + q];
          if (apq != 0.0) {
            const double tau = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
            const double tt = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + tt * tt); s = tt * c;
          }
        }
        cs[2 * k] = c; cs[2 * k + 1] = s;
      }
      __syncthreads();
      for (int e = t; e < n * half; e += T) {          // columns: A <- A R, V <- V R
        const int i = e / half, k = e % half;
        const int p = min(top[k], bot[k]), q = max(top[k], bot[k]);
        if (q >= n) continue;
        const double c = cs[2 * k], s = cs[2 * k + 1];
        if (s == 0.0) continue;
        const double aip = A[i * n + p], aiq = A[i * n + q];
        A[i * n + p] = c * aip - s * aiq; A[i * n + q] = s * aip + c * aiq;
        const double vip = V[i * n + p], viq = V[i * n + q];
        V[i * n + p] = c * vip - s * viq; V[i * n + q] = s * vip + c * viq;
      }
      __syncthreads();
      for (int e = t; e < n * half; e += T) {          // rows: A <- R^T A
        const int j = e / half, k = e % half;
        const int p = min(top[k], bot[k]), q = max(top[k], bot[k]);
        if (q >= n) continue;
        const double c = cs[2 * k], s = cs[2 * k + 1];
        if (s == 0.0) continue;
        const double apj = A[p * n + j], aqj = A[q * n + j];
        A[p * n + j] = c * apj - s * aqj; A[q * n + j] = s * apj + c * aqj;
      }
      __syncthreads();
      if (t == 0) {   // round-robin tournament: top[0] fixed, everything else rotates
        const int last_top = top[half - 1], first_bot = bot[0];
        for (int k = half - 1; k > 1; k--) top[k] = top[k - 1];
        if (half > 1) top[1] = first_bot;
        for (int k = 0; k < half - 1; k++) bot[k] = bot[k + 1];
        bot[half - 1] = last_top;
        if (half == 1) { /* single pair: nothing to rotate */ bot[0] = first_bot; }
      }
      __syncthreads();
    }
  }
  for (int i = t; i < n; i += T) w[i] = A[i * n + i];
  __syncthreads();
}

// H[idx[a]][idx[b]] += sum_r J[r][a] J[r][b] ; g[idx[a]] += sum_r J[r][a] res[r] for one factor (all threads, then a barrier)
__device__ void add_block(double* H, double* g, int ld, const int* idx, int width, const double* J, const double* res, int nr, int* touched) {
  for (int e = threadIdx.x; e < width * width + width; e += blockDim.x) {
    if (e < width * width) {
      const int a = e / width, b = e % width; double v = 0;
      for (int r = 0; r < nr; r++) v = fma(J[r * width + a], J[r * width + b], v);
      H[(size_t)idx[a] * ld + idx[b]] += v;
    } else {
      const int a = e - width * width; double v = 0;
      for (int r = 0; r < nr; r++) v = fma(J[r * width + a], res[r], v);
      g[idx[a]] += v; touched[idx[a]] = 1;
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(SOLVE_THREADS, 1) margin_kernel(SolveParams P, MargParams Q) {
  extern __shared__ __align__(16) double sm[];   // [0,512) reductions / small stages ; ints after
  __shared__ int s_sel[4];
  const int t = threadIdx.x, T = blockDim.x;
  const Win W = decode(P, Q.slot);
  const int N = W.N, D = W.D;
  const double* x = P.xout + (size_t)Q.slot * P.xout_stride;      // solved state
  const double* scr = P.scratch + (size_t)Q.slot * P.sl.total;
  double* red = sm; double* stage = sm + 64;                       // stage: >= 466 doubles
  int* idx = reinterpret_cast<int*>(sm + 64 + 480);                // 64 ints
  int* touched = Q.iws + 8; int* drop = touched + Q.Tcap; int* order = drop + Q.Tcap; int* lmloc = order + Q.Tcap; int* blkout = lmloc + Q.Mcap;
  int* top = blkout + 2 * (2 * P.Ncap + 2); int* bot = top + Q.Tcap;
  double* H = Q.ws + Q.oH; double* g = Q.ws + Q.oG;
  // ---- dropped landmarks (MARGIN_OLD): those anchored at frame 0, local index D + k in feature order
  const int32_t* ix = W.i(OFF_PROJ_IDX); const int32_t* lm_start = W.i(OFF_LM_START); const int32_t* lm_feat = W.i(OFF_LM_FEAT);
  if (t == 0) {
    int m0 = 0;
    for (int r = 0; r < W.h->n_lm; r++) lmloc[r] = (Q.flag == VILS_MARGIN_OLD && ix[lm_start[r]] == 0) ? m0++ : -1;
    s_sel[0] = m0;
  }
  __syncthreads();
  const int M0 = s_sel[0], Tp = D + M0;
  for (int e = t; e < Tp * Tp; e += T) H[e] = 0.0;
  for (int e = t; e < Tp; e += T) { g[e] = 0.0; touched[e] = 0; drop[e] = 0; }
  __syncthreads();
  // ---- prior (both flags)
  const int pn = W.h->prior_n;
  bool has_second = false;
  if (pn > 0) {
    const int32_t* blk = W.i(OFF_PRIOR_BLK); const int32_t* col = W.i(OFF_PRIOR_COL);
    for (int b = 0; b < W.h->prior_nblk; b++) if (blk[4 * b] == VILS_BLK_POSE && blk[4 * b + 1] == N - 2) has_second = true;
    if (Q.flag == VILS_MARGIN_OLD || has_second) {
      double* dxp = Q.ws + Q.oBr; double* rp = Q.ws + Q.oS;      // scratch vectors (free until the Schur step)
      const double* x0 = W.d(OFF_PRIOR_X0); const double* Jl = W.d(OFF_PRIOR_J); const double* rl = W.d(OFF_PRIOR_R);
      for (int b = t; b < W.h->prior_nblk; b += T) {
        const int type = blk[4 * b], bi = blk[4 * b + 1], xo = blk[4 * b + 2], c0 = blk[4 * b + 3];
        const double* xb = type == VILS_BLK_POSE ? x + XP(bi) : type == VILS_BLK_SPEEDBIAS ? x + XS(N, bi) : type == VILS_BLK_EXPOSE ? x + XE(N) : x + XT(N);
        if (type == VILS_BLK_POSE || type == VILS_BLK_EXPOSE) vf::prior_dx_pose(xb, x0 + xo, dxp + c0);
        else { const int sz = type == VILS_BLK_SPEEDBIAS ? 9 : 1; for (int k = 0; k < sz; k++) dxp[c0 + k] = xb[k] - x0[xo + k]; }
      }
      __syncthreads();
      for (int i = t; i < pn; i += T) { double r = rl[i]; for (int j = 0; j < pn; j++) r = fma(Jl[(size_t)j * pn + i], dxp[j], r); rp[i] = r; }
      __syncthreads();
      const double* A = scr + P.sl.priorA;
      for (int e = t; e < pn * pn + pn; e += T) {
        if (e < pn * pn) { const int a = e / pn, b = e % pn; H[(size_t)col[a] * Tp + col[b]] += A[e]; }
        else { const int a = e - pn * pn; double v = 0; for (int i = 0; i < pn; i++) v = fma(Jl[(size_t)a * pn + i], rp[i], v); g[col[a]] += v; touched[col[a]] = 1; }
      }
      __syncthreads();
    }
  }
  if (Q.flag == VILS_MARGIN_OLD) {
    // ---- the LAST ICP / LPS constraint whose first keyframe is frame 0 (ICPmarg / LPSmarg are overwritten in the loops,
    //      estimator.cpp:1311-1317,1381-1389)
    if (t == 0) {
      int is = -1, ls = -1;
      const double* icp = W.d(OFF_ICP); const double* lps = W.d(OFF_LPS);
      for (int k = 0; k < W.h->n_icp; k++) if ((int)icp[14 * k + 10] == 0) is = k;
      for (int k = 0; k < W.h->n_lps; k++) if ((int)lps[9 * k + 7] == 0) ls = k;
      s_sel[1] = is; s_sel[2] = ls;
    }
    __syncthreads();
    if (s_sel[1] >= 0) {
      const double* c = W.d(OFF_ICP) + 14 * s_sel[1];
      if (t == 0) {
        vf::icp_eval(c, x + XP((int)c[10]), x + XP((int)c[11]), x + XP((int)c[12]), x + XP((int)c[13]), stage, stage + 3);
        double rho, w; vf::cauchy(P.cfg.cauchy_a, stage[0] * stage[0] + stage[1] * stage[1] + stage[2] * stage[2], rho, w);
        for (int k = 0; k < 75; k++) stage[k] *= w;
        for (int a = 0; a < 24; a++) idx[a] = 15 * (int)c[10 + a / 6] + a % 6;
      }
      __syncthreads();
      add_block(H, g, Tp, idx, 24, stage + 3, stage, 3, touched);
    }
    if (s_sel[2] >= 0) {
      const double* c = W.d(OFF_LPS) + 9 * s_sel[2];
      if (t == 0) {
        vf::lps_eval(c, x + XP((int)c[7]), x + XP((int)c[8]), stage, stage + 3);
        double rho, w; vf::cauchy(P.cfg.cauchy_a, stage[0] * stage[0] + stage[1] * stage[1] + stage[2] * stage[2], rho, w);
        for (int k = 0; k < 39; k++) stage[k] *= w;
        for (int a = 0; a < 12; a++) idx[a] = 15 * (int)c[7 + a / 6] + a % 6;
      }
      __syncthreads();
      add_block(H, g, Tp, idx, 12, stage + 3, stage, 3, touched);
    }
    // ---- IMU factor(s) starting at frame 0 (estimator.cpp:1536-1543)
    for (int k = 0; k < W.h->n_imu; k++) {
      if (W.i(OFF_IMU_KF)[k] != 0) continue;
      const double* pre = W.d(OFF_IMU) + (size_t)k * 467;
      if (!(pre[16] < 10.0)) continue;
      double* J = stage; double* r = stage + 450;
      for (int e = t; e < 450; e += T) J[e] = 0.0;
      __syncthreads();
      if (t == 0) vf::imu_eval_raw(pre, P.cfg.G, x + XP(0), x + XS(N, 0), x + XP(1), x + XS(N, 1), r, J);
      __syncthreads();
      const double* Wk = scr + P.sl.w_imu + (size_t)k * 225;
      double* Jw = Q.ws + Q.oStage;     // 15 x 30 + 15
      for (int e = t; e < 450 + 15; e += T) {
        if (e < 450) { const int a = e / 30, cc = e % 30; double v = 0; for (int m2 = a; m2 < 15; m2++) v = fma(Wk[a * 15 + m2], J[m2 * 30 + cc], v); Jw[e] = v; }
        else { const int a = e - 450; double v = 0; for (int m2 = a; m2 < 15; m2++) v = fma(Wk[a * 15 + m2], r[m2], v); Jw[450 + a] = v; }
      }
      if (t < 30) idx[t] = t;            // pose0 sb0 pose1 sb1 = camera dims 0..29
      __syncthreads();
      add_block(H, g, Tp, idx, 30, Jw, Jw + 450, 15, touched);
    }
    // ---- projection factors anchored at frame 0 (estimator.cpp:1547-1588): evaluate all, then add one after the other
    const int np = W.h->n_proj; const double* c0 = W.d(OFF_PROJ);
    double* pst = Q.ws + Q.oStage;        // 44 doubles per factor: r(2) J(2x20) pad(2)
    for (int f = t; f < np; f += T) {
      if (ix[f] != 0) continue;
      double c[14];
      for (int k = 0; k < 14; k++) c[k] = c0[(size_t)k * np + f];
      const int feat = lm_feat[ix[2 * np + f]];
      double* o = pst + (size_t)f * 44;
      vf::proj_eval(P.cfg, c, x + XP(0), x + XP(ix[np + f]), x + XE(N), x[XL(N) + feat], x[XT(N)], o, o + 2);
      double rho, w; vf::cauchy(P.cfg.cauchy_a, o[0] * o[0] + o[1] * o[1], rho, w);
      for (int k = 0; k < 42; k++) o[k] *= w;
    }
    __syncthreads();
    for (int f = 0; f < np; f++) {
      if (ix[f] != 0) continue;
      const int j = ix[np + f], loc = lmloc[ix[2 * np + f]];
      if (t < 20) idx[t] = t < 6 ? t : t < 12 ? 15 * j + (t - 6) : t < 18 ? 15 * N + (t - 12) : t == 18 ? D + loc : 15 * N + 6;
      if (t == 0) drop[D + loc] = 1;
      __syncthreads();
      add_block(H, g, Tp, idx, 20, pst + (size_t)f * 44 + 2, pst + (size_t)f * 44, 2, touched);
    }
    for (int e = t; e < 15; e += T) drop[e] = 1;
  } else {
    for (int e = t; e < 6; e += T) drop[15 * (N - 2) + e] = 1;
  }
  __syncthreads();
  // ---- ordering [drop ; keep]
  if (t == 0) {
    if (!P.cfg.use_td) touched[15 * N + 6] = 0;
    int pos = 0, nb = 0;
    if (Q.flag == VILS_MARGIN_SECOND_NEW && !has_second) { Q.iws[0] = 0; Q.iws[1] = 0; Q.iws[2] = 0; Q.iws[3] = 0; }
    else {
      for (int i = 0; i < Tp; i++) if (touched[i] && drop[i]) order[pos++] = i;
      const int m = pos;
      auto keep = [&](int type, int bi, int off, int ls) {
        if (!touched[off] || drop[off]) return;
        int ni = bi;
        if (type == VILS_BLK_POSE || type == VILS_BLK_SPEEDBIAS) ni = (Q.flag == VILS_MARGIN_OLD) ? bi - 1 : (bi == N - 1 ? N - 2 : bi);   // addr_shift (estimator.cpp:1599-1611,1654-1675)
        blkout[2 * nb] = VILS_BLK_ID(type, ni); blkout[2 * nb + 1] = VILS_BLK_ID(type, bi); nb++;
        for (int a = 0; a < ls; a++) order[pos++] = off + a;
      };
      for (int k = 0; k < N; k++) keep(VILS_BLK_POSE, k, 15 * k, 6);
      for (int k = 0; k < N; k++) keep(VILS_BLK_SPEEDBIAS, k, 15 * k + 6, 9);
      keep(VILS_BLK_EXPOSE, 0, 15 * N, 6);
      keep(VILS_BLK_TD, 0, 15 * N + 6, 1);
      Q.iws[0] = pos - m; Q.iws[1] = m; Q.iws[2] = nb; Q.iws[3] = pos;
    }
  }
  __syncthreads();
  const int n = Q.iws[0], m = Q.iws[1], pos = Q.iws[3];
  if (pos == 0) return;
  double* A = Q.ws + Q.oA; double* bv = Q.ws + Q.oB;
  for (int e = t; e < pos * pos; e += T) A[e] = H[(size_t)order[e / pos] * Tp + order[e % pos]];
  for (int e = t; e < pos; e += T) bv[e] = g[order[e]];
  __syncthreads();
  // ---- Amm^+ (marginalization_factor.cpp:274-279)
  double* Amm = Q.ws + Q.oV2;          // reuse as the matrix that Jacobi destroys
  double* V = Q.ws + Q.oV; double* wv = Q.ws + Q.oW; double* Ainv = Q.ws + Q.oAinv;
  double* cs = sm + 64;                 // 2 * half doubles (<= 448)
  const double eps = 1e-8;
  if (m > 0) {
    for (int e = t; e < m * m; e += T) { const int i = e / m, j = e % m; Amm[e] = 0.5 * (A[(size_t)i * pos + j] + A[(size_t)j * pos + i]); }
    __syncthreads();
    jacobi_eigh(Amm, m, wv, V, top, bot, cs, red);
    for (int e = t; e < m * m; e += T) {
      const int i = e / m, j = e % m; double v = 0;
      for (int k = 0; k < m; k++) if (wv[k] > eps) v += V[i * m + k] * (1.0 / wv[k]) * V[j * m + k];
      Ainv[e] = v;
    }
    __syncthreads();
  }
  // ---- Schur (:282-290)
  double* Arm = Q.ws + Q.oArm; double* Ar = Q.ws + Q.oAr; double* br = Q.ws + Q.oBr;
  for (int e = t; e < n * m; e += T) { const int i = e / m, j = e % m; double v = 0; for (int k = 0; k < m; k++) v = fma(A[(size_t)(m + i) * pos + k], Ainv[k * m + j], v); Arm[e] = v; }
  __syncthreads();
  for (int e = t; e < n * n + n; e += T) {
    if (e < n * n) { const int i = e / n, j = e % n; double v = A[(size_t)(m + i) * pos + m + j]; for (int k = 0; k < m; k++) v -= Arm[i * m + k] * A[(size_t)k * pos + m + j]; Ar[e] = v; }
    else { const int i = e - n * n; double v = bv[m + i]; for (int k = 0; k < m; k++) v -= Arm[i * m + k] * bv[k]; br[i] = v; }
  }
  __syncthreads();
  for (int e = t; e < n * n; e += T) { const int i = e / n, j = e % n; if (j < i) { const double v = 0.5 * (Ar[e] + Ar[j * n + i]); Ar[e] = v; } }
  __syncthreads();
  for (int e = t; e < n * n; e += T) { const int i = e / n, j = e % n; if (j > i) Ar[e] = Ar[j * n + i]; }
  __syncthreads();
  // ---- A' = V2 S V2^T -> J_lin = sqrt(S) V2^T, r_lin = sqrt(S^+) V2^T b' (:301-309); rows in ascending eigenvalue order
  double* V2 = Q.ws + Q.oV2; double* S = Q.ws + Q.oS;
  double* Awork = Q.ws + Q.oA;           // A is no longer needed
  for (int e = t; e < n * n; e += T) Awork[e] = Ar[e];
  __syncthreads();
  jacobi_eigh(Awork, n, S, V2, top, bot, cs, red);
  int* rank = order;                      // order[] is free now
  for (int k = t; k < n; k += T) { int r = 0; for (int j = 0; j < n; j++) if (S[j] < S[k] || (S[j] == S[k] && j < k)) r++; rank[k] = r; }
  __syncthreads();
  double* Jout = Q.ws + Q.oJout; double* rout = Q.ws + Q.oRout;
  for (int e = t; e < n * n + n; e += T) {
    if (e < n * n) { const int j = e / n, k = e % n; const double sv = S[k] > eps ? S[k] : 0.0; Jout[(size_t)j * n + rank[k]] = sqrt(sv) * V2[j * n + k]; }
    else { const int k = e - n * n; const double si = S[k] > eps ? 1.0 / S[k] : 0.0; double v = 0; for (int j = 0; j < n; j++) v = fma(V2[j * n + k], br[j], v); rout[rank[k]] = sqrt(si) * v; }
  }
  // ---- x0 snapshots of the kept blocks at the solved state (getParameterBlocks, :318-338)
  if (t == 0) {
    double* x0o = Q.ws + Q.oX0; int o = 0;
    for (int b = 0; b < Q.iws[2]; b++) {
      const int id = blkout[2 * b + 1], type = VILS_BLK_TYPE(id), bi = VILS_BLK_INDEX(id);
      const double* xb = type == VILS_BLK_POSE ? x + XP(bi) : type == VILS_BLK_SPEEDBIAS ? x + XS(N, bi) : type == VILS_BLK_EXPOSE ? x + XE(N) : x + XT(N);
      const int gs = (type == VILS_BLK_POSE || type == VILS_BLK_EXPOSE) ? 7 : type == VILS_BLK_SPEEDBIAS ? 9 : 1;
      for (int k = 0; k < gs; k++) x0o[o++] = xb[k];
    }
    Q.iws[4] = o;
  }
}

}  // namespace vb
