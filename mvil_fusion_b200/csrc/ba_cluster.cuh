// ba_cluster.cuh — ONE window solved by a thread-block cluster (sm_90+ / sm_100a): the latency form of the Gauss-Newton solve.
//
// solve_kernel gives one SM to a window, which is the right shape for hundreds of windows but leaves 147 SMs idle when the estimator
// solves the ONE window a new frame produces (Estimator::optimization(), the drop-in case).  Here the G CTAs of a cluster (G = 2, 4 or 8
// SMs, co-scheduled by hardware, synchronised by barrier.cluster) split every phase whose work items are independent:
//   F  factor pass      keyframe pairs p = r, r+G, ... with their projection factors (evaluation + pair-local A^T A), IMU factors k = r, r+G, ...
//                       (both stages), LiDAR factors of keyframes k = r, r+G, ... (6x6 blocks into a small buffer);
//   L  landmarks        rank = r, r+G, ...: reduction of the factor partials, then the Schur SYRK over THOSE landmarks: a partial Hv per CTA;
//   G  gather           entries t = r, r+G, ... of the visual sub-system: sum of the G partial Hv + pair blocks -> H;
//   C  solve            CTA 0: IMU / LiDAR / ICP / LPS / prior contributions, damping, blocked Cholesky, back-substitution (the serial chain);
//   U  update           every CTA applies dx to its own copy of the camera state (identical arithmetic), rank-strided landmarks are updated by
//                       their owner and exchanged.
// All exchange goes through the window's L2-resident scratch (a cluster shares nothing but L2 and DSMEM; the partial systems are a few tens of
// KB); ordering = __threadfence + barrier.cluster.arrive.release / wait.acquire.  Results agree with solve_kernel to rounding (the partial
// sums are associated differently), every accumulator still has exactly one owner: bit-reproducible run to run.
#pragma once
#include "ba_device.cuh"

__device__ void warp_imu_sqrt_info(const double* cov, double* W, double* wk);   // vils_ba.cu

namespace vb {

constexpr int CL_MAX = 16;           // 8 is the portable cluster size limit; 16 is opt-in (cudaFuncAttributeNonPortableClusterSizeAllowed) and only
                                     // used when the device can co-schedule such a cluster
constexpr int CL_PORTABLE = 8;

__device__ __forceinline__ unsigned cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_size() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
// full cluster barrier that also orders the global-memory traffic of the phase before it
__device__ __forceinline__ void cluster_sync_all() {
  __threadfence();
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// prep_window split over the cluster: IMU sqrt_info factor by factor, prior A = J^T J / b0 entry-strided, zero pattern of E row-strided
__device__ void prep_window_cluster(const SolveParams& P, const Win& W, double* scr, double* work, int r, int G) {
  const int warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const double* pre = W.d(OFF_IMU);
  for (int k = r + G * warp; k < W.h->n_imu; k += G * nwarp) ::warp_imu_sqrt_info(pre + (size_t)k * 467 + 242, scr + P.sl.w_imu + (size_t)k * 225, work + warp * 450);
  const int n = W.h->prior_n;
  const double* J = W.d(OFF_PRIOR_J); const double* rl = W.d(OFF_PRIOR_R);
  for (int e = r + G * threadIdx.x; e < n * n + n; e += G * blockDim.x) {
    if (e < n * n) {
      const int a = e / n, b = e % n; double s = 0;
      for (int i = 0; i < n; i++) s = fma(J[(size_t)a * n + i], J[(size_t)b * n + i], s);
      scr[P.sl.priorA + e] = s;
    } else {
      const int a = e - n * n; double s = 0;
      for (int i = 0; i < n; i++) s = fma(J[(size_t)a * n + i], rl[i], s);
      scr[P.sl.priorb0 + a] = s;
    }
  }
  for (int64_t e = r + G * (int64_t)threadIdx.x; e < (int64_t)W.h->n_lm * P.sl.Dv_pad; e += G * blockDim.x) scr[P.sl.E + e] = 0.0;
  if (r == 0) for (int e = threadIdx.x; e < PAIR_LD * PAIR_LD; e += blockDim.x) scr[P.sl.pairpart + (int64_t)(P.Ncap * (P.Ncap - 1) / 2) * PAIR_LD * PAIR_LD + e] = 0.0;
  __syncthreads();
}


// ---- F: projection factors of the pairs owned by CTA r + IMU factors k = r (mod G) on the warps past PAIR_WARPS --------------------------
__device__ double pair_pass_cluster(const SolveParams& P, const Win& W, const double* x, double* stage, double* scr, bool need_cost, double* imu_stage,
                                    int imu_slots, const uint16_t* tbl, double* rot, int* own, int r, int G) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int np = W.h->n_proj, npair = W.h->n_pair;
  const double* c0 = W.d(OFF_PROJ); const int32_t* ix = W.i(OFF_PROJ_IDX);
  const int32_t* pairs = W.pairs(); const int32_t* perm = W.i(OFF_PAIR_PERM);
  const int32_t* lm_feat = W.lm_feat();
  const uint8_t* dfix = W.u(OFF_FIXED);
  double* part = scr + P.sl.part; double* E = scr + P.sl.E; double* pairpart = scr + P.sl.pairpart;
  const int ra = 5 * (lane >> 3), cb = 3 * (lane & 7);
  double cost = 0;
  for (int k = threadIdx.x; k <= W.N; k += blockDim.x) {
    const vm::m3 R = vm::q2R(vm::ldq((k < W.N ? x + XP(k) : x + XE(W.N)) + 3));
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int b = 0; b < 3; b++) rot[9 * k + 3 * a + b] = R.m[a][b];
  }
  // own[0] = number of own pairs, own[1 + q] = prefix of factor counts over the own pairs (q-th own pair = pair r + q G)
  if (threadIdx.x == 0) {
    int nq = 0, acc = 0; own[1] = 0;
    for (int p = r; p < npair; p += G) { acc += pairs[4 * p + 1]; nq++; own[1 + nq] = acc; }
    own[0] = nq;
  }
  __syncthreads();
  const int nq = own[0], own_np = own[1 + nq];
  const int nimu = imu_stage ? W.h->n_imu : 0;
  // own IMU factors: both stages by one warp each, while the other warps evaluate the first round of projection factors
  if (warp >= PAIR_WARPS && warp - PAIR_WARPS < imu_slots) {
    int j = 0;
    for (int k = r; k < nimu; k += G, j++) {
      if (j % imu_slots != warp - PAIR_WARPS) continue;
      double* slot = imu_stage + (size_t)(j % imu_slots) * IMU_SLOT2;
      imu_stage_R(P, W, tbl, x, k, slot);
      cost += imu_stage_P(P, W, k, slot, scr);
    }
  }
  for (int base = 0; base < own_np || base == 0; base += PAIR_CHUNK) {
    const int cnt_round = max(0, min(PAIR_CHUNK, own_np - base));
    const int t = threadIdx.x;
    if (t < cnt_round) {
      const int u = base + t;
      int q = 0; while (own[2 + q] <= u) q++;
      const int p = r + q * G;
      const int f = perm[pairs[4 * p] + (u - own[1 + q])];
      double c[14];
#pragma unroll
      for (int k = 0; k < 14; k++) c[k] = c0[(size_t)k * np + f];
      const int kfi = ix[f], kfj = ix[np + f], rank = ix[2 * np + f];
      const int feat = lm_feat[rank];
      double* row0 = stage + (2 * t) * STAGE_LD; double* row1 = row0 + STAGE_LD;
      double rr[2], J[40];
      vf::proj_eval_rows(P.cfg, c, vf::ldm(rot + 9 * kfi), vf::ldm(rot + 9 * kfj), vf::ldm(rot + 9 * W.N), vm::ld3(x + XP(kfi)), vm::ld3(x + XP(kfj)), vm::ld3(x + XE(W.N)),
                         x[XL(W.N) + feat], x[XT(W.N)], rr, J);
      double rho, w; const double s2 = rr[0] * rr[0] + rr[1] * rr[1];
      if (need_cost) vf::cauchy(P.cfg.cauchy_a, s2, rho, w); else { w = vf::cauchy_w(P.cfg.cauchy_a, s2); rho = s2; }
      cost += 0.5 * rho;
      rr[0] *= w; rr[1] *= w;
#pragma unroll
      for (int k = 0; k < 40; k++) J[k] *= w;
      const bool fixed = dfix[feat] != 0;
      const double jl0 = fixed ? 0.0 : J[18], jl1 = fixed ? 0.0 : J[38];
#pragma unroll
      for (int k = 0; k < 18; k++) { row0[k] = J[k]; row1[k] = J[20 + k]; }
      row0[18] = J[19]; row0[19] = rr[0]; row1[18] = J[39]; row1[19] = rr[1];
#pragma unroll
      for (int k = 20; k < STAGE_LD; k++) { row0[k] = 0; row1[k] = 0; }
      double* pt = part + (size_t)f * PART_LD;
      pt[0] = jl0 * jl0 + jl1 * jl1;
      pt[1] = jl0 * rr[0] + jl1 * rr[1];
#pragma unroll
      for (int k = 0; k < 6; k++) {
        pt[2 + k] = J[k] * jl0 + J[20 + k] * jl1;
        pt[8 + k] = J[12 + k] * jl0 + J[32 + k] * jl1;
        E[(size_t)rank * W.Dvp + 6 * kfj + k] = J[6 + k] * jl0 + J[26 + k] * jl1;
      }
      pt[14] = J[19] * jl0 + J[39] * jl1;
    }
    __syncthreads();
    for (int q = warp; q < nq; q += SOLVE_WARPS) {
      const int start = own[1 + q], cnt = own[2 + q] - own[1 + q];
      if (start >= base + cnt_round || start + cnt <= base) continue;
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
      double* out = pairpart + (size_t)(r + q * G) * PAIR_LD * PAIR_LD;
      const bool first = s0 == start;
#pragma unroll
      for (int a = 0; a < 5; a++)
#pragma unroll
        for (int b = 0; b < 3; b++)
          if (cb + b < PAIR_LD) { double* o = out + (ra + a) * PAIR_LD + cb + b; *o = first ? acc[a][b] : *o + acc[a][b]; }
    }
    __syncthreads();
    if (own_np == 0) break;
  }
  return cost;
}

// LiDAR plane + edge factors of the keyframes k = r (mod G): [lower 6x6 (21) | g (6)] per keyframe into blk (global), no access to H
__device__ double lidar_pass_cluster(const SolveParams& P, const Win& W, const double* x, double* blk, bool want_J, int r, int G) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int npl = W.h->n_plane, ned = W.h->n_edge;
  const double* pl = W.d(OFF_PLANE); const double* ed = W.d(OFF_EDGE);
  const int32_t* pls = W.i(OFF_PLANE_START); const int32_t* eds = W.i(OFF_EDGE_START);
  double cost = 0;
  int j = 0;
  for (int k = r; k < W.N; k += G, j++) {
    if (j % SOLVE_WARPS != warp) continue;
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
      double rs = vf::plane_eval(pose, pb, n, pl[(size_t)6 * npl + f], want_J ? J : nullptr);
      double rho, w; vf::huber(P.cfg.huber_a, rs * rs, rho, w);
      cost += 0.5 * rho;
      if (want_J) {
        rs *= w;
        int e = 0;
#pragma unroll
        for (int a = 0; a < 6; a++) { J[a] *= w; }
#pragma unroll
        for (int a = 0; a < 6; a++) { gg[a] = fma(J[a], rs, gg[a]);
#pragma unroll
          for (int b = 0; b <= a; b++) { A[e] = fma(J[a], J[b], A[e]); e++; } }
      }
    }
    if (ned) for (int f = eds[k] + lane; f < eds[k + 1]; f += 32) {
      const vm::v3 pb = vm::mk(ed[f], ed[(size_t)ned + f], ed[(size_t)2 * ned + f]);
      const vm::v3 a_ = vm::mk(ed[(size_t)3 * ned + f], ed[(size_t)4 * ned + f], ed[(size_t)5 * ned + f]);
      const vm::v3 b_ = vm::mk(ed[(size_t)6 * ned + f], ed[(size_t)7 * ned + f], ed[(size_t)8 * ned + f]);
      double rs[3], J[18];
      vf::edge_eval(pose, pb, a_, b_, rs, want_J ? J : nullptr);
      double rho, w; vf::huber(P.cfg.huber_a, rs[0] * rs[0] + rs[1] * rs[1] + rs[2] * rs[2], rho, w);
      cost += 0.5 * rho;
      if (want_J) {
#pragma unroll
        for (int m = 0; m < 3; m++) {
          const double rm = rs[m] * w;
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
        double* o = blk + (size_t)k * 28;
#pragma unroll
        for (int e = 0; e < 21; e++) o[e] = A[e];
#pragma unroll
        for (int e = 0; e < 6; e++) o[21 + e] = gg[e];
      }
    }
  }
  return cost;
}

// CTA 0: H += the LiDAR blocks the cluster left in blk
__device__ void lidar_add(const Win& W, double* H, double* g, double* hd, const double* blk) {
  for (int t = threadIdx.x; t < W.N * 27; t += blockDim.x) {
    const int k = t / 27, e = t % 27; const double v = blk[(size_t)k * 28 + e];
    if (e < 21) {
      int a = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
      while (a * (a + 1) / 2 > e) a--;
      while ((a + 1) * (a + 2) / 2 <= e) a++;
      const int b = e - a * (a + 1) / 2;
      H[tidx(15 * k + a, 15 * k + b)] += v; if (a == b) hd[15 * k + a] += v;
    } else g[15 * k + e - 21] += v;
  }
}

// ---- L: landmarks rank = r (mod G) ------------------------------------------------------------------------------------------------------
// jac_mode / craw / scl as landmark_reduce (dogleg: Jacobi scale of the landmark columns, fixed at the first linearisation)
__device__ void landmark_reduce_cluster(const SolveParams& P, const Win& W, double* cinv, double* glam, double* scr, double mu, int r, int G,
                                        int jac_mode = 0, double* craw = nullptr, double* scl = nullptr) {
  const int nlm = W.h->n_lm;
  const int32_t* lm_start = W.lm_start(); const int32_t* ix = W.i(OFF_PROJ_IDX);
  const double* part = scr + P.sl.part; double* E = scr + P.sl.E;
  for (int rnk = r + G * threadIdx.x; rnk < nlm; rnk += G * blockDim.x) {
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
    if (C > 0.0) {
      double dd = fmin(fmax(C, 1e-6), 1e32);
      if (jac_mode) {
        const double sj = jac_mode == 1 ? 1.0 / (1.0 + sqrt(C)) : scl[rnk];
        if (jac_mode == 1) scl[rnk] = sj;
        dd = fmin(fmax(C * sj * sj, 1e-6), 1e32) / (sj * sj);
        craw[rnk] = C;
      }
      cinv[rnk] = 1.0 / (C + mu * dd); glam[rnk] = s[1];
    } else { cinv[rnk] = 0.0; glam[rnk] = 0.0; if (jac_mode) craw[rnk] = 0.0; }
  }
}

// partial Hv(lower) = -sum over OWN landmarks cinv e e^T, partial gv; written to the cluster's exchange area (global)
__device__ void schur_syrk_cluster(const SolveParams& P, const Win& W, const double* cinv, const double* glam, double* Hv_out, double* gv_out,
                                   double* chunk, const double* scr, int chunk_cap, int r, int G) {
  const int Dv = W.Dv, Dvp = W.Dvp, nlm = W.h->n_lm;
  const int n_own = nlm > r ? (nlm - r + G - 1) / G : 0;
  const int ECH = max(1, min(chunk_cap / Dvp, max(n_own, 1)));
  const int nt = (Dv + 2) / 3;
  const int ntiles = nt * (nt + 1) / 2;
  const double* E = scr + P.sl.E;
  for (int tbase = 0; tbase < ntiles; tbase += blockDim.x) {
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
    for (int j0 = 0; j0 < n_own; j0 += ECH) {
      const int n = min(ECH, n_own - j0);
      __syncthreads();
      for (int k = threadIdx.x; k < n * Dvp; k += blockDim.x) { const int j = k / Dvp, c = k - j * Dvp; cp_async_f64(chunk + k, E + (size_t)(r + G * (j0 + j)) * Dvp + c); }
      cp_async_wait_all();
      __syncthreads();
      if (ti >= 0) {
        for (int j = 0; j < n; j++) {
          const int rnk = r + G * (j0 + j);
          const double ci = cinv[rnk];
          const double* e = chunk + j * Dvp;
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
            const double gl = glam[rnk] * ci;
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
        for (int b = 0; b < 3; b++) { const int ib = 3 * tj + b; if (ib < Dv && ib <= ia) Hv_out[ia * Dvp + ib] = -acc[a][b]; }
        if (tj == 0) gv_out[ia] = -gacc[a];
      }
    }
  }
}

// ---- G: entries t = r (mod G): sum of the G partial Hv / gv + pair blocks -> H (global tiles), g and hd (global, camera-indexed) ---------
__device__ void gather_cluster(const SolveParams& P, const Win& W, const double* hvpart, const double* gvpart, double* H, double* gsc, double* hdsc,
                               const double* scr, const int* pid, int zblk, int r, int G) {
  const int N = W.N, Dv = W.Dv, Dvp = W.Dvp, nsh = Dv - 6 * N, npair = W.h->n_pair, lane = threadIdx.x & 31;
  const double* __restrict__ pp = scr + P.sl.pairpart;
  const double* __restrict__ hvr = hvpart;
  const size_t hvs = (size_t)Dvp * Dvp;
  // the four entry families of gather_visual, each entry owned by one thread (or warp) of ONE CTA of the cluster; the G partial Schur
  // complements are summed on the way
  for (int e = r + G * threadIdx.x; e < 21 * N + nsh * 6 * N; e += G * blockDim.x) {                       // (1)
    const GatherTask T = gather_task1(e, N);
    double s = gather_walk(pp, pid, N, T.q, T.off_anchor, T.off_obs, zblk);
    for (int c = 0; c < G; c++) s += hvr[(size_t)c * hvs + T.a * Dvp + T.b];
    H[tidx(vis2cam(T.a, N), vis2cam(T.b, N))] = s;
  }
  for (int e = r + G * threadIdx.x; e < 12 * N; e += G * blockDim.x) {                                     // (2)
    const int a = e >> 1, q = a / 6, ao = a - 6 * q, which = e & 1;
    double s = gather_walk(pp, pid, N, q, ao * PAIR_LD + (which ? ao : 19), (6 + ao) * PAIR_LD + (which ? 6 + ao : 19), zblk);
    const int ca = vis2cam(a, N);
    if (which) hdsc[ca] = s;
    else { for (int c = 0; c < G; c++) s += gvpart[(size_t)c * Dvp + a]; gsc[ca] = s; }
  }
  {                                                                                                        // (3)
    const int nent = N * (N - 1) / 2 * 36;
#pragma unroll 2
    for (int e = r + G * threadIdx.x; e < nent; e += G * blockDim.x) {
      const int pi = e / 36, o = e - pi * 36, ao = o / 6, bo = o - ao * 6;
      int pb, qb; gather_pair_of(pi, pb, qb);
      const int pr = pid[qb * N + pb];
      double v = pp[(size_t)(pr < 0 ? zblk : pr) * (PAIR_LD * PAIR_LD) + (6 + ao) * PAIR_LD + bo];
      const int a = 6 * pb + ao, b = 6 * qb + bo;
      for (int c = 0; c < G; c++) v += hvr[(size_t)c * hvs + a * Dvp + b];
      H[tidx(vis2cam(a, N), vis2cam(b, N))] = v;
    }
  }
  for (int wt = r + G * (threadIdx.x >> 5); wt < nsh * (nsh + 1) / 2 + nsh; wt += G * SOLVE_WARPS) {       // (4)
    int ia, ib; const bool grad = wt < nsh;
    if (grad) { ia = wt; ib = wt; }
    else { const int u = wt - nsh; ia = 0; while ((ia + 1) * (ia + 2) / 2 <= u) ia++; ib = u - ia * (ia + 1) / 2; }
    const int la = ia < 6 ? 12 + ia : 18, lb = ib < 6 ? 12 + ib : 18;
    const int a = 6 * N + ia, b = 6 * N + ib;
    double s0 = 0, s1 = 0;
    for (int pr = lane; pr < npair; pr += 32) {
      const double* blk = pp + (size_t)pr * (PAIR_LD * PAIR_LD);
      if (grad) { s0 += blk[la * PAIR_LD + 19]; s1 += blk[la * PAIR_LD + la]; }
      else s0 += blk[la * PAIR_LD + lb];
    }
    if (lane < G) s0 += grad ? gvpart[(size_t)lane * Dvp + a] : hvr[(size_t)lane * hvs + a * Dvp + b];
    s0 = warp_sum(s0); if (grad) s1 = warp_sum(s1);
    if (lane == 0) {
      const int ca = vis2cam(a, N), cbm = vis2cam(b, N);
      if (grad) { gsc[ca] = s0; hdsc[ca] = s1; }
      else H[tidx(ca, cbm)] = s0;
    }
  }
}

// H += A of the marginalization prior (n x n, n = 136 for a 20-keyframe window) on the mapped columns, for a reduced system kept in GLOBAL
// memory: entries dealt to all CTAs, four in flight per thread (on CTA 0 alone: 48 dependent L2 read-modify-writes per thread).  The diagonal
// also goes to the cluster's hd exchange vector.  One (a, b) of the prior maps to one entry of H: no two threads meet.
__device__ void prior_add_H_cluster(const SolveParams& P, const Win& W, double* H, double* hdg, const double* scr, const int r, const int G) {
  const int n = W.h->prior_n;
  if (n == 0) return;
  const int32_t* col = W.i(OFF_PRIOR_COL);
  const double* A = scr + P.sl.priorA;
  const int nn = n * n, step = G * blockDim.x;
  for (int e0 = r * blockDim.x + threadIdx.x; e0 < nn; e0 += 4 * step) {
    double av[4], hv[4]; int hi[4], ca[4]; bool on[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int e = e0 + u * step;
      on[u] = e < nn; hi[u] = 0; ca[u] = -1; av[u] = 0; hv[u] = 0;
      if (on[u]) {
        const int a = e / n, b = e - a * n;
        const int c1 = col[a], c2 = col[b];
        on[u] = c1 >= c2;
        if (on[u]) { hi[u] = tidx(c1, c2); av[u] = A[e]; hv[u] = H[hi[u]]; if (c1 == c2) ca[u] = c1; }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) if (on[u]) { H[hi[u]] = hv[u] + av[u]; if (ca[u] >= 0) hdg[ca[u]] += av[u]; }
  }
}

// Cholesky of a reduced system kept in GLOBAL memory (20-keyframe windows), by the whole cluster: CTA 0 runs what cholesky_tiles_staged runs
// (write-back of the diagonal tile, panel column, next diagonal tile and its factorisation, tile inverses, b row), the trailing tiles of every
// tile row are dealt round-robin to ALL CTAs, each with the panel column staged in its own shared memory.  Two cluster barriers per tile row
// (panel column visible -> trailing tiles visible).  On CTA 0 alone the trailing update was 60 % of a config-4 solve on a cluster.
// Every CTA passes its own stage ((nb + 1) tiles); b, linv, dinv, flag are CTA 0's.
__device__ void cholesky_tiles_cluster(double* H, double* b, double* linv, double* dinv, int nb, int* flag, double* stage, const int r, const int G) {
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int dw = SOLVE_WARPS - 1;
  double* dcur = stage; double* dnext = stage + TSZ; double* psm = stage + 2 * TSZ;   // psm tile j = L(kb + 1 + j, kb)
  if (r == 0) {
    for (int e = t; e < 16 * TLD; e += blockDim.x) dcur[e] = H[e];
    __syncthreads();
    if (warp == dw) chol_diag_factor<true>(dcur, dinv, flag, lane);
    __syncthreads();
  }
  for (int kb = 0; kb < nb; kb++) {
    const int rem = nb - kb - 1;
    if (r == 0) {
      double* Akk_g = H + (size_t)(tri(kb) + kb) * TSZ;
      for (int e = t; e < 16 * TLD; e += blockDim.x) Akk_g[e] = dcur[e];
      const int nrows = rem * 16 + 1;
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
    }
    if (rem == 0) break;
    cluster_sync_all();                              // the panel column L(kb+1.., kb) is in global memory
    if (r != 0) {
      for (int e = t; e < rem * 16 * TLD; e += blockDim.x) {
        const int j = e / (16 * TLD), o = e - j * 16 * TLD;
        cp_async_f64(psm + (size_t)j * TSZ + o, H + (size_t)(tri(kb + 1 + j) + kb) * TSZ + o);
      }
      cp_async_wait_all();
      __syncthreads();
    } else {
      // the next diagonal tile, updated, into its shared-memory buffer
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
    }
    if (r == 0 && warp == dw) {
      chol_diag_factor<true>(dnext, dinv + (kb + 1) * 16, flag, lane);
    } else if (r == 0 && warp == (dw & 3)) {
      chol_diag_inverse(dcur, dinv + kb * 16, linv + kb * 256);
    } else if (r != 0 || (warp & 3) != (dw & 3)) {
      // trailing tiles tl = 1 .. tri(rem) - 1 (tile 0 is the next diagonal tile): tile tl belongs to CTA tl % G; CTA 0 also takes the b row
      const int ntile = tri(rem) - 1;
      const int first = r == 0 ? G : r, mine = first <= ntile ? (ntile - first) / G + 1 : 0;
      const int nb_tasks = r == 0 ? rem * 16 : 0, ntl = mine * 16 + nb_tasks;
      const int tu = r == 0 ? (warp - (warp >> 2)) * 32 + lane : t, stride = r == 0 ? (SOLVE_WARPS - SOLVE_WARPS / 4) * 32 : SOLVE_THREADS;
      for (int it = tu; it < ntl; it += stride) {
        if (it < mine * 16) {
          const int tl = first + G * (it >> 4), sub = it & 15;
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
            for (int c = 0; c < 4; c++) { acc[a][c] = 0; old[a][c] = Aij[(r0 + a) * TLD + c0 + c]; }
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
          const int q = it - mine * 16, j = q >> 4, c = q & 15;
          const double* Ljk = psm + (size_t)j * TSZ;
          const double* yk = b + kb * 16;
          double s0 = 0, s1 = 0;
#pragma unroll
          for (int m = 0; m < 16; m += 2) { s0 = fma(yk[m], Ljk[c * TLD + m], s0); s1 = fma(yk[m + 1], Ljk[c * TLD + m + 1], s1); }
          b[(kb + 1 + j) * 16 + c] -= s0 + s1;
        }
      }
    }
    cluster_sync_all();                              // every trailing tile of this row is in global memory
    if (r == 0) { double* sw = dcur; dcur = dnext; dnext = sw; }
  }
  if (r == 0) {
    if (warp == 0) chol_diag_inverse(dcur, dinv + (nb - 1) * 16, linv + (nb - 1) * 256);
    __syncthreads();
  }
}

}  // namespace vb
