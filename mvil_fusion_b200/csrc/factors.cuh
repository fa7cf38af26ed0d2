// factors.cuh — residual + tangent-space Jacobian of every factor family of the sliding window, FP64.
// Each function cites the reference code whose arithmetic it reproduces (paths relative to the reference tree).
// Jacobian blocks are TANGENT-space (pose blocks 6 wide): the reference's 7-wide global blocks have a zero 7th
// column and PoseLocalParameterization::ComputeJacobian is [I6;0] (factor/pose_local_parameterization.cpp:20-27).
#pragma once
#include "vm.cuh"

namespace vf {
using namespace vm;

enum { O_P = 0, O_R = 3, O_V = 6, O_BA = 9, O_BG = 12 };  // parameters.h:80-87

struct BaCfg {
  double s_info;       // FOCAL_LENGTH / 2  (estimator.cpp:18)
  double G[3];         // gravity, subtracted as a positive vector (parameters.cpp:32,103)
  double tr_over_row;  // TR / ROW
  double half_row;     // ROW / 2
  double cauchy_a;     // CauchyLoss(a) on projection / ICP / LPS
  double huber_a;      // HuberLoss(a) on LiDAR edge / plane
  int use_td;          // ESTIMATE_TD
  int est_ex;          // ESTIMATE_EXTRINSIC
};

// ---- robust losses [ceres loss_function.cc]; with rho'' <= 0 the corrector of ResidualBlockInfo::Evaluate
// (factor/marginalization_factor.cpp:37-68, branch :49-53) is r <- sqrt(rho') r, J <- sqrt(rho') J. ----
VHD void cauchy(double a, double s, double& rho, double& w) { const double b = a * a, sum = 1.0 + s / b; rho = b * log(sum); w = sqrt(fmax(2.2250738585072014e-308, 1.0 / sum)); }
// weight only (no log): used when the robust cost value itself is not needed
VHD double cauchy_w(double a, double s) { return sqrt(fmax(2.2250738585072014e-308, 1.0 / (1.0 + s / (a * a)))); }
VHD void huber(double a, double s, double& rho, double& w) {
  const double b = a * a;
  if (s > b) { const double r = sqrt(s); rho = 2 * a * r - b; w = sqrt(fmax(2.2250738585072014e-308, a / r)); }
  else { rho = s; w = 1.0; } }

// ---- ProjectionTdFactor::Evaluate (factor/projection_td_factor.cpp:34-141); with use_td == 0 it is
// ProjectionFactor::Evaluate (factor/projection_factor.cpp:21-121).
// c[14] = pts_i(3) pts_j(3) vel_i(2) vel_j(2) td_i td_j row_i row_j   (row = uv.y, ROW/2 subtracted here, :18-19)
// J (may be null): 2 x 20 row-major [pose_i 6 | pose_j 6 | ex 6 | lambda | td]
// loss_a > 0: the Cauchy(loss_a) corrector (r <- w r, J <- w J, w = sqrt(rho')) is folded into the weight s_info, so the
// caller gets corrected rows without a second pass over J.
VHD void proj_eval(const BaCfg& cfg, const double* c, const double* pose_i, const double* pose_j, const double* ex,
                   double lam, double td, double* r, double* J, double loss_a = 0.0) {
  const v3 Pi = ld3(pose_i), Pj = ld3(pose_j), tic = ld3(ex);
  const m3 Ri = q2R(ldq(pose_i + 3)), Rj = q2R(ldq(pose_j + 3)), ric = q2R(ldq(ex + 3));
  v3 pi = mk(c[0], c[1], c[2]), pj = mk(c[3], c[4], c[5]);
  const v3 vi = mk(c[6], c[7], 0.0), vj = mk(c[8], c[9], 0.0);
  if (cfg.use_td) {
    const double si = td - c[10] + cfg.tr_over_row * (c[12] - cfg.half_row);
    const double sj = td - c[11] + cfg.tr_over_row * (c[13] - cfg.half_row);
    pi = pi - vi * si; pj = pj - vj * sj;                       // :51-52
  }
  const double inv_lam = 1.0 / lam;
  const v3 pc_i = pi * inv_lam;                                   // :53
  const v3 pb_i = mul(ric, pc_i) + tic;
  const v3 pw = mul(Ri, pb_i) + Pi;
  const v3 pb_j = mulT(Rj, pw - Pj);
  const v3 pc_j = mulT(ric, pb_j - tic);
  const double iz = 1.0 / pc_j.z;
  double si = cfg.s_info;
  r[0] = si * (pc_j.x * iz - pj.x);                               // :63-67
  r[1] = si * (pc_j.y * iz - pj.y);
  if (loss_a > 0.0) { const double w = cauchy_w(loss_a, r[0] * r[0] + r[1] * r[1]); r[0] *= w; r[1] *= w; si *= w; }
  if (!J) return;
  const double red[2][3] = {{si * iz, 0.0, -si * pc_j.x * iz * iz}, {0.0, si * iz, -si * pc_j.y * iz * iz}};  // :87-90
  const m3 A = mulT(ric, transpose(Rj));                          // ric^T Rj^T
  const m3 ARi = mul(A, Ri);
  const m3 Bi = scale(mul(ARi, skew(pb_i)), -1.0);                // :97-100
  const m3 Bj = mulT(ric, skew(pb_j));                            // :109-112
  const m3 T = mul(ARi, ric);                                     // tmp_r :119
  const m3 Aex = mulT(ric, sub(mulT(Rj, Ri), eye()));             // :118
  const v3 tex = mulT(ric, mulT(Rj, mul(Ri, tic) + Pi - Pj) - tic);
  const m3 Bex = add(add(scale(mul(T, skew(pc_i)), -1.0), skew(mul(T, pc_i))), skew(tex));  // :120-121
  const v3 jl = mul(T, pi) * (-inv_lam * inv_lam);                // :129
  const v3 jt = mul(T, vi) * (-inv_lam);                          // :135
#pragma unroll
  for (int a = 0; a < 2; a++) {
    double* Jr = J + 20 * a;
#pragma unroll
    for (int b = 0; b < 3; b++) {
      Jr[b] = red[a][0] * A.m[0][b] + red[a][1] * A.m[1][b] + red[a][2] * A.m[2][b];
      Jr[3 + b] = red[a][0] * Bi.m[0][b] + red[a][1] * Bi.m[1][b] + red[a][2] * Bi.m[2][b];
      Jr[6 + b] = -Jr[b];                                         // :108
      Jr[9 + b] = red[a][0] * Bj.m[0][b] + red[a][1] * Bj.m[1][b] + red[a][2] * Bj.m[2][b];
      Jr[12 + b] = red[a][0] * Aex.m[0][b] + red[a][1] * Aex.m[1][b] + red[a][2] * Aex.m[2][b];
      Jr[15 + b] = red[a][0] * Bex.m[0][b] + red[a][1] * Bex.m[1][b] + red[a][2] * Bex.m[2][b];
    }
    Jr[18] = red[a][0] * jl.x + red[a][1] * jl.y + red[a][2] * jl.z;
    Jr[19] = cfg.use_td ? (red[a][0] * jt.x + red[a][1] * jt.y + red[a][2] * jt.z + si * (a == 0 ? vj.x : vj.y)) : 0.0;  // :135
  }
}

// ---- the same factor as ROW-VECTOR chains: instead of forming the 3x3 products A = ric^T Rj^T, A Ri, T = A Ri ric, Bex ...
// of projection_td_factor.cpp:94-121 and reducing each with the 2x3 matrix `reduce`, every Jacobian row a is propagated as a
// 3-vector through the rotations:  u = red_a ric^T,  v = u Rj^T (= red_a A),  w = v Ri (= red_a A Ri),  y = w ric (= red_a T).
//   d/dP_i = v, d/dtheta_i = -(w x pb_i), d/dP_j = -v, d/dtheta_j = u x pb_j, d/dt_ic = w - u  (ric^T (Rj^T Ri - I)),
//   d/dtheta_ic = -(y x pc_i) + red_a x pc_j   (skew(T pc_i) + skew(tex) = skew(pc_j)),  d/dlambda = -(y . p_i)/lambda^2,
//   d/dtd = -(y . v_i)/lambda + s v_j[a].
// Identical mathematics to proj_eval (same formulas re-associated), ~40 % fewer FP64 operations and three live matrices
// instead of ten.  Ri, Rj, ric are passed in so that callers can keep a per-keyframe rotation table.
VHD void proj_eval_rows(const BaCfg& cfg, const double* c, const m3& Ri, const m3& Rj, const m3& ric, const v3& Pi, const v3& Pj, const v3& tic,
                        double lam, double td, double* r, double* J, double loss_a = 0.0) {
  v3 pi = mk(c[0], c[1], c[2]), pj = mk(c[3], c[4], c[5]);
  const v3 vi = mk(c[6], c[7], 0.0), vj = mk(c[8], c[9], 0.0);
  if (cfg.use_td) {
    const double si = td - c[10] + cfg.tr_over_row * (c[12] - cfg.half_row);
    const double sj = td - c[11] + cfg.tr_over_row * (c[13] - cfg.half_row);
    pi = pi - vi * si; pj = pj - vj * sj;
  }
  const double inv_lam = 1.0 / lam;
  const v3 pc_i = pi * inv_lam;
  const v3 pb_i = mul(ric, pc_i) + tic;
  const v3 pw = mul(Ri, pb_i) + Pi;
  const v3 pb_j = mulT(Rj, pw - Pj);
  const v3 pc_j = mulT(ric, pb_j - tic);
  const double iz = 1.0 / pc_j.z;
  double s = cfg.s_info;
  r[0] = s * (pc_j.x * iz - pj.x);
  r[1] = s * (pc_j.y * iz - pj.y);
  if (loss_a > 0.0) { const double w = cauchy_w(loss_a, r[0] * r[0] + r[1] * r[1]); r[0] *= w; r[1] *= w; s *= w; }
  if (!J) return;
  const double siz = s * iz;
#pragma unroll
  for (int a = 0; a < 2; a++) {
    double* Jr = J + 20 * a;
    const v3 ra = a == 0 ? mk(siz, 0.0, -siz * pc_j.x * iz) : mk(0.0, siz, -siz * pc_j.y * iz);
    const double rxy = a == 0 ? ra.x : ra.y;   // the non-zero one of (ra.x, ra.y)
    const v3 u = a == 0 ? mk(rxy * ric.m[0][0] + ra.z * ric.m[0][2], rxy * ric.m[1][0] + ra.z * ric.m[1][2], rxy * ric.m[2][0] + ra.z * ric.m[2][2])
                        : mk(rxy * ric.m[0][1] + ra.z * ric.m[0][2], rxy * ric.m[1][1] + ra.z * ric.m[1][2], rxy * ric.m[2][1] + ra.z * ric.m[2][2]);
    const v3 v = mul(Rj, u);        // (u Rj^T)_k = sum_m u_m Rj[k][m]
    const v3 w = mulT(Ri, v);       // (v Ri)_k   = sum_m v_m Ri[m][k]
    const v3 y = mulT(ric, w);      // (w ric)_k
    const v3 ji = cross(pb_i, w);   // -(w x pb_i)
    const v3 jj = cross(u, pb_j);
    const v3 je = cross(pc_i, y) + cross(ra, pc_j);
    Jr[0] = v.x; Jr[1] = v.y; Jr[2] = v.z; Jr[3] = ji.x; Jr[4] = ji.y; Jr[5] = ji.z;
    Jr[6] = -v.x; Jr[7] = -v.y; Jr[8] = -v.z; Jr[9] = jj.x; Jr[10] = jj.y; Jr[11] = jj.z;
    Jr[12] = w.x - u.x; Jr[13] = w.y - u.y; Jr[14] = w.z - u.z; Jr[15] = je.x; Jr[16] = je.y; Jr[17] = je.z;
    Jr[18] = -dot(y, pi) * inv_lam * inv_lam;
    Jr[19] = cfg.use_td ? (-dot(y, vi) * inv_lam + s * (a == 0 ? vj.x : vj.y)) : 0.0;
  }
}

// ---- the same factor with everything that only depends on the keyframe PAIR (i, j) and the extrinsic hoisted out.
// In a window ~900 projection factors share ~45 (i, j) pairs, so A = ric^T Rj^T, A Ri, T = A Ri ric, ric^T (Rj^T Ri - I)
// and the translation term of the extrinsic Jacobian are built once per pair per linearisation (proj_pair_ctx) and the
// per-factor work drops to the point chain plus a few 2x3 * 3x3 products (proj_eval_ctx).  Same formulas as proj_eval.
constexpr int PCTX_LD = 84;   // doubles per pair context
// ctx: A(0) ARi(9) T(18) Aex(27) ric(36) Ri(45) Rj(54) Pi(63) Pj(66) tic(69) tex(72) [75..83 pad]
VHD void proj_pair_ctx(const double* pose_i, const double* pose_j, const double* ex, double* ctx) {
  const v3 Pi = ld3(pose_i), Pj = ld3(pose_j), tic = ld3(ex);
  const m3 Ri = q2R(ldq(pose_i + 3)), Rj = q2R(ldq(pose_j + 3)), ric = q2R(ldq(ex + 3));
  const m3 A = mulT(ric, transpose(Rj)), ARi = mul(A, Ri), T = mul(ARi, ric), Aex = mulT(ric, sub(mulT(Rj, Ri), eye()));
  const v3 tex = mulT(ric, mulT(Rj, mul(Ri, tic) + Pi - Pj) - tic);
  const m3* ms[7] = {&A, &ARi, &T, &Aex, &ric, &Ri, &Rj};
  for (int k = 0; k < 7; k++) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) ctx[9 * k + 3 * i + j] = ms[k]->m[i][j];
  ctx[63] = Pi.x; ctx[64] = Pi.y; ctx[65] = Pi.z; ctx[66] = Pj.x; ctx[67] = Pj.y; ctx[68] = Pj.z;
  ctx[69] = tic.x; ctx[70] = tic.y; ctx[71] = tic.z; ctx[72] = tex.x; ctx[73] = tex.y; ctx[74] = tex.z;
}
VHD m3 ldm(const double* p) { m3 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.m[i][j] = p[3 * i + j];
  return r; }
// J: 2 x 20 row-major [pose_i 6 | pose_j 6 | ex 6 | lambda | td] as proj_eval
VHD void proj_eval_ctx(const BaCfg& cfg, const double* ctx, const double* c, double lam, double td, double* r, double* J) {
  const m3 A = ldm(ctx), ARi = ldm(ctx + 9), T = ldm(ctx + 18), Aex = ldm(ctx + 27), ric = ldm(ctx + 36), Ri = ldm(ctx + 45), Rj = ldm(ctx + 54);
  const v3 Pi = ld3(ctx + 63), Pj = ld3(ctx + 66), tic = ld3(ctx + 69), tex = ld3(ctx + 72);
  v3 pi = mk(c[0], c[1], c[2]), pj = mk(c[3], c[4], c[5]);
  const v3 vi = mk(c[6], c[7], 0.0), vj = mk(c[8], c[9], 0.0);
  if (cfg.use_td) {
    const double si = td - c[10] + cfg.tr_over_row * (c[12] - cfg.half_row);
    const double sj = td - c[11] + cfg.tr_over_row * (c[13] - cfg.half_row);
    pi = pi - vi * si; pj = pj - vj * sj;
  }
  const double inv_lam = 1.0 / lam;
  const v3 pc_i = pi * inv_lam;
  const v3 pb_i = mul(ric, pc_i) + tic;
  const v3 pw = mul(Ri, pb_i) + Pi;
  const v3 pb_j = mulT(Rj, pw - Pj);
  const v3 pc_j = mulT(ric, pb_j - tic);
  const double iz = 1.0 / pc_j.z;
  r[0] = cfg.s_info * (pc_j.x * iz - pj.x);
  r[1] = cfg.s_info * (pc_j.y * iz - pj.y);
  if (!J) return;
  const double red[2][3] = {{cfg.s_info * iz, 0.0, -cfg.s_info * pc_j.x * iz * iz}, {0.0, cfg.s_info * iz, -cfg.s_info * pc_j.y * iz * iz}};
  const v3 Tpc = mul(T, pc_i);
  const v3 sk = Tpc + tex;                          // skew(T pc_i) + skew(tex) = skew(T pc_i + tex)
#pragma unroll
  for (int a = 0; a < 2; a++) {
    double* Jr = J + 20 * a;
    // row vectors red_a * M for the shared matrices
    const v3 rA = mk(red[a][0] * A.m[0][0] + red[a][1] * A.m[1][0] + red[a][2] * A.m[2][0], red[a][0] * A.m[0][1] + red[a][1] * A.m[1][1] + red[a][2] * A.m[2][1], red[a][0] * A.m[0][2] + red[a][1] * A.m[1][2] + red[a][2] * A.m[2][2]);
    const v3 rARi = mk(red[a][0] * ARi.m[0][0] + red[a][1] * ARi.m[1][0] + red[a][2] * ARi.m[2][0], red[a][0] * ARi.m[0][1] + red[a][1] * ARi.m[1][1] + red[a][2] * ARi.m[2][1], red[a][0] * ARi.m[0][2] + red[a][1] * ARi.m[1][2] + red[a][2] * ARi.m[2][2]);
    const v3 rT = mk(red[a][0] * T.m[0][0] + red[a][1] * T.m[1][0] + red[a][2] * T.m[2][0], red[a][0] * T.m[0][1] + red[a][1] * T.m[1][1] + red[a][2] * T.m[2][1], red[a][0] * T.m[0][2] + red[a][1] * T.m[1][2] + red[a][2] * T.m[2][2]);
    const v3 rAex = mk(red[a][0] * Aex.m[0][0] + red[a][1] * Aex.m[1][0] + red[a][2] * Aex.m[2][0], red[a][0] * Aex.m[0][1] + red[a][1] * Aex.m[1][1] + red[a][2] * Aex.m[2][1], red[a][0] * Aex.m[0][2] + red[a][1] * Aex.m[1][2] + red[a][2] * Aex.m[2][2]);
    const v3 rricT = mk(red[a][0] * ric.m[0][0] + red[a][1] * ric.m[0][1] + red[a][2] * ric.m[0][2], red[a][0] * ric.m[1][0] + red[a][1] * ric.m[1][1] + red[a][2] * ric.m[1][2], red[a][0] * ric.m[2][0] + red[a][1] * ric.m[2][1] + red[a][2] * ric.m[2][2]);   // red_a ric^T
    const v3 ra_ = mk(red[a][0], red[a][1], red[a][2]);
    // u^T skew(b) = (u x b)^T
    const v3 ji_rot = cross(rARi, pb_i) * -1.0;      // -red ARi skew(pb_i)
    const v3 jj_rot = cross(rricT, pb_j);            //  red ric^T skew(pb_j)
    const v3 jex_rot = cross(rT, pc_i) * -1.0 + cross(ra_, sk);
    Jr[0] = rA.x; Jr[1] = rA.y; Jr[2] = rA.z; Jr[3] = ji_rot.x; Jr[4] = ji_rot.y; Jr[5] = ji_rot.z;
    Jr[6] = -rA.x; Jr[7] = -rA.y; Jr[8] = -rA.z; Jr[9] = jj_rot.x; Jr[10] = jj_rot.y; Jr[11] = jj_rot.z;
    Jr[12] = rAex.x; Jr[13] = rAex.y; Jr[14] = rAex.z; Jr[15] = jex_rot.x; Jr[16] = jex_rot.y; Jr[17] = jex_rot.z;
    Jr[18] = -dot(rT, pi) * inv_lam * inv_lam;
    Jr[19] = cfg.use_td ? (-dot(rT, vi) * inv_lam + cfg.s_info * (a == 0 ? vj.x : vj.y)) : 0.0;
  }
}

// ---- IMUFactor (factor/imu_factor.h:19-181) ---------------------------------------------------------------------
// sqrt_info = LLT(covariance.inverse()).matrixL().transpose() (:64). Same algorithm family as Eigen: partial-pivot LU
// inverse, then Cholesky. cov: 15x15 column-major (symmetric); W: 15x15 row-major, upper triangular.
VHD bool imu_sqrt_info(const double* cov, double* W) {
  double a[225], inv[225]; int piv[15];
  for (int i = 0; i < 225; i++) a[i] = cov[i];   // symmetric: row-major == column-major up to rounding of the propagation
  for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) a[i * 15 + j] = cov[j * 15 + i];
  for (int i = 0; i < 15; i++) piv[i] = i;
  for (int k = 0; k < 15; k++) {
    int p = k; double best = fabs(a[k * 15 + k]);
    for (int i = k + 1; i < 15; i++) if (fabs(a[i * 15 + k]) > best) { best = fabs(a[i * 15 + k]); p = i; }
    if (best == 0.0) return false;
    if (p != k) { for (int j = 0; j < 15; j++) { double t = a[k * 15 + j]; a[k * 15 + j] = a[p * 15 + j]; a[p * 15 + j] = t; } int t = piv[k]; piv[k] = piv[p]; piv[p] = t; }
    for (int i = k + 1; i < 15; i++) { a[i * 15 + k] /= a[k * 15 + k]; const double l = a[i * 15 + k]; for (int j = k + 1; j < 15; j++) a[i * 15 + j] -= l * a[k * 15 + j]; }
  }
  for (int c = 0; c < 15; c++) {
    double x[15];
    for (int i = 0; i < 15; i++) x[i] = (piv[i] == c) ? 1.0 : 0.0;
    for (int i = 0; i < 15; i++) { double s = x[i]; for (int k = 0; k < i; k++) s -= a[i * 15 + k] * x[k]; x[i] = s; }
    for (int i = 14; i >= 0; i--) { double s = x[i]; for (int k = i + 1; k < 15; k++) s -= a[i * 15 + k] * x[k]; x[i] = s / a[i * 15 + i]; }
    for (int i = 0; i < 15; i++) inv[i * 15 + c] = x[i];
  }
  // lower Cholesky of inv (reads the lower triangle like Eigen's LLT), W = L^T
  for (int j = 0; j < 15; j++) {
    double d = inv[j * 15 + j];
    for (int k = 0; k < j; k++) d -= inv[j * 15 + k] * inv[j * 15 + k];
    if (!(d > 0.0)) return false;
    d = sqrt(d); inv[j * 15 + j] = d;
    for (int i = j + 1; i < 15; i++) { double s = inv[i * 15 + j]; for (int k = 0; k < j; k++) s -= inv[i * 15 + k] * inv[j * 15 + k]; inv[i * 15 + j] = s / d; }
  }
  for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) W[i * 15 + j] = (j >= i) ? inv[j * 15 + i] : 0.0;
  return true;
}

VHD m3 blk_cm15(const double* cm, int r, int c) { m3 m;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) m.m[i][j] = cm[(c + j) * 15 + r + i];
  return m; }
VHD void put33(double* J, int ld, int r, int c, const m3& m, double s = 1.0) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) J[(r + i) * ld + c + j] = s * m.m[i][j]; }
// bottom-right 3x3 of Utility::Qleft / Qright (utility.h:46-64)
VHD m3 qleft33(const q4& q) { m3 r = skew(mk(q.x, q.y, q.z)); r.m[0][0] += q.w; r.m[1][1] += q.w; r.m[2][2] += q.w; return r; }
VHD m3 qright33(const q4& q) { m3 r = scale(skew(mk(q.x, q.y, q.z)), -1.0); r.m[0][0] += q.w; r.m[1][1] += q.w; r.m[2][2] += q.w; return r; }

// Unweighted residual (15) and Jacobian (15 x 30 row-major over [pose_i 6 | sb_i 9 | pose_j 6 | sb_j 9]; J must be
// zero-filled by the caller).  IntegrationBase::evaluate (integration_base.h:175-201) + imu_factor.h:88-177.
// pre = vils_preint as 467 doubles: dp(3) dq(4 xyzw) dv(3) lin_ba(3) lin_bg(3) sum_dt jac(225 cm) cov(225 cm).
VHD void imu_eval_raw(const double* pre, const double* G3, const double* pose_i, const double* sb_i, const double* pose_j,
                      const double* sb_j, double* r, double* J) {
  const v3 Pi = ld3(pose_i), Pj = ld3(pose_j), Vi = ld3(sb_i), Vj = ld3(sb_j), Bai = ld3(sb_i + 3), Bgi = ld3(sb_i + 6), Baj = ld3(sb_j + 3), Bgj = ld3(sb_j + 6);
  const q4 Qi = ldq(pose_i + 3), Qj = ldq(pose_j + 3);
  const v3 dp = ld3(pre), dv = ld3(pre + 7), G = ld3(G3);
  const q4 dq = ldq(pre + 3);
  const double dt = pre[16];
  const double* jac = pre + 17;
  const m3 dp_dba = blk_cm15(jac, O_P, O_BA), dp_dbg = blk_cm15(jac, O_P, O_BG), dq_dbg = blk_cm15(jac, O_R, O_BG);
  const m3 dv_dba = blk_cm15(jac, O_V, O_BA), dv_dbg = blk_cm15(jac, O_V, O_BG);
  const v3 dba = Bai - ld3(pre + 10), dbg = Bgi - ld3(pre + 13);
  const v3 th = mul(dq_dbg, dbg);
  const q4 cdq = qmul(dq, mkq(1.0, th.x * 0.5, th.y * 0.5, th.z * 0.5));       // :188 deltaQ, not normalised
  const v3 cdv = dv + mul(dv_dba, dba) + mul(dv_dbg, dbg);
  const v3 cdp = dp + mul(dp_dba, dba) + mul(dp_dbg, dbg);
  const q4 Qi_inv = qinv(Qi);
  const v3 ap = qrot(Qi_inv, G * (0.5 * dt * dt) + Pj - Pi - Vi * dt);
  const v3 av = qrot(Qi_inv, G * dt + Vj - Vi);
  const v3 rp = ap - cdp, rv = av - cdv;
  const q4 qe = qmul(qinv(cdq), qmul(Qi_inv, Qj));
  r[0] = rp.x; r[1] = rp.y; r[2] = rp.z; r[3] = 2 * qe.x; r[4] = 2 * qe.y; r[5] = 2 * qe.z; r[6] = rv.x; r[7] = rv.y; r[8] = rv.z;
  r[9] = Baj.x - Bai.x; r[10] = Baj.y - Bai.y; r[11] = Baj.z - Bai.z; r[12] = Bgj.x - Bgi.x; r[13] = Bgj.y - Bgi.y; r[14] = Bgj.z - Bgi.z;
  if (!J) return;
  const m3 RiT = q2R(Qi_inv);
  const q4 qji = qmul(qinv(Qj), Qi);
  // pose_i (:88-113)
  put33(J, 30, O_P, 0, RiT, -1.0);
  put33(J, 30, O_P, 3, skew(ap));
  {  // -(Qleft(Qj^-1 Qi) Qright(cdq)).bottomRight3x3 — 4x4 product restricted to rows/cols 1..3
    const m3 L = qleft33(qji), R = qright33(cdq);
    const v3 lv = mk(qji.x, qji.y, qji.z), rv_ = mk(cdq.x, cdq.y, cdq.z);
    m3 LR = mul(L, R);   // + column 0 of L (= q.vec) times row 0 of R (= -q.vec^T)
    const double lcol[3] = {lv.x, lv.y, lv.z}, rrow[3] = {-rv_.x, -rv_.y, -rv_.z};
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) LR.m[i][j] += lcol[i] * rrow[j];
    put33(J, 30, O_R, 3, LR, -1.0);
  }
  put33(J, 30, O_V, 3, skew(av));
  // speed-bias_i (:114-142), columns 6..14
  put33(J, 30, O_P, 6, RiT, -dt);
  put33(J, 30, O_P, 9, dp_dba, -1.0);
  put33(J, 30, O_P, 12, dp_dbg, -1.0);
  put33(J, 30, O_R, 12, mul(qleft33(qmul(qji, dq)), dq_dbg), -1.0);             // :127 un-corrected delta_q
  put33(J, 30, O_V, 6, RiT, -1.0);
  put33(J, 30, O_V, 9, dv_dba, -1.0);
  put33(J, 30, O_V, 12, dv_dbg, -1.0);
  put33(J, 30, O_BA, 9, eye(), -1.0);
  put33(J, 30, O_BG, 12, eye(), -1.0);
  // pose_j (:143-161), columns 15..20
  put33(J, 30, O_P, 15, RiT);
  put33(J, 30, O_R, 18, qleft33(qe));
  // speed-bias_j (:162-177), columns 21..29
  put33(J, 30, O_V, 21, RiT);
  put33(J, 30, O_BA, 24, eye());
  put33(J, 30, O_BG, 27, eye());
}

// ---- warp-cooperative form of imu_eval_raw.  The quaternion algebra (imu_core, ~400 flops) is done by one lane; the 450-entry
// Jacobian, which is 18 scaled copies of 3x3 blocks, is then ASSEMBLED by all lanes from a flat source array through a
// constant table:  J[e] = scale(code_e) * core[off_e],  scale in {+1, -1, -sum_dt}.
// core layout (doubles): RiT 0..8 | ap 9..11 | av 12..14 | LR 15..23 | M7 16+8.. = 24..32 | Lqe 33..41 | dt 42 | r 43..57 |
//                        one 58 | zero 59 | dp_dba 60..68 | dp_dbg 69..77 | dv_dba 78..86 | dv_dbg 87..95   (row-major 3x3 each)
constexpr int IMU_CORE_LD = 96;
enum { IC_RIT = 0, IC_AP = 9, IC_AV = 12, IC_LR = 15, IC_M7 = 24, IC_LQE = 33, IC_DT = 42, IC_R = 43, IC_ONE = 58, IC_ZERO = 59, IC_JAC = 60 };
// Fills core[0..59].  Same statements as imu_eval_raw up to the put33 calls.
VHD void imu_core(const double* pre, const double* G3, const double* pose_i, const double* sb_i, const double* pose_j, const double* sb_j, double* core) {
  const v3 Pi = ld3(pose_i), Pj = ld3(pose_j), Vi = ld3(sb_i), Vj = ld3(sb_j), Bai = ld3(sb_i + 3), Bgi = ld3(sb_i + 6), Baj = ld3(sb_j + 3), Bgj = ld3(sb_j + 6);
  const q4 Qi = ldq(pose_i + 3), Qj = ldq(pose_j + 3);
  const v3 dp = ld3(pre), dv = ld3(pre + 7), G = ld3(G3);
  const q4 dq = ldq(pre + 3);
  const double dt = pre[16];
  const double* jac = pre + 17;
  const m3 dp_dba = blk_cm15(jac, O_P, O_BA), dp_dbg = blk_cm15(jac, O_P, O_BG), dq_dbg = blk_cm15(jac, O_R, O_BG);
  const m3 dv_dba = blk_cm15(jac, O_V, O_BA), dv_dbg = blk_cm15(jac, O_V, O_BG);
  const v3 dba = Bai - ld3(pre + 10), dbg = Bgi - ld3(pre + 13);
  const v3 th = mul(dq_dbg, dbg);
  const q4 cdq = qmul(dq, mkq(1.0, th.x * 0.5, th.y * 0.5, th.z * 0.5));
  const v3 cdv = dv + mul(dv_dba, dba) + mul(dv_dbg, dbg);
  const v3 cdp = dp + mul(dp_dba, dba) + mul(dp_dbg, dbg);
  const q4 Qi_inv = qinv(Qi);
  const v3 ap = qrot(Qi_inv, G * (0.5 * dt * dt) + Pj - Pi - Vi * dt);
  const v3 av = qrot(Qi_inv, G * dt + Vj - Vi);
  const v3 rp = ap - cdp, rv = av - cdv;
  const q4 qe = qmul(qinv(cdq), qmul(Qi_inv, Qj));
  double* r = core + IC_R;
  r[0] = rp.x; r[1] = rp.y; r[2] = rp.z; r[3] = 2 * qe.x; r[4] = 2 * qe.y; r[5] = 2 * qe.z; r[6] = rv.x; r[7] = rv.y; r[8] = rv.z;
  r[9] = Baj.x - Bai.x; r[10] = Baj.y - Bai.y; r[11] = Baj.z - Bai.z; r[12] = Bgj.x - Bgi.x; r[13] = Bgj.y - Bgi.y; r[14] = Bgj.z - Bgi.z;
  const m3 RiT = q2R(Qi_inv);
  const q4 qji = qmul(qinv(Qj), Qi);
  m3 LR = mul(qleft33(qji), qright33(cdq));
  {
    const double lcol[3] = {qji.x, qji.y, qji.z}, rrow[3] = {-cdq.x, -cdq.y, -cdq.z};
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) LR.m[i][j] += lcol[i] * rrow[j];
  }
  const m3 M7 = mul(qleft33(qmul(qji, dq)), dq_dbg);
  const m3 Lqe = qleft33(qe);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) { core[IC_RIT + 3 * i + j] = RiT.m[i][j]; core[IC_LR + 3 * i + j] = LR.m[i][j]; core[IC_M7 + 3 * i + j] = M7.m[i][j]; core[IC_LQE + 3 * i + j] = Lqe.m[i][j]; }
  core[IC_AP] = ap.x; core[IC_AP + 1] = ap.y; core[IC_AP + 2] = ap.z; core[IC_AV] = av.x; core[IC_AV + 1] = av.y; core[IC_AV + 2] = av.z;
  core[IC_DT] = dt; core[IC_ONE] = 1.0; core[IC_ZERO] = 0.0;
}
// core[60 + e], e < 36: the four bias Jacobian blocks of the pre-integration, row-major 3x3 each
VHD double imu_core_jac(const double* pre, int e) {
  const int b = e / 9, i = (e % 9) / 3, j = e % 3;
  const int r = (b < 2) ? O_P : O_V, c = (b & 1) ? O_BG : O_BA;
  return pre[17 + (c + j) * 15 + r + i];
}
// Assembly table: entry e = a * 30 + c of the 15 x 30 row-major Jacobian -> (off << 2) | code, code 0: +1, 1: -1, 2: -dt.
inline void imu_build_table(uint16_t* tbl) {
  for (int e = 0; e < 450; e++) tbl[e] = (uint16_t)(IC_ZERO << 2);
  auto blk = [&](int r, int c, int off, int code) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) tbl[(r + i) * 30 + c + j] = (uint16_t)(((off + 3 * i + j) << 2) | code); };
  auto skw = [&](int r, int c, int off) {   // skew(v) = [0 -z y; z 0 -x; -y x 0]
    const int idx[3][3] = {{-1, 2, 1}, {2, -1, 0}, {1, 0, -1}}, neg[3][3] = {{0, 1, 0}, {0, 0, 1}, {1, 0, 0}};
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) if (idx[i][j] >= 0) tbl[(r + i) * 30 + c + j] = (uint16_t)(((off + idx[i][j]) << 2) | neg[i][j]); };
  auto eye_ = [&](int r, int c, int code) { for (int i = 0; i < 3; i++) tbl[(r + i) * 30 + c + i] = (uint16_t)((IC_ONE << 2) | code); };
  blk(O_P, 0, IC_RIT, 1); skw(O_P, 3, IC_AP); blk(O_R, 3, IC_LR, 1); skw(O_V, 3, IC_AV);                      // pose_i   (imu_factor.h:88-113)
  blk(O_P, 6, IC_RIT, 2); blk(O_P, 9, IC_JAC, 1); blk(O_P, 12, IC_JAC + 9, 1); blk(O_R, 12, IC_M7, 1);        // speed-bias_i (:114-142)
  blk(O_V, 6, IC_RIT, 1); blk(O_V, 9, IC_JAC + 18, 1); blk(O_V, 12, IC_JAC + 27, 1); eye_(O_BA, 9, 1); eye_(O_BG, 12, 1);
  blk(O_P, 15, IC_RIT, 0); blk(O_R, 18, IC_LQE, 0);                                                            // pose_j   (:143-161)
  blk(O_V, 21, IC_RIT, 0); eye_(O_BA, 24, 0); eye_(O_BG, 27, 0);                                                // speed-bias_j (:162-177)
}
VHD double imu_tbl_value(uint16_t d, const double* core) {
  const int code = d & 3; const double v = core[d >> 2];
  return code == 0 ? v : (code == 1 ? -1.0 * v : -core[IC_DT] * v);
}

// The same arithmetic as imu_eval_raw, split into IMU_PARTS independent pieces so that the lanes of a warp can each
// produce one 3x3 block (part 0..17) or the residual (part 18) after the shared quaternion algebra.
constexpr int IMU_PARTS = 19;
VHD void imu_eval_part(int part, const double* pre, const double* G3, const double* pose_i, const double* sb_i, const double* pose_j,
                       const double* sb_j, double* r, double* J) {
  const v3 Pi = ld3(pose_i), Pj = ld3(pose_j), Vi = ld3(sb_i), Vj = ld3(sb_j), Bgi = ld3(sb_i + 6);
  const q4 Qi = ldq(pose_i + 3), Qj = ldq(pose_j + 3);
  const v3 G = ld3(G3);
  const q4 dq = ldq(pre + 3);
  const double dt = pre[16];
  const double* jac = pre + 17;
  const q4 Qi_inv = qinv(Qi);
  const m3 dq_dbg = blk_cm15(jac, O_R, O_BG);
  const v3 th = mul(dq_dbg, Bgi - ld3(pre + 13));
  const q4 cdq = qmul(dq, mkq(1.0, th.x * 0.5, th.y * 0.5, th.z * 0.5));
  switch (part) {
    case 0: put33(J, 30, O_P, 0, q2R(Qi_inv), -1.0); break;
    case 1: put33(J, 30, O_P, 3, skew(qrot(Qi_inv, G * (0.5 * dt * dt) + Pj - Pi - Vi * dt))); break;
    case 2: {
      const q4 qji = qmul(qinv(Qj), Qi);
      const m3 L = qleft33(qji), R = qright33(cdq);
      m3 LR = mul(L, R);
      const double lcol[3] = {qji.x, qji.y, qji.z}, rrow[3] = {-cdq.x, -cdq.y, -cdq.z};
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) LR.m[i][j] += lcol[i] * rrow[j];
      put33(J, 30, O_R, 3, LR, -1.0);
    } break;
    case 3: put33(J, 30, O_V, 3, skew(qrot(Qi_inv, G * dt + Vj - Vi))); break;
    case 4: put33(J, 30, O_P, 6, q2R(Qi_inv), -dt); break;
    case 5: put33(J, 30, O_P, 9, blk_cm15(jac, O_P, O_BA), -1.0); break;
    case 6: put33(J, 30, O_P, 12, blk_cm15(jac, O_P, O_BG), -1.0); break;
    case 7: put33(J, 30, O_R, 12, mul(qleft33(qmul(qmul(qinv(Qj), Qi), dq)), dq_dbg), -1.0); break;
    case 8: put33(J, 30, O_V, 6, q2R(Qi_inv), -1.0); break;
    case 9: put33(J, 30, O_V, 9, blk_cm15(jac, O_V, O_BA), -1.0); break;
    case 10: put33(J, 30, O_V, 12, blk_cm15(jac, O_V, O_BG), -1.0); break;
    case 11: put33(J, 30, O_BA, 9, eye(), -1.0); break;
    case 12: put33(J, 30, O_BG, 12, eye(), -1.0); break;
    case 13: put33(J, 30, O_P, 15, q2R(Qi_inv)); break;
    case 14: put33(J, 30, O_R, 18, qleft33(qmul(qinv(cdq), qmul(Qi_inv, Qj)))); break;
    case 15: put33(J, 30, O_V, 21, q2R(Qi_inv)); break;
    case 16: put33(J, 30, O_BA, 24, eye()); break;
    case 17: put33(J, 30, O_BG, 27, eye()); break;
    default: imu_eval_raw(pre, G3, pose_i, sb_i, pose_j, sb_j, r, nullptr); break;
  }
}

// ---- LiDAR point factors attached to keyframe k through the fixed LiDAR<->body extrinsic.
// pb = RLB^T (p_l - TLB) is precomputed at pack time; p_w = R_k pb + P_k.
// LidarPlaneNormFactor (lidar_mapping/src/lidarFactor.hpp:113-125): r = n . p_w + d.  J: 1 x 6.
VHD double plane_eval(const double* pose, const v3& pb, const v3& n, double d, double* J) {
  const m3 R = q2R(ldq(pose + 3));
  const v3 pw = mul(R, pb) + ld3(pose);
  if (J) { const v3 jr = cross(pb, mulT(R, n)); J[0] = n.x; J[1] = n.y; J[2] = n.z; J[3] = jr.x; J[4] = jr.y; J[5] = jr.z; }
  return dot(n, pw) + d; }
// LidarEdgeFactor with s = 1 (lidarFactor.hpp:18-43; only call site localMapping.cpp:664):
// r = (p_w - a) x (p_w - b) / |a - b|.  J: 3 x 6 row-major = skew(b - a)/|a-b| [I | -R skew(pb)].
VHD void edge_eval(const double* pose, const v3& pb, const v3& a, const v3& b, double* r, double* J) {
  const m3 R = q2R(ldq(pose + 3));
  const v3 pw = mul(R, pb) + ld3(pose);
  const v3 de = a - b;
  const double inv = 1.0 / sqrt(dot(de, de));
  const v3 nu = cross(pw - a, pw - b);
  r[0] = nu.x * inv; r[1] = nu.y * inv; r[2] = nu.z * inv;
  if (!J) return;
  const m3 S = scale(skew(b - a), inv);
  const m3 Jr = scale(mul(mul(S, R), skew(pb)), -1.0);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) { J[i * 6 + j] = S.m[i][j]; J[i * 6 + 3 + j] = Jr.m[i][j]; }
}

// ---- the reference's two ceres::AutoDiffCostFunction functors, differentiated with forward-mode duals.
// Their pose blocks carry PoseLocalParameterization, so the kept columns are d r / d (px py pz qx qy qz) of the raw
// 7-vector (pose_local_parameterization.cpp:20-27) — reproduced as is.
template <int N> VHD void seed_pose(const double* p, int base, V3<Dual<N>>& P, Q4<Dual<N>>& Q) {
  P = mk(seed<N>(p[0], base), seed<N>(p[1], base + 1), seed<N>(p[2], base + 2));
  Q = mkq(Dual<N>(p[6]), seed<N>(p[3], base + 3), seed<N>(p[4], base + 4), seed<N>(p[5], base + 5)); }
// LPSConstraint::operator() (vils_estimator/src/lidar_backend.h:45-80). c = tl tr tk qx qy qz qw. r: 3, J: 3 x 12.
VHD void lps_eval(const double* c, const double* pa, const double* pb, double* r, double* J) {
  typedef Dual<12> D;
  V3<D> Pa, Pb; Q4<D> Qa, Qb;
  seed_pose<12>(pa, 0, Pa, Qa); seed_pose<12>(pb, 6, Pb, Qb);
  const double t = (c[2] - c[0]) / (c[1] - c[0]);
  const Q4<D> Qi = qslerp(Qa, t, Qb);
  const Q4<D> Q12 = qmul(qinv(Qi), mkq(D(c[6]), D(c[3]), D(c[4]), D(c[5])));
  const D out[3] = {Q12.x * D(2.0) / D(0.01), Q12.y * D(2.0) / D(0.01), Q12.z * D(2.0) / D(0.01)};
  for (int i = 0; i < 3; i++) { r[i] = out[i].a; if (J) for (int k = 0; k < 12; k++) J[i * 12 + k] = out[i].v[k]; }
}
// LidarICPConstraint_b::operator() (lidar_backend.h:107-169). c = ta tb tc td ti tj trans_t(3) sqrt_info. r: 3, J: 3 x 24.
VHD void icp_eval(const double* c, const double* pa, const double* pb, const double* pc, const double* pd, double* r, double* J) {
  typedef Dual<24> D;
  V3<D> Pa, Pb, Pc, Pd; Q4<D> Qa, Qb, Qc, Qd;
  seed_pose<24>(pa, 0, Pa, Qa); seed_pose<24>(pb, 6, Pb, Qb); seed_pose<24>(pc, 12, Pc, Qc); seed_pose<24>(pd, 18, Pd, Qd);
  const double t_i = (c[4] - c[0]) / (c[1] - c[0]), t_j = (c[5] - c[2]) / (c[3] - c[2]);
  const Q4<D> Qi = qslerp(Qa, t_i, Qb), Qj = qslerp(Qc, t_j, Qd);
  const V3<D> Pi = Pa + (Pb - Pa) * (D(1.0) / D(c[1] - c[0])) * D(c[4] - c[0]);
  const V3<D> Pj = Pc + (Pd - Pc) * (D(1.0) / D(c[3] - c[2])) * D(c[5] - c[2]);
  const Q4<D> temQ = qmul(qinv(Qj), Qi);
  const V3<D> temP = qrot(qinv(Qi), Pj - Pi);
  const V3<D> PIJ = mk(D(c[6]), D(c[7]), D(c[8]));
  const V3<D> RES = qrot(temQ, PIJ - temP);
  const D out[3] = {RES.x * D(c[9]), D(0.0), RES.z * D(c[9])};
  for (int i = 0; i < 3; i++) { r[i] = out[i].a; if (J) for (int k = 0; k < 24; k++) J[i * 24 + k] = out[i].v[k]; }
}

// ---- MarginalizationFactor::Evaluate dx (factor/marginalization_factor.cpp:364-383) for one pose block ----
VHD void prior_dx_pose(const double* x, const double* x0, double* dx) {
  dx[0] = x[0] - x0[0]; dx[1] = x[1] - x0[1]; dx[2] = x[2] - x0[2];
  const q4 d = qmul(qinv(ldq(x0 + 3)), ldq(x + 3));
  const double s = (d.w >= 0) ? 2.0 : -2.0;
  dx[3] = s * d.x; dx[4] = s * d.y; dx[5] = s * d.z; }

}  // namespace vf
