// vils_oracle.cpp — CPU FP64 restatement of the reference hot path (Stan994265/mVIL-Fusion).
//
// *** TEST INFRASTRUCTURE ONLY. ***  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library; the product (mvil_fusion_b200/) never does.
//
// PARITY STATUS: "parity unpinned" against Ceres/Eigen.  The reference cannot be compiled here (every
// hot-path translation unit includes <ros/...>, <ceres/ceres.h>, Eigen; none exist in this image) and
// it ships no tests or golden vectors.  This file restates the reference arithmetic line by line from
// the files cited at each function; what pins it is (i) forward-difference Jacobian checks modelled on
// ProjectionFactor::check (factor/projection_factor.cpp:123-225), (ii) the marginalization identities
// J^T J = A, J^T r = b (factor/marginalization_factor.cpp:313-314), (iii) zero-residual fixed points on
// noise-free synthetic windows, (iv) scipy.optimize.least_squares on the same residuals
// (tests/test_oracle_*.py).  Third-party algorithms whose source is not in the reference tree and that
// are restated from their published behaviour: ceres-solver (loss functions, robust corrector, Jet
// autodiff, LM trust region; version unpinned by the reference, API implies <= 2.1) and Eigen 3.3
// (quaternion slerp / toRotationMatrix, LLT, inverse, SelfAdjointEigenSolver -> Jacobi here).
//
// All paths below are relative to the reference tree.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../include/vils_cabi.h"
#include "vo_math.h"

using namespace vo;

namespace {

// StateOrder (vils_estimator/src/parameters.h:80-87)
enum { O_P = 0, O_R = 3, O_V = 6, O_BA = 9, O_BG = 12 };

inline Vec3 v3(const double* p) { return Vec3(p[0], p[1], p[2]); }
inline Quat qxyzw(const double* p) { return Quat(p[3], p[0], p[1], p[2]); }  // p -> (x y z w)
inline Quat pose_q(const double* pose) { return Quat(pose[6], pose[3], pose[4], pose[5]); }  // imu_factor.h:23

struct M15 {  // 15x15 row-major
  double a[15][15];
  M15() { std::memset(a, 0, sizeof(a)); }
};
inline void set33(double* base, int ld, int r, int c, const Mat3& m) {
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) base[(r + i) * ld + c + j] = m(i, j);
}
inline Mat3 get33cm(const double* cm15, int r, int c) {  // from column-major 15x15
  Mat3 m; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m(i, j) = cm15[(c + j) * 15 + r + i]; return m;
}

// ------------------------------------------------------------------------------------------
// IntegrationBase::midPointIntegration + propagate (factor/integration_base.h:54-158)
// ------------------------------------------------------------------------------------------
struct Preint {
  Vec3 acc_0, gyr_0, ba, bg, dp, dv;
  Quat dq;
  double sum_dt;
  double J[15][15], P[15][15];
  double noise[18];
};

void preint_init(Preint& s, Vec3 acc0, Vec3 gyr0, Vec3 ba, Vec3 bg, const double n[4]) {
  s.acc_0 = acc0; s.gyr_0 = gyr0; s.ba = ba; s.bg = bg; s.dp = Vec3(); s.dv = Vec3(); s.dq = Quat(); s.sum_dt = 0;
  std::memset(s.J, 0, sizeof(s.J)); std::memset(s.P, 0, sizeof(s.P));
  for (int i = 0; i < 15; i++) s.J[i][i] = 1.0;  // :17
  // :21-27  ACC_N GYR_N ACC_N GYR_N ACC_W GYR_W
  const double d[6] = {n[0] * n[0], n[1] * n[1], n[0] * n[0], n[1] * n[1], n[2] * n[2], n[3] * n[3]};
  for (int i = 0; i < 18; i++) s.noise[i] = d[i / 3];
}

void preint_propagate(Preint& s, double dt, Vec3 acc_1, Vec3 gyr_1) {
  // :63-71
  Vec3 un_acc_0 = rotate(s.dq, s.acc_0 - s.ba);
  Vec3 un_gyr = 0.5 * (s.gyr_0 + gyr_1) - s.bg;
  Quat rq = s.dq * Quat(1, un_gyr.x * dt / 2, un_gyr.y * dt / 2, un_gyr.z * dt / 2);
  Vec3 un_acc_1 = rotate(rq, acc_1 - s.ba);
  Vec3 un_acc = 0.5 * (un_acc_0 + un_acc_1);
  Vec3 rp = s.dp + s.dv * dt + 0.5 * un_acc * dt * dt;
  Vec3 rv = s.dv + un_acc * dt;
  // :75-88
  Vec3 w_x = 0.5 * (s.gyr_0 + gyr_1) - s.bg;
  Vec3 a_0_x = s.acc_0 - s.ba, a_1_x = acc_1 - s.ba;
  Mat3 R_w_x = skew(w_x), R_a_0_x = skew(a_0_x), R_a_1_x = skew(a_1_x);
  Mat3 Rd = toR(s.dq), Rr = toR(rq), I3 = Mat3::I();
  // :90-105
  static thread_local double F[15][15], V[15][18];
  std::memset(F, 0, sizeof(F)); std::memset(V, 0, sizeof(V));
  double* Fp = &F[0][0]; double* Vp = &V[0][0];
  set33(Fp, 15, 0, 0, I3);
  set33(Fp, 15, 0, 3, (Rd * R_a_0_x) * (-0.25 * dt * dt) + (Rr * R_a_1_x * (I3 - R_w_x * dt)) * (-0.25 * dt * dt));
  set33(Fp, 15, 0, 6, I3 * dt);
  set33(Fp, 15, 0, 9, (Rd + Rr) * (-0.25 * dt * dt));
  set33(Fp, 15, 0, 12, (Rr * R_a_1_x) * (-0.25 * dt * dt * -dt));
  set33(Fp, 15, 3, 3, I3 - R_w_x * dt);
  set33(Fp, 15, 3, 12, I3 * (-1.0 * dt));
  set33(Fp, 15, 6, 3, (Rd * R_a_0_x) * (-0.5 * dt) + (Rr * R_a_1_x * (I3 - R_w_x * dt)) * (-0.5 * dt));
  set33(Fp, 15, 6, 6, I3);
  set33(Fp, 15, 6, 9, (Rd + Rr) * (-0.5 * dt));
  set33(Fp, 15, 6, 12, (Rr * R_a_1_x) * (-0.5 * dt * -dt));
  set33(Fp, 15, 9, 9, I3);
  set33(Fp, 15, 12, 12, I3);
  // :108-120
  Mat3 V03 = (-Rr * R_a_1_x) * (0.25 * dt * dt * 0.5 * dt);
  Mat3 V63 = (-Rr * R_a_1_x) * (0.5 * dt * 0.5 * dt);
  set33(Vp, 18, 0, 0, Rd * (0.25 * dt * dt));
  set33(Vp, 18, 0, 3, V03);
  set33(Vp, 18, 0, 6, Rr * (0.25 * dt * dt));
  set33(Vp, 18, 0, 9, V03);
  set33(Vp, 18, 3, 3, I3 * (0.5 * dt));
  set33(Vp, 18, 3, 9, I3 * (0.5 * dt));
  set33(Vp, 18, 6, 0, Rd * (0.5 * dt));
  set33(Vp, 18, 6, 3, V63);
  set33(Vp, 18, 6, 6, Rr * (0.5 * dt));
  set33(Vp, 18, 6, 9, V63);
  set33(Vp, 18, 9, 12, I3 * dt);
  set33(Vp, 18, 12, 15, I3 * dt);
  // :124-125  jacobian = F*jacobian ; covariance = F*cov*F^T + V*noise*V^T
  double T[15][15], T2[15][15];
  for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) { double a = 0; for (int k = 0; k < 15; k++) a += F[i][k] * s.J[k][j]; T[i][j] = a; }
  std::memcpy(s.J, T, sizeof(T));
  for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) { double a = 0; for (int k = 0; k < 15; k++) a += F[i][k] * s.P[k][j]; T[i][j] = a; }
  for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) { double a = 0; for (int k = 0; k < 15; k++) a += T[i][k] * F[j][k]; T2[i][j] = a; }
  for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) { double a = 0; for (int k = 0; k < 18; k++) a += V[i][k] * s.noise[k] * V[j][k]; T2[i][j] += a; }
  std::memcpy(s.P, T2, sizeof(T2));
  // :148-156
  s.dp = rp; s.dq = normalized(rq); s.dv = rv;
  s.sum_dt += dt; s.acc_0 = acc_1; s.gyr_0 = gyr_1;
}

void preint_export(const Preint& s, vils_preint* o) {
  o->delta_p[0] = s.dp.x; o->delta_p[1] = s.dp.y; o->delta_p[2] = s.dp.z;
  o->delta_q[0] = s.dq.x; o->delta_q[1] = s.dq.y; o->delta_q[2] = s.dq.z; o->delta_q[3] = s.dq.w;
  o->delta_v[0] = s.dv.x; o->delta_v[1] = s.dv.y; o->delta_v[2] = s.dv.z;
  o->lin_ba[0] = s.ba.x; o->lin_ba[1] = s.ba.y; o->lin_ba[2] = s.ba.z;
  o->lin_bg[0] = s.bg.x; o->lin_bg[1] = s.bg.y; o->lin_bg[2] = s.bg.z;
  o->sum_dt = s.sum_dt;
  for (int r = 0; r < 15; r++) for (int c = 0; c < 15; c++) { o->jacobian[c * 15 + r] = s.J[r][c]; o->covariance[c * 15 + r] = s.P[r][c]; }
}

// ------------------------------------------------------------------------------------------
// IMUFactor::Evaluate (factor/imu_factor.h:19-181) + IntegrationBase::evaluate (integration_base.h:175-201)
// Jacobians: row-major 15x7, 15x9, 15x7, 15x9 (global size, 7th pose column zero), any may be null.
// ------------------------------------------------------------------------------------------
void imu_sqrt_info(const vils_preint* pre, double W[15][15]) {
  // :64  LLT(covariance.inverse()).matrixL().transpose()
  double cov[225], inv[225];
  for (int r = 0; r < 15; r++) for (int c = 0; c < 15; c++) cov[r * 15 + c] = pre->covariance[c * 15 + r];
  inverse_lu(cov, 15, inv);
  cholesky_lower(inv, 15);
  for (int r = 0; r < 15; r++) for (int c = 0; c < 15; c++) W[r][c] = inv[c * 15 + r];
}

void imu_evaluate(const vils_preint* pre, const double* G3, const double* pi, const double* sbi, const double* pj,
                  const double* sbj, double* residuals, double* Jpi, double* Jsbi, double* Jpj, double* Jsbj) {
  Vec3 Pi = v3(pi), Pj = v3(pj); Quat Qi = pose_q(pi), Qj = pose_q(pj);
  Vec3 Vi = v3(sbi), Bai = v3(sbi + 3), Bgi = v3(sbi + 6), Vj = v3(sbj), Baj = v3(sbj + 3), Bgj = v3(sbj + 6);
  Vec3 G = v3(G3);
  const double sum_dt = pre->sum_dt;
  Vec3 delta_p = v3(pre->delta_p), delta_v = v3(pre->delta_v); Quat delta_q = qxyzw(pre->delta_q);
  Vec3 lba = v3(pre->lin_ba), lbg = v3(pre->lin_bg);
  // integration_base.h:180-199
  Mat3 dp_dba = get33cm(pre->jacobian, O_P, O_BA), dp_dbg = get33cm(pre->jacobian, O_P, O_BG);
  Mat3 dq_dbg = get33cm(pre->jacobian, O_R, O_BG);
  Mat3 dv_dba = get33cm(pre->jacobian, O_V, O_BA), dv_dbg = get33cm(pre->jacobian, O_V, O_BG);
  Vec3 dba = Bai - lba, dbg = Bgi - lbg;
  Quat corrected_delta_q = delta_q * deltaQ(dq_dbg * dbg);
  Vec3 corrected_delta_v = delta_v + dv_dba * dba + dv_dbg * dbg;
  Vec3 corrected_delta_p = delta_p + dp_dba * dba + dp_dbg * dbg;
  double r[15];
  Vec3 rp = rotate(inverse(Qi), 0.5 * G * sum_dt * sum_dt + Pj - Pi - Vi * sum_dt) - corrected_delta_p;
  Vec3 rq = 2.0 * (inverse(corrected_delta_q) * (inverse(Qi) * Qj)).vec();
  Vec3 rv = rotate(inverse(Qi), G * sum_dt + Vj - Vi) - corrected_delta_v;
  Vec3 rba = Baj - Bai, rbg = Bgj - Bgi;
  for (int k = 0; k < 3; k++) { r[O_P + k] = rp[k]; r[O_R + k] = rq[k]; r[O_V + k] = rv[k]; r[O_BA + k] = rba[k]; r[O_BG + k] = rbg[k]; }
  double W[15][15];
  imu_sqrt_info(pre, W);
  for (int i = 0; i < 15; i++) { double a = 0; for (int k = 0; k < 15; k++) a += W[i][k] * r[k]; residuals[i] = a; }

  auto premul = [&](double* J, int cols) {  // J = sqrt_info * J
    std::vector<double> T(15 * cols);
    for (int i = 0; i < 15; i++) for (int j = 0; j < cols; j++) { double a = 0; for (int k = 0; k < 15; k++) a += W[i][k] * J[k * cols + j]; T[i * cols + j] = a; }
    std::memcpy(J, T.data(), sizeof(double) * 15 * cols);
  };
  Mat3 RiT = toR(inverse(Qi));
  if (Jpi) {  // :88-113
    std::memset(Jpi, 0, sizeof(double) * 15 * 7);
    set33(Jpi, 7, O_P, O_P, -RiT);
    set33(Jpi, 7, O_P, O_R, skew(rotate(inverse(Qi), 0.5 * G * sum_dt * sum_dt + Pj - Pi - Vi * sum_dt)));
    double L[4][4], R[4][4];
    Qleft44(inverse(Qj) * Qi, L); Qright44(corrected_delta_q, R);
    Mat3 LR; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double a = 0; for (int k = 0; k < 4; k++) a += L[1 + i][k] * R[k][1 + j]; LR(i, j) = a; }
    set33(Jpi, 7, O_R, O_R, -LR);
    set33(Jpi, 7, O_V, O_R, skew(rotate(inverse(Qi), G * sum_dt + Vj - Vi)));
    premul(Jpi, 7);
  }
  if (Jsbi) {  // :114-142
    std::memset(Jsbi, 0, sizeof(double) * 15 * 9);
    set33(Jsbi, 9, O_P, O_V - O_V, RiT * (-sum_dt));
    set33(Jsbi, 9, O_P, O_BA - O_V, -dp_dba);
    set33(Jsbi, 9, O_P, O_BG - O_V, -dp_dbg);
    set33(Jsbi, 9, O_R, O_BG - O_V, -(Qleft33(inverse(Qj) * Qi * delta_q) * dq_dbg));  // :127 uses UNcorrected delta_q
    set33(Jsbi, 9, O_V, O_V - O_V, -RiT);
    set33(Jsbi, 9, O_V, O_BA - O_V, -dv_dba);
    set33(Jsbi, 9, O_V, O_BG - O_V, -dv_dbg);
    set33(Jsbi, 9, O_BA, O_BA - O_V, -Mat3::I());
    set33(Jsbi, 9, O_BG, O_BG - O_V, -Mat3::I());
    premul(Jsbi, 9);
  }
  if (Jpj) {  // :143-161
    std::memset(Jpj, 0, sizeof(double) * 15 * 7);
    set33(Jpj, 7, O_P, O_P, RiT);
    set33(Jpj, 7, O_R, O_R, Qleft33(inverse(corrected_delta_q) * inverse(Qi) * Qj));
    premul(Jpj, 7);
  }
  if (Jsbj) {  // :162-177
    std::memset(Jsbj, 0, sizeof(double) * 15 * 9);
    set33(Jsbj, 9, O_V, O_V - O_V, RiT);
    set33(Jsbj, 9, O_BA, O_BA - O_V, Mat3::I());
    set33(Jsbj, 9, O_BG, O_BG - O_V, Mat3::I());
    premul(Jsbj, 9);
  }
}

// ------------------------------------------------------------------------------------------
// ProjectionTdFactor::Evaluate (factor/projection_td_factor.cpp:34-141) and
// ProjectionFactor::Evaluate (factor/projection_factor.cpp:21-121; = the td variant with pts unshifted and no J_td).
// row_i/row_j are uv.y; the ctor subtracts ROW/2 (:18-19).
// ------------------------------------------------------------------------------------------
void proj_evaluate(const vils_config* cfg, int use_td, const double* pts_i3, const double* pts_j3, const double* vel_i2,
                   const double* vel_j2, double td_i, double td_j, double row_i_uv, double row_j_uv, const double* pi,
                   const double* pj, const double* pex, double inv_dep_i, double td, double* residual, double* Ji,
                   double* Jj, double* Jex, double* Jf, double* Jtd) {
  const double s_info = cfg->focal_length / 2.0;  // estimator.cpp:18-19
  Vec3 Pi = v3(pi), Pj = v3(pj), tic = v3(pex); Quat Qi = pose_q(pi), Qj = pose_q(pj), qic = pose_q(pex);
  Vec3 pts_i = v3(pts_i3), pts_j = v3(pts_j3);
  Vec3 velocity_i(vel_i2[0], vel_i2[1], 0), velocity_j(vel_j2[0], vel_j2[1], 0);  // :12-17
  Vec3 pts_i_td = pts_i, pts_j_td = pts_j;
  if (use_td) {
    double row_i = row_i_uv - cfg->row / 2, row_j = row_j_uv - cfg->row / 2;
    pts_i_td = pts_i - (td - td_i + cfg->tr / cfg->row * row_i) * velocity_i;  // :51-52
    pts_j_td = pts_j - (td - td_j + cfg->tr / cfg->row * row_j) * velocity_j;
  }
  Vec3 pts_camera_i = pts_i_td / inv_dep_i;
  Vec3 pts_imu_i = rotate(qic, pts_camera_i) + tic;
  Vec3 pts_w = rotate(Qi, pts_imu_i) + Pi;
  Vec3 pts_imu_j = rotate(inverse(Qj), pts_w - Pj);
  Vec3 pts_camera_j = rotate(inverse(qic), pts_imu_j - tic);
  double dep_j = pts_camera_j.z;
  residual[0] = s_info * (pts_camera_j.x / dep_j - pts_j_td.x);  // :63-67
  residual[1] = s_info * (pts_camera_j.y / dep_j - pts_j_td.y);
  if (!(Ji || Jj || Jex || Jf || Jtd)) return;
  Mat3 Ri = toR(Qi), Rj = toR(Qj), ric = toR(qic);
  double reduce[2][3] = {{s_info * 1. / dep_j, 0, s_info * -pts_camera_j.x / (dep_j * dep_j)},
                         {0, s_info * 1. / dep_j, s_info * -pts_camera_j.y / (dep_j * dep_j)}};  // :87-90
  auto red = [&](const Mat3& A, const Mat3& B, double* out /*2x7 row-major*/) {
    for (int r = 0; r < 2; r++) { for (int c = 0; c < 3; c++) { double a = 0, b = 0; for (int k = 0; k < 3; k++) { a += reduce[r][k] * A(k, c); b += reduce[r][k] * B(k, c); } out[r * 7 + c] = a; out[r * 7 + 3 + c] = b; } out[r * 7 + 6] = 0; }
  };
  Mat3 ricT = transpose(ric), RjT = transpose(Rj);
  if (Ji) red(ricT * RjT, ricT * RjT * Ri * (-skew(pts_imu_i)), Ji);  // :92-102
  if (Jj) red(ricT * (-RjT), ricT * skew(pts_imu_j), Jj);             // :104-114
  Mat3 tmp_r = ricT * RjT * Ri * ric;
  if (Jex) {  // :115-125
    Mat3 A = ricT * (RjT * Ri - Mat3::I());
    Mat3 B = -tmp_r * skew(pts_camera_i) + skew(tmp_r * pts_camera_i) + skew(ricT * (RjT * (Ri * tic + Pi - Pj) - tic));
    red(A, B, Jex);
  }
  if (Jf) {  // :126-130
    Vec3 v = tmp_r * pts_i_td * (-1.0 / (inv_dep_i * inv_dep_i));
    for (int r = 0; r < 2; r++) Jf[r] = reduce[r][0] * v.x + reduce[r][1] * v.y + reduce[r][2] * v.z;
  }
  if (Jtd) {  // :131-136
    if (use_td) {
      Vec3 v = tmp_r * velocity_i / inv_dep_i * -1.0;
      for (int r = 0; r < 2; r++) Jtd[r] = reduce[r][0] * v.x + reduce[r][1] * v.y + reduce[r][2] * v.z + s_info * velocity_j[r];
    } else { Jtd[0] = Jtd[1] = 0; }
  }
}

// ------------------------------------------------------------------------------------------
// Local-parameterisation Jacobians.
//  * reference behaviour for autodiff factors: PoseLocalParameterization::ComputeJacobian returns
//    [I6;0] (factor/pose_local_parameterization.cpp:20-27), so Ceres keeps the first six columns
//    of the 7-wide global Jacobian.
//  * true derivative of Plus (pose_local_parameterization.cpp:3-19) at delta = 0:
//    d p / d dp = I ; d q(xyzw) / d dtheta = 1/2 * Qleft(q)[:,1:4] re-ordered.
// ------------------------------------------------------------------------------------------
void plus_jacobian_true(const double* pose7, double Jp[7][6]) {
  std::memset(Jp, 0, sizeof(double) * 42);
  for (int i = 0; i < 3; i++) Jp[i][i] = 1;
  Quat q = pose_q(pose7);
  // q (x) [0, e_k/2]: w' = -q.vec . e/2 ; vec' = (w e + q.vec x e)/2
  for (int k = 0; k < 3; k++) {
    Vec3 e; e[k] = 0.5;
    Vec3 dv = e * q.w + cross(q.vec(), e);
    double dw = -dot(q.vec(), e);
    Jp[3][3 + k] = dv.x; Jp[4][3 + k] = dv.y; Jp[5][3 + k] = dv.z; Jp[6][3 + k] = dw;
  }
}

template <class T> inline QuatT<T> jq(const T* p) { return QuatT<T>(p[6], p[3], p[4], p[5]); }
template <class T> inline Vec3T<T> jp(const T* p) { return Vec3T<T>(p[0], p[1], p[2]); }

// LidarPlaneNormFactor::operator() (lidar_mapping/src/lidarFactor.hpp:113-125), attached to a window keyframe:
// (q_w_curr, t_w_curr) = LiDAR pose of keyframe k = (Q_k * RLB^T, P_k + Q_k * TBL)  (estimator.cpp:451,484-485).
template <class T>
void plane_functor(const vils_config* cfg, const T* pose, const double* p_l, const double* n, double d, T* res) {
  Mat3 RLB; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) RLB(i, j) = cfg->rlb[i * 3 + j];
  Vec3 TLB = v3(cfg->tlb);
  Vec3 pb = transpose(RLB) * (v3(p_l) - TLB);  // p_b = RLB^T (p_l - TLB)
  Vec3T<T> cp{T(pb.x), T(pb.y), T(pb.z)};
  Vec3T<T> point_w = rotate(jq(pose), cp) + jp(pose);
  Vec3T<T> nrm{T(n[0]), T(n[1]), T(n[2])};
  res[0] = dot(nrm, point_w) + T(d);
}
// LidarEdgeFactor::operator() (lidarFactor.hpp:18-43) with s = 1.0 (the only call site, localMapping.cpp:664):
// q_identity.slerp(1, q) == q including derivatives, t_last_curr = t.
template <class T>
void edge_functor(const vils_config* cfg, const T* pose, const double* p_l, const double* a, const double* b, T* res) {
  Mat3 RLB; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) RLB(i, j) = cfg->rlb[i * 3 + j];
  Vec3 TLB = v3(cfg->tlb);
  Vec3 pb = transpose(RLB) * (v3(p_l) - TLB);
  Vec3T<T> cp{T(pb.x), T(pb.y), T(pb.z)};
  Vec3T<T> lp = rotate(jq(pose), cp) + jp(pose);
  Vec3T<T> lpa{T(a[0]), T(a[1]), T(a[2])}, lpb{T(b[0]), T(b[1]), T(b[2])};
  Vec3T<T> nu = cross(lp - lpa, lp - lpb);
  Vec3T<T> de = lpa - lpb;
  T den = norm(de);
  res[0] = nu.x / den; res[1] = nu.y / den; res[2] = nu.z / den;
}
// LPSConstraint::operator() (vils_estimator/src/lidar_backend.h:45-80). Qi.normalized() discards its result (:53).
template <class T>
void lps_functor(const vils_lps* c, const T* POSEa, const T* POSEb, T* residuals) {
  QuatT<T> Qa = jq(POSEa), Qb = jq(POSEb);
  double t_i = (c->tk - c->tl) / (c->tr - c->tl);
  QuatT<T> Qi = slerp(Qa, T(t_i), Qb);
  QuatT<T> Q1{T(c->q[3]), T(c->q[0]), T(c->q[1]), T(c->q[2])};
  QuatT<T> Q12 = inverse(Qi) * Q1;
  residuals[0] = T(2) * Q12.x / T(0.01);
  residuals[1] = T(2) * Q12.y / T(0.01);
  residuals[2] = T(2) * Q12.z / T(0.01);
}
// LidarICPConstraint_b::operator() (lidar_backend.h:107-169)
template <class T>
void icp_functor(const vils_icp* c, const T* POSEa, const T* POSEb, const T* POSEc, const T* POSEd, T* residuals) {
  QuatT<T> Qa = jq(POSEa), Qb = jq(POSEb), Qc = jq(POSEc), Qd = jq(POSEd);
  Vec3T<T> Pa = jp(POSEa), Pb = jp(POSEb), Pc = jp(POSEc), Pd = jp(POSEd);
  double t_i = (c->ti - c->ta) / (c->tb - c->ta);
  double t_j = (c->tj - c->tc) / (c->td - c->tc);
  QuatT<T> Qi = slerp(Qa, T(t_i), Qb);
  QuatT<T> Qj = slerp(Qc, T(t_j), Qd);
  Vec3T<T> Pi = Pa + (Pb - Pa) / T(c->tb - c->ta) * T(c->ti - c->ta);
  Vec3T<T> Pj = Pc + (Pd - Pc) / T(c->td - c->tc) * T(c->tj - c->tc);
  QuatT<T> temQ = inverse(Qj) * Qi;
  Vec3T<T> temPIJ = rotate(inverse(Qi), Pj - Pi);
  Vec3T<T> PIJ{T(c->trans_t[0]), T(c->trans_t[1]), T(c->trans_t[2])};
  Vec3T<T> RES = rotate(temQ, PIJ - temPIJ);
  residuals[0] = RES.x * T(c->sqrt_info);
  residuals[1] = T(0.0);
  residuals[2] = RES.z * T(c->sqrt_info);
}

// ceres::AutoDiffCostFunction over NB pose blocks of 7: residuals (NR) + global Jacobians (NR x 7 per block, row-major).
template <int NR, int NB, class F>
void autodiff(F f, const double* const* poses, double* res, double Jg[NB][NR * 7]) {
  typedef Jet<7 * NB> J;
  J x[NB][7];
  for (int b = 0; b < NB; b++) for (int k = 0; k < 7; k++) x[b][k] = J(poses[b][k], b * 7 + k);
  J r[NR];
  const J* xp[NB];
  for (int b = 0; b < NB; b++) xp[b] = x[b];
  f(xp, r);
  for (int i = 0; i < NR; i++) { res[i] = r[i].a; for (int b = 0; b < NB; b++) for (int k = 0; k < 7; k++) Jg[b][i * 7 + k] = r[i].v[b * 7 + k]; }
}

// ------------------------------------------------------------------------------------------
// Loss functions [upstream Ceres loss_function.cc] and the robust corrector as restated in-tree by
// ResidualBlockInfo::Evaluate (factor/marginalization_factor.cpp:37-68).
// ------------------------------------------------------------------------------------------
enum { LOSS_NONE = 0, LOSS_CAUCHY = 1, LOSS_HUBER = 2 };
void loss_eval(int kind, double a, double s, double rho[3]) {
  if (kind == LOSS_CAUCHY) { double b = a * a, c = 1 / b; double sum = 1 + s * c, inv = 1 / sum; rho[0] = b * std::log(sum); rho[1] = std::max(std::numeric_limits<double>::min(), inv); rho[2] = -c * (inv * inv); }
  else if (kind == LOSS_HUBER) { double b = a * a; if (s > b) { double r = std::sqrt(s); rho[0] = 2 * a * r - b; rho[1] = std::max(std::numeric_limits<double>::min(), a / r); rho[2] = -rho[1] / (2 * s); } else { rho[0] = s; rho[1] = 1; rho[2] = 0; } }
  else { rho[0] = s; rho[1] = 1; rho[2] = 0; }
}
// r: nr ; J: nr x nc row-major. Returns rho(s) (cost contribution is rho/2).
double corrector(int kind, double a, int nr, int nc, double* r, double* J) {
  double sq_norm = 0; for (int i = 0; i < nr; i++) sq_norm += r[i] * r[i];
  if (kind == LOSS_NONE) return sq_norm;
  double rho[3]; loss_eval(kind, a, sq_norm, rho);
  double sqrt_rho1_ = std::sqrt(rho[1]);
  double residual_scaling_, alpha_sq_norm_;
  if ((sq_norm == 0.0) || (rho[2] <= 0.0)) { residual_scaling_ = sqrt_rho1_; alpha_sq_norm_ = 0.0; }  // :49-53
  else { const double D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1]; const double alpha = 1.0 - std::sqrt(D); residual_scaling_ = sqrt_rho1_ / (1 - alpha); alpha_sq_norm_ = alpha / sq_norm; }
  if (J) {  // :62-65  J = sqrt_rho1 * (J - alpha_sq_norm * r * (r^T J))
    std::vector<double> rtJ(nc, 0.0);
    for (int i = 0; i < nr; i++) for (int j = 0; j < nc; j++) rtJ[j] += r[i] * J[i * nc + j];
    for (int i = 0; i < nr; i++) for (int j = 0; j < nc; j++) J[i * nc + j] = sqrt_rho1_ * (J[i * nc + j] - alpha_sq_norm_ * r[i] * rtJ[j]);
  }
  for (int i = 0; i < nr; i++) r[i] *= residual_scaling_;  // :67
  return rho[0];
}

// ------------------------------------------------------------------------------------------
// Window state + generic residual blocks
// ------------------------------------------------------------------------------------------
struct State {
  int N, M;
  dvec pose, sb, ex, lam; double td;
  void from(const vils_window* w) {
    N = w->n_kf; M = w->n_feat; pose.assign(w->pose, w->pose + 7 * N); sb.assign(w->speedbias, w->speedbias + 9 * N);
    ex.assign(w->ex_pose, w->ex_pose + 7); lam.assign(w->inv_depth, w->inv_depth + M); td = w->td;
  }
  int D() const { return 15 * N + 7; }
};
// tangent offsets: pose k -> 15k ; sb k -> 15k+6 ; ex -> 15N ; td -> 15N+6 ; lambda f -> D + f
struct Block { int off, size; };
struct ResBlock {
  int nr; int nb; Block blk[6]; dvec r; dvec J[6];  // J[b]: nr x size row-major
  double rho;                                       // rho(|r|^2) (cost = rho/2)
};

inline void pose_plus(double* x, const double* d) {  // PoseLocalParameterization::Plus (pose_local_parameterization.cpp:3-19)
  x[0] += d[0]; x[1] += d[1]; x[2] += d[2];
  Quat q = normalized(pose_q(x) * deltaQ(Vec3(d[3], d[4], d[5])));
  x[3] = q.x; x[4] = q.y; x[5] = q.z; x[6] = q.w;
}

void state_plus(State& s, const double* dc, const double* dl) {
  for (int k = 0; k < s.N; k++) { pose_plus(&s.pose[7 * k], dc + 15 * k); for (int i = 0; i < 9; i++) s.sb[9 * k + i] += dc[15 * k + 6 + i]; }
  pose_plus(s.ex.data(), dc + 15 * s.N);
  s.td += dc[15 * s.N + 6];
  for (int f = 0; f < s.M; f++) s.lam[f] += dl[f];
}

// MarginalizationFactor::Evaluate (factor/marginalization_factor.cpp:352-400): dx and residual.
void prior_dx(const vils_window* w, const State& s, dvec& dx) {
  int n = w->prior_n; dx.assign(n, 0.0);
  int col = 0; const double* x0 = w->prior_x0;
  for (int b = 0; b < w->prior_nblk; b++) {
    int id = w->prior_blk[b], type = VILS_BLK_TYPE(id), idx = VILS_BLK_INDEX(id);
    const double* x; int size;
    if (type == VILS_BLK_POSE) { x = &s.pose[7 * idx]; size = 7; }
    else if (type == VILS_BLK_SPEEDBIAS) { x = &s.sb[9 * idx]; size = 9; }
    else if (type == VILS_BLK_EXPOSE) { x = s.ex.data(); size = 7; }
    else { x = &s.td; size = 1; }
    if (size != 7) { for (int i = 0; i < size; i++) dx[col + i] = x[i] - x0[i]; col += size; }
    else {
      for (int i = 0; i < 3; i++) dx[col + i] = x[i] - x0[i];
      Quat dq = inverse(pose_q(x0)) * pose_q(x);  // :376
      double sgn = (dq.w >= 0) ? 1.0 : -1.0;      // :377-380
      dx[col + 3] = sgn * 2.0 * dq.x; dx[col + 4] = sgn * 2.0 * dq.y; dx[col + 5] = sgn * 2.0 * dq.z;
      col += 6;
    }
    x0 += size;
  }
}

int blk_local_size(int type) { return type == VILS_BLK_POSE ? 6 : type == VILS_BLK_SPEEDBIAS ? 9 : type == VILS_BLK_EXPOSE ? 6 : 1; }
int blk_global_size(int type) { return type == VILS_BLK_POSE ? 7 : type == VILS_BLK_SPEEDBIAS ? 9 : type == VILS_BLK_EXPOSE ? 7 : 1; }
int blk_offset(int id, int N) {
  int type = VILS_BLK_TYPE(id), idx = VILS_BLK_INDEX(id);
  return type == VILS_BLK_POSE ? 15 * idx : type == VILS_BLK_SPEEDBIAS ? 15 * idx + 6 : type == VILS_BLK_EXPOSE ? 15 * N : 15 * N + 6;
}

inline void take6(const double* J7, int nr, dvec& out) { out.resize(nr * 6); for (int i = 0; i < nr; i++) for (int j = 0; j < 6; j++) out[i * 6 + j] = J7[i * 7 + j]; }

// Which factors enter and which carry a loss: estimator.cpp:1171-1398.
//   prior: no loss (:1175) ; IMU: no loss (:1185) ; projection: CauchyLoss(1.0) (:1215) ; LPS: Cauchy (:1322) ;
//   ICP: Cauchy (:1395) ; LiDAR edge/plane: HuberLoss(0.1) (lidar_mapping/src/localMapping.cpp:597).
// `which` selects factor families (bitmask) so marginalization can reuse this.
enum { F_IMU = 1, F_PROJ = 2, F_PLANE = 4, F_EDGE = 8, F_ICP = 16, F_LPS = 32, F_PRIOR = 64, F_ALL = 127 };

void build_blocks(const vils_config* cfg, const vils_window* w, const State& s, bool want_J, bool apply_loss,
                  std::vector<ResBlock>& out, int which = F_ALL) {
  const int N = s.N, D = s.D();
  out.clear();
  if (which & F_IMU)
    for (int k = 0; k < w->n_imu; k++) {
      int i = w->imu_kf[k], j = i + 1;
      ResBlock rb; rb.nr = 15; rb.nb = 4; rb.r.resize(15);
      rb.blk[0] = {15 * i, 6}; rb.blk[1] = {15 * i + 6, 9}; rb.blk[2] = {15 * j, 6}; rb.blk[3] = {15 * j + 6, 9};
      double J0[105], J1[135], J2[105], J3[135];
      imu_evaluate(&w->imu[k], cfg->gravity, &s.pose[7 * i], &s.sb[9 * i], &s.pose[7 * j], &s.sb[9 * j], rb.r.data(),
                   want_J ? J0 : nullptr, want_J ? J1 : nullptr, want_J ? J2 : nullptr, want_J ? J3 : nullptr);
      if (want_J) { take6(J0, 15, rb.J[0]); rb.J[1].assign(J1, J1 + 135); take6(J2, 15, rb.J[2]); rb.J[3].assign(J3, J3 + 135); }
      rb.rho = 0; for (double v : rb.r) rb.rho += v * v;
      out.push_back(std::move(rb));
    }
  if (which & F_PROJ)
    for (int k = 0; k < w->n_proj; k++) {
      int i = w->kf_i[k], j = w->kf_j[k], f = w->feat[k];
      ResBlock rb; rb.nr = 2; rb.nb = 5; rb.r.resize(2);
      rb.blk[0] = {15 * i, 6}; rb.blk[1] = {15 * j, 6}; rb.blk[2] = {15 * N, 6}; rb.blk[3] = {D + f, 1}; rb.blk[4] = {15 * N + 6, 1};
      double Ji[14], Jj[14], Je[14], Jf[2], Jt[2];
      proj_evaluate(cfg, cfg->estimate_td, &w->pts_i[3 * k], &w->pts_j[3 * k], &w->vel_i[2 * k], &w->vel_j[2 * k], w->td_i[k],
                    w->td_j[k], w->row_i[k], w->row_j[k], &s.pose[7 * i], &s.pose[7 * j], s.ex.data(), s.lam[f], s.td,
                    rb.r.data(), want_J ? Ji : nullptr, want_J ? Jj : nullptr, want_J ? Je : nullptr, want_J ? Jf : nullptr,
                    want_J ? Jt : nullptr);
      if (want_J) {
        // one contiguous 2x20 [pose_i pose_j ex lambda td] so the corrector sees the whole row
        double Jall[40];
        for (int r = 0; r < 2; r++) { for (int c = 0; c < 6; c++) { Jall[r * 20 + c] = Ji[r * 7 + c]; Jall[r * 20 + 6 + c] = Jj[r * 7 + c]; Jall[r * 20 + 12 + c] = Je[r * 7 + c]; } Jall[r * 20 + 18] = Jf[r]; Jall[r * 20 + 19] = Jt[r]; }
        rb.rho = corrector(apply_loss ? LOSS_CAUCHY : LOSS_NONE, cfg->cauchy_visual, 2, 20, rb.r.data(), Jall);
        for (int b = 0; b < 3; b++) { rb.J[b].resize(12); for (int r = 0; r < 2; r++) for (int c = 0; c < 6; c++) rb.J[b][r * 6 + c] = Jall[r * 20 + 6 * b + c]; }
        rb.J[3] = {Jall[18], Jall[38]}; rb.J[4] = {Jall[19], Jall[39]};
      } else rb.rho = corrector(apply_loss ? LOSS_CAUCHY : LOSS_NONE, cfg->cauchy_visual, 2, 0, rb.r.data(), nullptr);
      out.push_back(std::move(rb));
    }
  auto lidar_pt = [&](int kf, int nr, auto functor) {
    ResBlock rb; rb.nr = nr; rb.nb = 1; rb.r.resize(nr); rb.blk[0] = {15 * kf, 6};
    const double* pose = &s.pose[7 * kf];
    double Jg[1][3 * 7];
    const double* poses[1] = {pose};
    if (nr == 1) autodiff<1, 1>(functor, poses, rb.r.data(), reinterpret_cast<double(*)[7]>(Jg));
    else autodiff<3, 1>(functor, poses, rb.r.data(), Jg);
    dvec Jt(nr * 6, 0.0);
    if (want_J) {  // chain with the true derivative of Plus (EigenQuaternionParameterization plays this role in localMapping.cpp:598)
      double Jp[7][6]; plus_jacobian_true(pose, Jp);
      for (int r = 0; r < nr; r++) for (int c = 0; c < 6; c++) { double a = 0; for (int k = 0; k < 7; k++) a += Jg[0][r * 7 + k] * Jp[k][c]; Jt[r * 6 + c] = a; }
    }
    rb.rho = corrector(apply_loss ? LOSS_HUBER : LOSS_NONE, cfg->huber_lidar, nr, want_J ? 6 : 0, rb.r.data(), want_J ? Jt.data() : nullptr);
    if (want_J) rb.J[0] = Jt;
    out.push_back(std::move(rb));
  };
  if (which & F_PLANE)
    for (int k = 0; k < w->n_plane; k++) {
      const double* p = &w->plane_p[3 * k]; const double* n = &w->plane_n[3 * k]; double d = w->plane_d[k];
      lidar_pt(w->plane_kf[k], 1, [&](const Jet<7>* const* x, Jet<7>* r) { plane_functor<Jet<7>>(cfg, x[0], p, n, d, r); });
    }
  if (which & F_EDGE)
    for (int k = 0; k < w->n_edge; k++) {
      const double* p = &w->edge_p[3 * k]; const double* a = &w->edge_a[3 * k]; const double* b = &w->edge_b[3 * k];
      lidar_pt(w->edge_kf[k], 3, [&](const Jet<7>* const* x, Jet<7>* r) { edge_functor<Jet<7>>(cfg, x[0], p, a, b, r); });
    }
  if (which & F_ICP)
    for (int k = 0; k < w->n_icp; k++) {
      const vils_icp* c = &w->icp[k];
      ResBlock rb; rb.nr = 3; rb.nb = 4; rb.r.resize(3);
      const double* poses[4];
      for (int b = 0; b < 4; b++) { rb.blk[b] = {15 * c->kf[b], 6}; poses[b] = &s.pose[7 * c->kf[b]]; }
      double Jg[4][21];
      autodiff<3, 4>([&](const Jet<28>* const* x, Jet<28>* r) { icp_functor<Jet<28>>(c, x[0], x[1], x[2], x[3], r); }, poses, rb.r.data(), Jg);
      double Jall[3 * 24];
      for (int b = 0; b < 4; b++) for (int r = 0; r < 3; r++) for (int cc = 0; cc < 6; cc++) Jall[r * 24 + 6 * b + cc] = Jg[b][r * 7 + cc];  // [I6;0]
      rb.rho = corrector(apply_loss ? LOSS_CAUCHY : LOSS_NONE, cfg->cauchy_visual, 3, want_J ? 24 : 0, rb.r.data(), want_J ? Jall : nullptr);
      if (want_J) for (int b = 0; b < 4; b++) { rb.J[b].resize(18); for (int r = 0; r < 3; r++) for (int cc = 0; cc < 6; cc++) rb.J[b][r * 6 + cc] = Jall[r * 24 + 6 * b + cc]; }
      out.push_back(std::move(rb));
    }
  if (which & F_LPS)
    for (int k = 0; k < w->n_lps; k++) {
      const vils_lps* c = &w->lps[k];
      ResBlock rb; rb.nr = 3; rb.nb = 2; rb.r.resize(3);
      const double* poses[2];
      for (int b = 0; b < 2; b++) { rb.blk[b] = {15 * c->kf[b], 6}; poses[b] = &s.pose[7 * c->kf[b]]; }
      double Jg[2][21];
      autodiff<3, 2>([&](const Jet<14>* const* x, Jet<14>* r) { lps_functor<Jet<14>>(c, x[0], x[1], r); }, poses, rb.r.data(), Jg);
      double Jall[3 * 12];
      for (int b = 0; b < 2; b++) for (int r = 0; r < 3; r++) for (int cc = 0; cc < 6; cc++) Jall[r * 12 + 6 * b + cc] = Jg[b][r * 7 + cc];
      rb.rho = corrector(apply_loss ? LOSS_CAUCHY : LOSS_NONE, cfg->cauchy_visual, 3, want_J ? 12 : 0, rb.r.data(), want_J ? Jall : nullptr);
      if (want_J) for (int b = 0; b < 2; b++) { rb.J[b].resize(18); for (int r = 0; r < 3; r++) for (int cc = 0; cc < 6; cc++) rb.J[b][r * 6 + cc] = Jall[r * 12 + 6 * b + cc]; }
      out.push_back(std::move(rb));
    }
  if ((which & F_PRIOR) && w->prior_n > 0) {
    int n = w->prior_n;
    dvec dx; prior_dx(w, s, dx);
    // the prior spans up to 2N+2 blocks: emit one ResBlock per prior block PAIR is unnecessary — keep a
    // single wide block list by splitting into chunks of <= 6 blocks sharing the same residual vector.
    dvec r(n);
    for (int i = 0; i < n; i++) { double a = w->prior_r[i]; for (int j = 0; j < n; j++) a += w->prior_J[j * n + i] * dx[j]; r[i] = a; }  // :383
    // Emit as pseudo-blocks: the assembly code special-cases nb == -1 (dense prior).
    ResBlock rb; rb.nr = n; rb.nb = -1; rb.r = r; rb.rho = 0; for (double v : r) rb.rho += v * v;
    out.push_back(std::move(rb));
  }
}

// Dense normal equations over [camera (D) | landmarks (M)], row-major (D+M)^2, g = J^T r.
struct Normal { int D, M; dvec H, g; double cost; };

void add_prior_normal(const vils_window* w, int N, const dvec& r, dvec& H, dvec& g, int ld) {
  int n = w->prior_n; std::vector<int> col(n);
  int c = 0;
  for (int b = 0; b < w->prior_nblk; b++) { int id = w->prior_blk[b]; int ls = blk_local_size(VILS_BLK_TYPE(id)); int off = blk_offset(id, N); for (int i = 0; i < ls; i++) col[c++] = off + i; }
  // H += J^T J ; g += J^T r   (J column-major n x n)
  for (int a = 0; a < n; a++) {
    const double* Ja = &w->prior_J[a * n];
    double ga = 0; for (int i = 0; i < n; i++) ga += Ja[i] * r[i];
    g[col[a]] += ga;
    for (int b = 0; b < n; b++) { const double* Jb = &w->prior_J[b * n]; double h = 0; for (int i = 0; i < n; i++) h += Ja[i] * Jb[i]; H[col[a] * ld + col[b]] += h; }
  }
}

void assemble(const vils_config* cfg, const vils_window* w, const State& s, Normal& ne) {
  std::vector<ResBlock> blocks;
  build_blocks(cfg, w, s, true, true, blocks);
  int D = s.D(), M = s.M, T = D + M;
  ne.D = D; ne.M = M; ne.H.assign((size_t)T * T, 0.0); ne.g.assign(T, 0.0); ne.cost = 0;
  for (const ResBlock& rb : blocks) {
    ne.cost += 0.5 * rb.rho;
    if (rb.nb == -1) { add_prior_normal(w, s.N, rb.r, ne.H, ne.g, T); continue; }
    for (int a = 0; a < rb.nb; a++) {
      const Block& A = rb.blk[a];
      for (int i = 0; i < A.size; i++) { double ga = 0; for (int r = 0; r < rb.nr; r++) ga += rb.J[a][r * A.size + i] * rb.r[r]; ne.g[A.off + i] += ga; }
      for (int b = 0; b < rb.nb; b++) {
        const Block& B = rb.blk[b];
        for (int i = 0; i < A.size; i++) for (int j = 0; j < B.size; j++) {
          double h = 0; for (int r = 0; r < rb.nr; r++) h += rb.J[a][r * A.size + i] * rb.J[b][r * B.size + j];
          ne.H[(size_t)(A.off + i) * T + B.off + j] += h;
        }
      }
    }
  }
  // constant parameter blocks (problem.SetParameterBlockConstant): estimator.cpp:1154-1158 (ex), :1162-1166 (td absent),
  // :1217-1221 (LiDAR-depth features), :1368-1370 (zero-velocity frame).  Realised as identity rows/cols with zero gradient.
  std::vector<char> fixed(T, 0);
  if (!cfg->estimate_extrinsic) for (int i = 0; i < 6; i++) fixed[15 * s.N + i] = 1;
  if (!cfg->estimate_td) fixed[15 * s.N + 6] = 1;
  if (w->kf_fixed) for (int k = 0; k < s.N; k++) if (w->kf_fixed[k]) for (int i = 0; i < 15; i++) fixed[15 * k + i] = 1;
  for (int f = 0; f < M; f++) if ((w->depth_fixed && w->depth_fixed[f]) || ne.H[(size_t)(D + f) * T + D + f] == 0.0) fixed[D + f] = 1;
  for (int i = 0; i < T; i++) if (fixed[i]) { for (int j = 0; j < T; j++) { ne.H[(size_t)i * T + j] = 0; ne.H[(size_t)j * T + i] = 0; } ne.H[(size_t)i * T + i] = 1; ne.g[i] = 0; }
}

double total_cost(const vils_config* cfg, const vils_window* w, const State& s) {
  std::vector<ResBlock> blocks; build_blocks(cfg, w, s, false, true, blocks);
  double c = 0; for (const ResBlock& rb : blocks) c += 0.5 * rb.rho; return c;
}

// Schur-eliminate the (diagonal) landmark block with per-parameter damping d2 (added to the diagonal), then
// Cholesky-solve the reduced camera system.  Returns false on breakdown.  Outputs step (dc, dl) solving (H + diag(d2)) d = -g.
bool schur_solve(const Normal& ne, const dvec& d2, dvec& dc, dvec& dl, dvec* S_out = nullptr, dvec* gr_out = nullptr) {
  int D = ne.D, M = ne.M, T = D + M;
  dvec S((size_t)D * D), gr(D);
  for (int i = 0; i < D; i++) { gr[i] = ne.g[i]; for (int j = 0; j < D; j++) S[(size_t)i * D + j] = ne.H[(size_t)i * T + j]; S[(size_t)i * D + i] += d2[i]; }
  dvec Cinv(M);
  for (int f = 0; f < M; f++) Cinv[f] = 1.0 / (ne.H[(size_t)(D + f) * T + D + f] + d2[D + f]);
  for (int f = 0; f < M; f++) {
    double ci = Cinv[f], gl = ne.g[D + f];
    // E_f = H[0:D, D+f] is sparse: collect non-zeros
    int nz[512]; int n = 0;
    for (int i = 0; i < D; i++) if (ne.H[(size_t)i * T + D + f] != 0.0) nz[n++] = i;
    for (int a = 0; a < n; a++) {
      double ea = ne.H[(size_t)nz[a] * T + D + f] * ci;
      gr[nz[a]] -= ea * gl;
      for (int b = 0; b < n; b++) S[(size_t)nz[a] * D + nz[b]] -= ea * ne.H[(size_t)nz[b] * T + D + f];
    }
  }
  if (S_out) *S_out = S;
  if (gr_out) *gr_out = gr;
  if (!cholesky_lower(S.data(), D)) return false;
  dc.resize(D); for (int i = 0; i < D; i++) dc[i] = -gr[i];
  chol_solve(S.data(), D, dc.data());
  dl.resize(M);
  for (int f = 0; f < M; f++) { double a = ne.g[D + f]; for (int i = 0; i < D; i++) { double e = ne.H[(size_t)i * T + D + f]; if (e != 0.0) a += e * dc[i]; } dl[f] = -a * Cinv[f]; }
  for (double v : dc) if (!std::isfinite(v)) return false;
  return true;
}

// Ceres clamps the SQUARED column norms diag(J^T J) to [min_lm_diagonal, max_lm_diagonal] = [1e-6, 1e32] before taking the
// square root (levenberg_marquardt_strategy.cc / dogleg_strategy.cc [upstream]).
const double kMinDiag = 1e-6;
const double kMaxDiag = 1e32;

int solve_window(const vils_config* cfg, const vils_window* w, const vils_solve_opts* o, State& s, vils_summary* sum) {
  s.from(w);
  std::memset(sum, 0, sizeof(*sum));
  Normal ne; dvec d2, dc, dl;
  int D = s.D(), M = s.M;
  if (o->mode == VILS_MODE_GN) {
    for (int it = 0; it < o->max_iters; it++) {
      assemble(cfg, w, s, ne);
      if (it == 0) sum->cost_initial = ne.cost;
      if (!std::isfinite(ne.cost)) { sum->status = VILS_ERR_NOT_FINITE; return sum->status; }
      d2.assign(D + M, 0.0);
      for (int i = 0; i < D + M; i++) d2[i] = o->mu * std::min(std::max(ne.H[(size_t)i * (D + M) + i], kMinDiag), kMaxDiag);
      if (!schur_solve(ne, d2, dc, dl)) { sum->status = VILS_ERR_CHOLESKY; return sum->status; }
      state_plus(s, dc.data(), dl.data());
      sum->iterations++; sum->accepted++;
    }
    sum->cost_final = total_cost(cfg, w, s);
    if (!std::isfinite(sum->cost_final)) sum->status = VILS_ERR_NOT_FINITE;
    return sum->status;
  }
  if (o->mode == VILS_MODE_DOGLEG) {
    // ceres TrustRegionMinimizer + DoglegStrategy (TRADITIONAL_DOGLEG) [upstream; what estimator.cpp:1402-1411 configures], written in
    // the unscaled parameter space.  Jacobi scaling (Solver::Options::jacobi_scaling = true): column i of J is multiplied by
    // s_i = 1 / (1 + |J_i|) with the norms taken ONCE at x0; the strategy's diagonal is clamp(|J_i s_i|^2, 1e-6, 1e32), i.e. in
    // unscaled space the elliptical trust region is |diag(sqrt(dd)) step| <= radius with dd_i = clamp(H_ii s_i^2) / s_i^2.
    //   Gauss-Newton point : (H + mu diag(dd)) y = -g, mu from min_mu = 1e-8, x10 on a failed factorisation (max_mu = 1)
    //   Cauchy point       : -alpha g / dd, alpha = |g / sqrt(dd)|^2 / (v^T H v), v = g / dd
    //   step = a p + b y (p = -g / dd): the three dogleg cases;  accepted: radius = max(radius, 3 |step|) if quality > 0.75, halved if
    //   < 0.25, mu = max(1e-8, 2 mu / 10);  rejected: radius halved and the SAME y, g reused;  invalid (no descent): mu x10.
    const int T = D + M;
    double radius = o->lm_initial_radius, mu = 1e-8;
    const double min_mu = 1e-8, max_mu = 1.0, mu_inc = 10.0;
    assemble(cfg, w, s, ne); sum->iterations = 1;
    double x_cost = ne.cost; sum->cost_initial = x_cost; sum->cost_final = x_cost;
    if (!std::isfinite(x_cost)) { sum->status = VILS_ERR_NOT_FINITE; return sum->status; }
    dvec sc(T);
    for (int i = 0; i < T; i++) sc[i] = 1.0 / (1.0 + std::sqrt(ne.H[(size_t)i * T + i]));
    dvec dd(T), y(T);
    double gg = 0, pHp = 0, yy = 0, gy = 0, alpha = 0;
    bool reuse = false;
    int invalid = 0;
    for (int it = 0; it < o->max_iters; it++) {
      if (!reuse) {
        for (int i = 0; i < T; i++) dd[i] = std::min(std::max(ne.H[(size_t)i * T + i] * sc[i] * sc[i], kMinDiag), kMaxDiag) / (sc[i] * sc[i]);
        bool ok = false;
        while (mu < max_mu) {
          d2.assign(T, 0.0);
          for (int i = 0; i < T; i++) d2[i] = mu * dd[i];
          if (schur_solve(ne, d2, dc, dl)) { ok = true; break; }
          mu *= mu_inc;
        }
        if (!ok) { sum->status = VILS_ERR_CHOLESKY; return sum->status; }
        for (int i = 0; i < D; i++) y[i] = dc[i];
        for (int f = 0; f < M; f++) y[D + f] = dl[f];
        gg = 0; yy = 0; gy = 0; pHp = 0;
        dvec v(T);
        for (int i = 0; i < T; i++) { v[i] = ne.g[i] / dd[i]; gg += ne.g[i] * ne.g[i] / dd[i]; yy += dd[i] * y[i] * y[i]; gy += ne.g[i] * y[i]; }
        for (int i = 0; i < T; i++) { if (v[i] == 0.0) continue; double a = 0; for (int j = 0; j < T; j++) a += ne.H[(size_t)i * T + j] * v[j]; pHp += v[i] * a; }
        alpha = gg / pHp;
      }
      // dogleg_strategy.cc ComputeTraditionalDoglegStep
      double a, b, step_norm;
      const double gn_norm = std::sqrt(yy), g_norm = std::sqrt(gg);
      if (gn_norm <= radius) { a = 0; b = 1; step_norm = gn_norm; }
      else if (g_norm * alpha >= radius) { a = radius / g_norm; b = 0; step_norm = radius; }
      else {
        const double b_dot_a = -alpha * gy, a_sq = (alpha * g_norm) * (alpha * g_norm), bma = a_sq - 2 * b_dot_a + yy, c = b_dot_a - a_sq;
        const double dsc = std::sqrt(c * c + bma * (radius * radius - a_sq));
        const double beta = (c <= 0) ? (dsc - c) / bma : (radius * radius - a_sq) / (dsc + c);
        a = alpha * (1.0 - beta); b = beta;
        step_norm = std::sqrt(std::max(0.0, a * a * gg - 2 * a * b * gy + b * b * yy));
      }
      // model_cost_change = -g^T step - 1/2 step^T H step, with H y = -g - mu dd y
      const double g_step = -a * gg + b * gy;
      const double sHs = a * a * pHp + 2 * a * b * (gg + mu * gy) + b * b * (-gy - mu * yy);
      const double model_change = -g_step - 0.5 * sHs;
      dvec stc(D), stl(M);
      for (int i = 0; i < D; i++) stc[i] = -a * ne.g[i] / dd[i] + b * y[i];
      for (int f = 0; f < M; f++) stl[f] = -a * ne.g[D + f] / dd[D + f] + b * y[D + f];
      if (!(model_change > 0)) {   // StepIsInvalid
        mu *= mu_inc; reuse = false;
        if (++invalid >= 5 || mu >= max_mu) break;    // max_num_consecutive_invalid_steps
        continue;
      }
      invalid = 0;
      State cand = s; state_plus(cand, stc.data(), stl.data());
      const double new_cost = total_cost(cfg, w, cand);
      const double quality = std::isfinite(new_cost) ? (x_cost - new_cost) / model_change : -1;
      if (quality > o->min_relative_decrease) {
        double xn = 0, dn = 0;
        for (double v : s.pose) xn += v * v; for (double v : s.sb) xn += v * v; for (double v : s.ex) xn += v * v; for (double v : s.lam) xn += v * v; xn += s.td * s.td;
        for (double v : stc) dn += v * v; for (double v : stl) dn += v * v;
        s = cand; sum->accepted++;
        if (quality < 0.25) radius *= 0.5;
        if (quality > 0.75) radius = std::max(radius, 3.0 * step_norm);
        mu = std::max(min_mu, 2.0 * mu / mu_inc); reuse = false;
        const double change = x_cost - new_cost; x_cost = new_cost; sum->cost_final = x_cost;
        if (std::fabs(change) / (x_cost + 1e-300) < o->function_tolerance) break;
        if (std::sqrt(dn) <= o->parameter_tolerance * (std::sqrt(xn) + o->parameter_tolerance)) break;
        if (it + 1 < o->max_iters) { assemble(cfg, w, s, ne); sum->iterations++; }
      } else {
        radius *= 0.5; reuse = true;
        if (radius < 1e-32) break;
      }
    }
    return sum->status;
  }
  // Levenberg-Marquardt, following ceres TrustRegionMinimizer + LevenbergMarquardtStrategy [upstream]:
  // (H + diag(clamp(H_ii))/radius) d = -g ; rho = (cost - new_cost) / model_cost_change ;
  // accept if rho > min_relative_decrease, radius /= max(1/3, 1 - (2 rho - 1)^3) ; else radius /= decrease_factor, decrease_factor *= 2.
  double radius = o->lm_initial_radius, decrease_factor = 2.0;
  assemble(cfg, w, s, ne); sum->iterations = 1;
  double x_cost = ne.cost; sum->cost_initial = x_cost; sum->cost_final = x_cost;
  if (!std::isfinite(x_cost)) { sum->status = VILS_ERR_NOT_FINITE; return sum->status; }
  for (int it = 0; it < o->max_iters; it++) {
    d2.assign(D + M, 0.0);
    for (int i = 0; i < D + M; i++) d2[i] = std::min(std::max(ne.H[(size_t)i * (D + M) + i], kMinDiag), kMaxDiag) / radius;
    bool ok = schur_solve(ne, d2, dc, dl);
    double rho = -1;
    State cand = s; double new_cost = 0;
    if (ok) {
      double gd = 0, dd = 0;
      for (int i = 0; i < D; i++) { gd += ne.g[i] * dc[i]; dd += d2[i] * dc[i] * dc[i]; }
      for (int f = 0; f < M; f++) { gd += ne.g[D + f] * dl[f]; dd += d2[D + f] * dl[f] * dl[f]; }
      double model_change = -0.5 * gd + 0.5 * dd;  // = -g^T d - 1/2 d^T H d with (H + D2) d = -g
      state_plus(cand, dc.data(), dl.data());
      new_cost = total_cost(cfg, w, cand);
      rho = (std::isfinite(new_cost) && model_change > 0) ? (x_cost - new_cost) / model_change : -1;
    }
    if (ok && rho > o->min_relative_decrease) {
      double xn = 0, dn = 0;
      for (double v : s.pose) xn += v * v; for (double v : s.sb) xn += v * v; for (double v : s.ex) xn += v * v; for (double v : s.lam) xn += v * v; xn += s.td * s.td;
      for (double v : dc) dn += v * v; for (double v : dl) dn += v * v;
      s = cand; sum->accepted++;
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3)); radius = std::min(radius, 1e16); decrease_factor = 2.0;
      double change = x_cost - new_cost; x_cost = new_cost; sum->cost_final = x_cost;
      if (std::fabs(change) / (x_cost + 1e-300) < o->function_tolerance) break;  // FUNCTION_TOLERANCE (ceres compares against the pre-step cost; same order)
      if (std::sqrt(dn) <= o->parameter_tolerance * (std::sqrt(xn) + o->parameter_tolerance)) break;
      if (it + 1 < o->max_iters) { assemble(cfg, w, s, ne); sum->iterations++; }
    } else {
      radius /= decrease_factor; decrease_factor *= 2.0;
      if (radius < 1e-32) break;
    }
  }
  return sum->status;
}

// ------------------------------------------------------------------------------------------
// Marginalization: Estimator::optimization() tail (estimator.cpp:1483-1684) +
// MarginalizationInfo::{addResidualBlockInfo,preMarginalize,marginalize,getParameterBlocks}
// (factor/marginalization_factor.cpp:89-129,176-338).
// Canonical block order (the reference's unordered_map<long,...> order is not reproducible):
//   dropped: pose0, sb0 (MARGIN_OLD) / pose[N-2] (SECOND_NEW), then dropped landmarks by feature index;
//   kept: poses ascending, then speed-bias ascending, ex, td — only those touched by a participating factor.
// ------------------------------------------------------------------------------------------
int marginalize_window(const vils_config* cfg, const vils_window* w, int flag, vils_prior_out* out) {
  State s; s.from(w);
  const int N = s.N, D = s.D(), M = s.M, T = D + M;
  // --- select factors ---
  vils_window sub = *w;
  std::vector<int> proj_sel;
  std::vector<char> prior_has(T, 0);
  dvec H((size_t)T * T, 0.0), g(T, 0.0);
  std::vector<char> touched(T, 0);
  std::vector<ResBlock> blocks;
  auto accumulate = [&](const ResBlock& rb) {
    for (int a = 0; a < rb.nb; a++) {
      const Block& A = rb.blk[a];
      for (int i = 0; i < A.size; i++) { touched[A.off + i] = 1; double ga = 0; for (int r = 0; r < rb.nr; r++) ga += rb.J[a][r * A.size + i] * rb.r[r]; g[A.off + i] += ga; }
      for (int b = 0; b < rb.nb; b++) { const Block& B = rb.blk[b];
        for (int i = 0; i < A.size; i++) for (int j = 0; j < B.size; j++) { double h = 0; for (int r = 0; r < rb.nr; r++) h += rb.J[a][r * A.size + i] * rb.J[b][r * B.size + j]; H[(size_t)(A.off + i) * T + B.off + j] += h; } }
    }
  };
  auto add_prior = [&]() {
    if (w->prior_n <= 0) return;
    build_blocks(cfg, w, s, true, true, blocks, F_PRIOR);
    add_prior_normal(w, N, blocks[0].r, H, g, T);
    for (int b = 0; b < w->prior_nblk; b++) { int id = w->prior_blk[b]; int ls = blk_local_size(VILS_BLK_TYPE(id)), off = blk_offset(id, N); for (int i = 0; i < ls; i++) touched[off + i] = 1; }
  };
  std::vector<char> drop(T, 0);
  if (flag == VILS_MARGIN_OLD) {
    add_prior();  // :1489-1505, drop pose0 / sb0 if present
    // ICP / LPS touching frame 0: the LAST such constraint only (ICPmarg / LPSmarg are overwritten in the loops :1381-1389, :1311-1317)
    int icp_sel = -1, lps_sel = -1;
    for (int k = 0; k < w->n_icp; k++) if (w->icp[k].kf[0] == 0) icp_sel = k;
    for (int k = 0; k < w->n_lps; k++) if (w->lps[k].kf[0] == 0) lps_sel = k;
    build_blocks(cfg, w, s, true, true, blocks, F_ICP);
    if (icp_sel >= 0) accumulate(blocks[icp_sel]);  // :1508-1518
    build_blocks(cfg, w, s, true, true, blocks, F_LPS);
    if (lps_sel >= 0) accumulate(blocks[lps_sel]);  // :1522-1532
    build_blocks(cfg, w, s, true, true, blocks, F_IMU);
    for (int k = 0; k < w->n_imu; k++) if (w->imu_kf[k] == 0 && w->imu[k].sum_dt < 10.0) accumulate(blocks[k]);  // :1536-1543
    build_blocks(cfg, w, s, true, true, blocks, F_PROJ);
    for (int k = 0; k < w->n_proj; k++) if (w->kf_i[k] == 0) { accumulate(blocks[k]); drop[D + w->feat[k]] = 1; }  // :1547-1588 drop {pose_i, lambda}
    for (int i = 0; i < 15; i++) drop[i] = 1;  // pose0 + sb0
  } else {
    // MARGIN_SECOND_NEW (:1618-1683): only the prior, dropping Pose[N-2]; skipped unless the prior contains it (:1620-1621)
    bool has = false;
    for (int b = 0; b < w->prior_nblk; b++) if (w->prior_blk[b] == VILS_BLK_ID(VILS_BLK_POSE, N - 2)) has = true;
    if (!has) { out->n = 0; out->nblk = 0; out->m = 0; return VILS_OK; }
    add_prior();
    for (int i = 0; i < 6; i++) drop[15 * (N - 2) + i] = 1;
  }
  // td not in the problem when !estimate_td (projection factor has no td block)
  if (!cfg->estimate_td) touched[15 * N + 6] = 0;
  // --- ordering [drop ; keep] (:178-196) ---
  std::vector<int> order; int m = 0;
  for (int i = 0; i < T; i++) if (touched[i] && drop[i]) order.push_back(i);
  m = (int)order.size();
  std::vector<int> keep_ids;
  auto keep_block = [&](int id) { int ls = blk_local_size(VILS_BLK_TYPE(id)), off = blk_offset(id, N); if (!touched[off] || drop[off]) return; keep_ids.push_back(id); for (int i = 0; i < ls; i++) order.push_back(off + i); };
  for (int k = 0; k < N; k++) keep_block(VILS_BLK_ID(VILS_BLK_POSE, k));
  for (int k = 0; k < N; k++) keep_block(VILS_BLK_ID(VILS_BLK_SPEEDBIAS, k));
  keep_block(VILS_BLK_ID(VILS_BLK_EXPOSE, 0));
  keep_block(VILS_BLK_ID(VILS_BLK_TD, 0));
  int pos = (int)order.size(), n = pos - m;
  if (n > out->capacity_n) return VILS_ERR_CAPACITY;
  dvec A((size_t)pos * pos), b(pos);
  for (int i = 0; i < pos; i++) { b[i] = g[order[i]]; for (int j = 0; j < pos; j++) A[(size_t)i * pos + j] = H[(size_t)order[i] * T + order[j]]; }
  // :274-290
  dvec Amm((size_t)m * m), wv(m), V((size_t)m * m), Amm_inv((size_t)m * m, 0.0);
  for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) Amm[(size_t)i * m + j] = 0.5 * (A[(size_t)i * pos + j] + A[(size_t)j * pos + i]);
  if (m > 0) eigh_jacobi(Amm.data(), m, wv.data(), V.data());
  const double eps = 1e-8;  // marginalization_factor.h:70
  for (int k = 0; k < m; k++) { if (!(wv[k] > eps)) continue; double iv = 1.0 / wv[k]; for (int i = 0; i < m; i++) { double vi = V[(size_t)i * m + k] * iv; for (int j = 0; j < m; j++) Amm_inv[(size_t)i * m + j] += vi * V[(size_t)j * m + k]; } }
  dvec Arm_Ainv((size_t)n * m, 0.0);
  for (int i = 0; i < n; i++) for (int j = 0; j < m; j++) { double a = 0; for (int k = 0; k < m; k++) a += A[(size_t)(m + i) * pos + k] * Amm_inv[(size_t)k * m + j]; Arm_Ainv[(size_t)i * m + j] = a; }
  dvec Ar((size_t)n * n), br(n);
  for (int i = 0; i < n; i++) {
    double bb = b[m + i]; for (int k = 0; k < m; k++) bb -= Arm_Ainv[(size_t)i * m + k] * b[k]; br[i] = bb;
    for (int j = 0; j < n; j++) { double a = A[(size_t)(m + i) * pos + m + j]; for (int k = 0; k < m; k++) a -= Arm_Ainv[(size_t)i * m + k] * A[(size_t)k * pos + m + j]; Ar[(size_t)i * n + j] = a; }
  }
  // :301-309 (SelfAdjointEigenSolver reads the lower triangle; symmetrise to be order-independent)
  for (int i = 0; i < n; i++) for (int j = 0; j < i; j++) { double a = 0.5 * (Ar[(size_t)i * n + j] + Ar[(size_t)j * n + i]); Ar[(size_t)i * n + j] = Ar[(size_t)j * n + i] = a; }
  dvec S(n), V2((size_t)n * n);
  eigh_jacobi(Ar.data(), n, S.data(), V2.data());
  for (int k = 0; k < n; k++) {
    double sv = S[k] > eps ? S[k] : 0.0, sinv = S[k] > eps ? 1.0 / S[k] : 0.0;
    double ssq = std::sqrt(sv), sisq = std::sqrt(sinv);
    double rb = 0;
    for (int j = 0; j < n; j++) { out->J[(size_t)j * n + k] = ssq * V2[(size_t)j * n + k]; rb += V2[(size_t)j * n + k] * br[j]; }  // J_lin(k,j) col-major
    out->r[k] = sisq * rb;
  }
  // getParameterBlocks + addr_shift (:318-338, estimator.cpp:1599-1611 / :1654-1675)
  out->n = n; out->m = m; out->nblk = (int)keep_ids.size();
  double* x0 = out->x0;
  for (size_t bidx = 0; bidx < keep_ids.size(); bidx++) {
    int id = keep_ids[bidx], type = VILS_BLK_TYPE(id), idx = VILS_BLK_INDEX(id);
    const double* x = type == VILS_BLK_POSE ? &s.pose[7 * idx] : type == VILS_BLK_SPEEDBIAS ? &s.sb[9 * idx] : type == VILS_BLK_EXPOSE ? s.ex.data() : &s.td;
    int gs = blk_global_size(type); for (int i = 0; i < gs; i++) *x0++ = x[i];
    int nidx = idx;
    if (type == VILS_BLK_POSE || type == VILS_BLK_SPEEDBIAS) {
      if (flag == VILS_MARGIN_OLD) nidx = idx - 1;
      else nidx = (idx == N - 1) ? N - 2 : idx;
    }
    out->blk[bidx] = VILS_BLK_ID(type, nidx);
  }
  return VILS_OK;
}

}  // namespace

// ============================================================================================
// extern "C" surface (ctypes)
// ============================================================================================
extern "C" {

int vo_preintegrate(int n_intervals, const int32_t* off, const double* dt, const double* acc, const double* gyr,
                    const double* acc0, const double* gyr0, const double* ba, const double* bg, const double noise[4],
                    vils_preint* out) {
  for (int k = 0; k < n_intervals; k++) {
    Preint s; preint_init(s, v3(acc0 + 3 * k), v3(gyr0 + 3 * k), v3(ba + 3 * k), v3(bg + 3 * k), noise);
    for (int i = off[k]; i < off[k + 1]; i++) preint_propagate(s, dt[i], v3(acc + 3 * i), v3(gyr + 3 * i));
    preint_export(s, &out[k]);
  }
  return 0;
}

void vo_imu_sqrt_info(const vils_preint* pre, double* W_rowmajor) { double W[15][15]; imu_sqrt_info(pre, W); std::memcpy(W_rowmajor, W, sizeof(W)); }

void vo_imu_evaluate(const vils_preint* pre, const double* G3, const double* pi, const double* sbi, const double* pj,
                     const double* sbj, double* residuals, double* Jpi, double* Jsbi, double* Jpj, double* Jsbj) {
  imu_evaluate(pre, G3, pi, sbi, pj, sbj, residuals, Jpi, Jsbi, Jpj, Jsbj);
}

void vo_proj_evaluate(const vils_config* cfg, int use_td, const double* pts_i, const double* pts_j, const double* vel_i,
                      const double* vel_j, double td_i, double td_j, double row_i, double row_j, const double* pi,
                      const double* pj, const double* pex, double inv_dep, double td, double* residual, double* Ji,
                      double* Jj, double* Jex, double* Jf, double* Jtd) {
  proj_evaluate(cfg, use_td, pts_i, pts_j, vel_i, vel_j, td_i, td_j, row_i, row_j, pi, pj, pex, inv_dep, td, residual, Ji, Jj, Jex, Jf, Jtd);
}

// Residual-only functors for finite-difference tests (global parameters, plain doubles).
void vo_plane_residual(const vils_config* cfg, const double* pose, const double* p, const double* n, double d, double* r) { plane_functor<double>(cfg, pose, p, n, d, r); }
void vo_edge_residual(const vils_config* cfg, const double* pose, const double* p, const double* a, const double* b, double* r) { edge_functor<double>(cfg, pose, p, a, b, r); }
void vo_icp_residual(const vils_icp* c, const double* pa, const double* pb, const double* pc, const double* pd, double* r) { icp_functor<double>(c, pa, pb, pc, pd, r); }
void vo_lps_residual(const vils_lps* c, const double* pa, const double* pb, double* r) { lps_functor<double>(c, pa, pb, r); }
void vo_pose_plus(double* pose7, const double* delta6) { pose_plus(pose7, delta6); }

int vo_residual_count(const vils_window* w) { return 15 * w->n_imu + 2 * w->n_proj + w->n_plane + 3 * w->n_edge + 3 * w->n_icp + 3 * w->n_lps + w->prior_n; }
int vo_jacobian_count(const vils_window* w) { return 450 * w->n_imu + 40 * w->n_proj + 6 * w->n_plane + 18 * w->n_edge + 72 * w->n_icp + 36 * w->n_lps; }

// Same layout as vils_ba_evaluate (include/vils_cabi.h).
int vo_evaluate_window(const vils_config* cfg, const vils_window* w, int apply_loss, double* residuals, double* jacobians, double* cost) {
  State s; s.from(w);
  std::vector<ResBlock> blocks; build_blocks(cfg, w, s, true, apply_loss != 0, blocks);
  double c = 0; double* r = residuals; double* J = jacobians;
  for (const ResBlock& rb : blocks) {
    c += 0.5 * rb.rho;
    if (r) { std::memcpy(r, rb.r.data(), sizeof(double) * rb.nr); r += rb.nr; }
    if (rb.nb == -1 || !J) continue;
    int width = 0; for (int b = 0; b < rb.nb; b++) width += rb.blk[b].size;
    for (int i = 0; i < rb.nr; i++) { int c0 = 0; for (int b = 0; b < rb.nb; b++) { int sz = rb.blk[b].size; for (int j = 0; j < sz; j++) J[i * width + c0 + j] = rb.J[b][i * sz + j]; c0 += sz; } }
    J += rb.nr * width;
  }
  if (cost) *cost = c;
  return 0;
}

// Reduced camera system without damping: S (D x D row-major), g_r (D), cost.
int vo_linearize_window(const vils_config* cfg, const vils_window* w, double* S, double* g, double* cost) {
  State s; s.from(w); Normal ne; assemble(cfg, w, s, ne);
  dvec d2(ne.D + ne.M, 0.0), dc, dl, So, go;
  schur_solve(ne, d2, dc, dl, &So, &go);
  std::memcpy(S, So.data(), sizeof(double) * So.size()); std::memcpy(g, go.data(), sizeof(double) * go.size());
  if (cost) *cost = ne.cost;
  return 0;
}

int vo_solve_window(const vils_config* cfg, const vils_window* w, const vils_solve_opts* opts, double* pose, double* sb,
                    double* ex, double* lam, double* td, vils_summary* sum) {
  State s; int st = solve_window(cfg, w, opts, s, sum);
  std::memcpy(pose, s.pose.data(), sizeof(double) * 7 * s.N); std::memcpy(sb, s.sb.data(), sizeof(double) * 9 * s.N);
  std::memcpy(ex, s.ex.data(), sizeof(double) * 7); if (s.M) std::memcpy(lam, s.lam.data(), sizeof(double) * s.M); *td = s.td;
  return st;
}

int vo_marginalize_window(const vils_config* cfg, const vils_window* w, int flag, vils_prior_out* out) { return marginalize_window(cfg, w, flag, out); }

// Estimator::double2vector gauge re-anchoring (estimator.cpp:962-1011). pose0_before = para_Pose[0] snapshot before the solve.
int vo_double2vector(int n_kf, const double* pose0_before, double* pose, double* sb) {
  Mat3 R0 = toR(pose_q(pose0_before)); Vec3 origin_R0 = R2ypr(R0); Vec3 origin_P0 = v3(pose0_before);
  Mat3 R00m = toR(pose_q(pose)); Vec3 origin_R00 = R2ypr(R00m);
  double y_diff = origin_R0.x - origin_R00.x;
  Mat3 rot_diff = ypr2R(Vec3(y_diff, 0, 0));
  if (std::fabs(std::fabs(origin_R0.y) - 90) < 1.0 || std::fabs(std::fabs(origin_R00.y) - 90) < 1.0) rot_diff = R0 * transpose(R00m);
  Vec3 P0 = v3(pose);
  for (int i = 0; i < n_kf; i++) {
    double* p = pose + 7 * i;
    Mat3 R = rot_diff * toR(normalized(pose_q(p)));
    Vec3 P = rot_diff * (v3(p) - P0) + origin_P0;
    Vec3 V = rot_diff * v3(sb + 9 * i);
    Quat q = fromR(R);  // vector2double(): Quaterniond q{Rs[i]} (estimator.cpp:923)
    p[0] = P.x; p[1] = P.y; p[2] = P.z; p[3] = q.x; p[4] = q.y; p[5] = q.z; p[6] = q.w;
    sb[9 * i] = V.x; sb[9 * i + 1] = V.y; sb[9 * i + 2] = V.z;
  }
  return 0;
}

// ---- LiDAR ---------------------------------------------------------------------------------
// TransformToEnd (vils_estimator/src/lidar_frontend.cpp:1001-1041), FP32, in place.
int vo_deskew(float* xyzi, int n, int stride, const float q[4], const float t[3], float time_factor, double min_r, double max_r) {
  typedef QuatT<float> Qf; typedef Vec3T<float> Vf;
  Qf q_e(q[3], q[0], q[1], q[2]); Vf te(t[0], t[1], t[2]);
  for (int i = 0; i < n; i++) {
    float* p = xyzi + (size_t)i * stride;
    float& intensity = p[stride >= 8 ? 4 : 3];
    double distance = std::sqrt(p[0] * p[0] + p[1] * p[1]);            // :1008 (float product, double sqrt)
    float s = time_factor * (intensity - int(intensity));                // :1009
    if (s < 0 || s > 1.001 || distance < min_r || distance > max_r) {    // :1011
      p[0] = p[1] = p[2] = std::numeric_limits<float>::quiet_NaN(); continue; }
    p[0] -= s * te.x; p[1] -= s * te.y; p[2] -= s * te.z;               // :1021-1023
    Qf q_s = slerp(Qf(), s, q_e);                                        // :1026-1029
    Vf v(p[0], p[1], p[2]);
    v = rotate(normalized(conj(q_s)), v);                                // :1031
    v = rotate(q_e, v);                                                  // :1034
    p[0] = v.x + te.x; p[1] = v.y + te.y; p[2] = v.z + te.z;             // :1035-1037
    intensity = int(intensity);                                          // :1038
  }
  return 0;
}

// PointProcessor::PointToRing (lidar_compensator/src/PointProcessor.cc:127-341), non-DEBUG_ORIGIN branch,
// infer_start_ori_ = false (PointProcessor.h:47). Instead of re-ordering into per-ring clouds the ring id is
// returned per point (ring_out = -1 for dropped points); intensity <- int(I) + rel_time (:331).
int vo_stamp_rings(float* xyzi, int n, int stride, float lower_deg, float upper_deg, int n_rings, float scan_period, int32_t* ring_out) {
  float factor = (n_rings - 1) / (upper_deg - lower_deg);  // PointProcessor.cc:15
  bool start_flag = false; float start_ori = 0;
  std::vector<float> azi(n);
  for (int i = 0; i < n; i++) {
    float* p = xyzi + (size_t)i * stride; ring_out[i] = -1;
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;   // :161-166
    float dis = std::sqrt(p[0] * p[0] + p[1] * p[1]);                                     // :168
    float ele_rad = std::atan2(p[2], dis);                                                // :169
    float azi_rad = 2 * M_PI - std::atan2(p[1], p[0]);                                    // :170 (double expr -> float)
    if (azi_rad >= 2 * M_PI) azi_rad -= 2 * M_PI;                                         // :174-177
    float deg = ele_rad * 180.0 / M_PI;                                                   // RadToDeg (math_utils.h)
    int scan_id = int((deg - lower_deg) * factor + 0.5);                                  // PointProcessor.h:77-81
    if (scan_id >= n_rings || scan_id < 0) continue;                                      // :181-184
    if (!start_flag) { start_ori = azi_rad; start_flag = true; }                          // :186-190
    azi[i] = azi_rad; ring_out[i] = scan_id;
  }
  for (int i = 0; i < n; i++) {
    if (ring_out[i] < 0) continue;
    float* p = xyzi + (size_t)i * stride; float& intensity = p[stride >= 8 ? 4 : 3];
    float azi_rad_rel = azi[i] - start_ori;                                               // :318
    if (azi_rad_rel < 0) azi_rad_rel += 2 * M_PI;                                         // :319-322
    float rel_time = scan_period * azi_rad_rel / (2 * M_PI);                              // :324
    intensity = int(intensity) + rel_time;                                                // :331
  }
  return 0;
}

// Sweep start->end transform in the LiDAR frame (estimator.cpp:190-232). Rwb*/Pwb* row-major/3-vectors of the bracketing keyframes.
int vo_sweep_transform(const double* Rwbi, const double* Pwbi, const double* Rwbj, const double* Pwbj, double ta, double tb, double tl,
                       double lidar_time_step, const double* rlb9, const double* tlb3, float q_out[4], float t_out[3]) {
  Mat3 Ri, Rj, RLB; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { Ri(i, j) = Rwbi[i * 3 + j]; Rj(i, j) = Rwbj[i * 3 + j]; RLB(i, j) = rlb9[i * 3 + j]; }
  Vec3 TLB = v3(tlb3);
  double tls = tl - 0.5 * lidar_time_step, tle = tl + 0.5 * lidar_time_step;
  Quat temQa = fromR(Ri), temQb = fromR(Rj);
  double ss = (tls - ta) / (tb - ta), se = (tle - ta) / (tb - ta);
  Quat temQls = slerp(temQa, ss, temQb), temQle = slerp(temQa, se, temQb);
  Mat3 Rb = toR(inverse(temQle) * temQls);
  Vec3 tb_ = (-(transpose(Rj) * v3(Pwbj)) + transpose(Rj) * v3(Pwbi)) * (lidar_time_step / (tb - ta));
  // trans_l = T_lb * T_b * T_lb^-1
  Mat3 Rl = RLB * Rb * transpose(RLB);
  Vec3 tl_ = RLB * tb_ + TLB - Rl * TLB;
  Mat3 Rlf; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Rlf(i, j) = (double)(float)Rl(i, j);
  Quat q = fromR(Rlf);
  q_out[0] = (float)q.x; q_out[1] = (float)q.y; q_out[2] = (float)q.z; q_out[3] = (float)q.w;
  t_out[0] = (float)tl_.x; t_out[1] = (float)tl_.y; t_out[2] = (float)tl_.z;
  return 0;
}

}  // extern "C"
