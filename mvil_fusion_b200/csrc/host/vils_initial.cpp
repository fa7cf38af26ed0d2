// vils_initial.cpp — the one-shot bootstrap of the estimator: Estimator::initialStructure / relativePose / visualInitialAlign
// (vils_estimator/src/estimator.cpp:618-901) with GlobalSFM (initial/initial_sfm.cpp:5-309), MotionEstimator::solveRelativeRT + the
// vendored cv::recoverPose (initial/solve_5pts.cpp:3-230) and this fork's VisualIMUAlignment = Estimate_ric_td_bg + Estimate_vel_g_s_tic
// (initial/initial_aligment.cpp:193-519, functors initial/initial_alignment.h:38-193).
//
// Host code, as SURVEY.md 8f ranks it ("one-shot; port last, on host").  What runs on the device: the RANSAC fundamental matrix
// (vils_reject_with_f), the pre-integration of every image interval (vils_preintegrate) and the triangulation that follows the alignment.
// Third-party pieces restated from their published behaviour (none of them is in the reference tree, parity unpinned):
//   ceres::Solve on the three small problems -> a dense Levenberg-Marquardt (ceres' default trust-region strategy; the reference asks DOGLEG
//     for two of them) with central-difference Jacobians in the tangent space and box constraints by projection;
//   cv::solvePnPRansac(..., useExtrinsicGuess) with an 8-unit reprojection threshold on normalised coordinates (every point is an inlier)
//     -> iterative refinement of the guess over all points (SOLVEPNP_ITERATIVE);
//   cv::triangulatePoints / JacobiSVD -> the null vector of the 4 x 4 design matrix by a cyclic Jacobi eigen-solver on A^T A.
// Reference behaviour kept: Bgs / Bas are REPLACED by the estimated values (:338,:497); per-frame scales s_i.  Two reference defects are NOT
// reproduced (both documented in DESIGN.md): (1) visualInitialAlign reads the time offsets through an uninitialised pointer
// (estimator.cpp:773-809) — here td becomes the mean of the offsets Estimate_ric_td_bg estimated; (2) the loop that moves the frame
// rotations from the camera to the body frame stops one frame early (initial_aligment.cpp:331-347), which leaves the NEWEST frame rotated by
// RIC (180 degrees for the configured extrinsic) — here every frame is converted.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <vector>

#include "vils_host.h"
#include "vils_hostmath.h"

namespace vils {
using namespace hm;
namespace {

typedef std::vector<double> Vec;

// ---- dense linear algebra -----------------------------------------------------------------------------------------------------------------
bool chol_solve(Vec& A, Vec& b, int n) {   // A (row-major, symmetric positive definite) x = b, in place
  for (int j = 0; j < n; j++) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; k++) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0.0)) return false;
    d = std::sqrt(d); A[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = A[(size_t)i * n + j];
      for (int k = 0; k < j; k++) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
      A[(size_t)i * n + j] = s / d;
    }
  }
  for (int i = 0; i < n; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= A[(size_t)i * n + k] * b[k]; b[i] = s / A[(size_t)i * n + i]; }
  for (int i = n - 1; i >= 0; i--) { double s = b[i]; for (int k = i + 1; k < n; k++) s -= A[(size_t)k * n + i] * b[k]; b[i] = s / A[(size_t)i * n + i]; }
  return true;
}

// cyclic Jacobi eigen-decomposition of a small symmetric matrix: A = V diag(w) V^T, eigenvalues ascending
void jacobi_eig(const double* A_in, int n, double* w, double* V) {
  Vec A(A_in, A_in + n * n);
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) V[i * n + j] = i == j;
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0; for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) off += A[i * n + j] * A[i * n + j];
    if (off < 1e-300) break;
    for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) {
      if (std::fabs(A[p * n + q]) < 1e-300) continue;
      const double th = (A[q * n + q] - A[p * n + p]) / (2 * A[p * n + q]);
      const double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1)), c = 1 / std::sqrt(t * t + 1), s = t * c;
      for (int k = 0; k < n; k++) { const double akp = A[k * n + p], akq = A[k * n + q]; A[k * n + p] = c * akp - s * akq; A[k * n + q] = s * akp + c * akq; }
      for (int k = 0; k < n; k++) { const double apk = A[p * n + k], aqk = A[q * n + k]; A[p * n + k] = c * apk - s * aqk; A[q * n + k] = s * apk + c * aqk; }
      for (int k = 0; k < n; k++) { const double vkp = V[k * n + p], vkq = V[k * n + q]; V[k * n + p] = c * vkp - s * vkq; V[k * n + q] = s * vkp + c * vkq; }
    }
  }
  std::vector<int> idx(n); for (int i = 0; i < n; i++) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&](int a, int b) { return A[a * n + a] < A[b * n + b]; });
  Vec Vs(n * n);
  for (int k = 0; k < n; k++) { w[k] = A[idx[k] * n + idx[k]]; for (int i = 0; i < n; i++) Vs[i * n + k] = V[i * n + idx[k]]; }
  std::memcpy(V, Vs.data(), sizeof(double) * n * n);
}

double det3(const double* M) { return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]); }

// SVD of a 3 x 3 matrix, singular values descending: E = U diag(S) Vt
void svd3(const double* E, double* U, double* S, double* Vt) {
  double EtE[9], w[3], V[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double a = 0; for (int k = 0; k < 3; k++) a += E[3 * k + i] * E[3 * k + j]; EtE[3 * i + j] = a; }
  jacobi_eig(EtE, 3, w, V);
  double Vd[9];                                              // columns in descending order of the eigenvalue
  for (int k = 0; k < 3; k++) { S[k] = std::sqrt(std::max(w[2 - k], 0.0)); for (int i = 0; i < 3; i++) Vd[3 * i + k] = V[3 * i + 2 - k]; }
  double u[3][3];
  for (int k = 0; k < 2; k++) {
    double n2 = 0;
    for (int i = 0; i < 3; i++) { u[k][i] = E[3 * i] * Vd[k] + E[3 * i + 1] * Vd[3 + k] + E[3 * i + 2] * Vd[6 + k]; n2 += u[k][i] * u[k][i]; }
    n2 = std::sqrt(n2); for (int i = 0; i < 3; i++) u[k][i] /= (n2 > 0 ? n2 : 1.0);
  }
  // re-orthogonalise the second column against the first, the third is their cross product (rank-2 essential matrices)
  double d = u[0][0] * u[1][0] + u[0][1] * u[1][1] + u[0][2] * u[1][2];
  double n2 = 0; for (int i = 0; i < 3; i++) { u[1][i] -= d * u[0][i]; n2 += u[1][i] * u[1][i]; }
  n2 = std::sqrt(n2); for (int i = 0; i < 3; i++) u[1][i] /= (n2 > 0 ? n2 : 1.0);
  u[2][0] = u[0][1] * u[1][2] - u[0][2] * u[1][1]; u[2][1] = u[0][2] * u[1][0] - u[0][0] * u[1][2]; u[2][2] = u[0][0] * u[1][1] - u[0][1] * u[1][0];
  for (int k = 0; k < 3; k++) for (int i = 0; i < 3; i++) { U[3 * i + k] = u[k][i]; Vt[3 * k + i] = Vd[3 * i + k]; }
}

// ---- Levenberg-Marquardt (ceres TrustRegionMinimizer + LevenbergMarquardtStrategy defaults) on a tangent space ----------------------------
struct Lsq {
  int m = 0;                                                               // tangent dimension
  std::function<void(const Vec& x, Vec& r)> residual;
  std::function<void(const Vec& x, const double* d, Vec& xn)> plus;        // x [+] d, including the projection onto box constraints
  int max_iter = 50;
};
struct LsqSummary { double cost = 0; int iterations = 0; bool converged = false; };

LsqSummary lm_solve(const Lsq& P, Vec& x) {
  LsqSummary sum;
  Vec r, rp, rm, xn, J, H, g(P.m), d(P.m), dd(P.m);
  P.residual(x, r);
  const int nr = (int)r.size();
  double cost = 0; for (double v : r) cost += 0.5 * v * v;
  double radius = 1e4, decrease = 2.0;
  J.resize((size_t)nr * P.m); H.resize((size_t)P.m * P.m);
  const double h = 1e-6;
  for (int it = 0; it < P.max_iter; it++) {
    sum.iterations = it + 1;
    for (int k = 0; k < P.m; k++) {                                        // central differences in the tangent space
      std::fill(d.begin(), d.end(), 0.0);
      d[k] = h; P.plus(x, d.data(), xn); P.residual(xn, rp);
      d[k] = -h; P.plus(x, d.data(), xn); P.residual(xn, rm);
      for (int i = 0; i < nr; i++) J[(size_t)i * P.m + k] = (rp[i] - rm[i]) / (2 * h);
    }
    double gmax = 0;
    for (int a = 0; a < P.m; a++) {
      double ga = 0; for (int i = 0; i < nr; i++) ga += J[(size_t)i * P.m + a] * r[i];
      g[a] = ga; gmax = std::max(gmax, std::fabs(ga));
      for (int b = a; b < P.m; b++) { double s = 0; for (int i = 0; i < nr; i++) s += J[(size_t)i * P.m + a] * J[(size_t)i * P.m + b]; H[(size_t)a * P.m + b] = H[(size_t)b * P.m + a] = s; }
    }
    if (gmax < 1e-10) { sum.converged = true; break; }
    bool accepted = false;
    for (int trial = 0; trial < 20 && !accepted; trial++) {
      Vec A = H, b(P.m);
      for (int a = 0; a < P.m; a++) { dd[a] = std::min(std::max(H[(size_t)a * P.m + a], 1e-6), 1e32) / radius; A[(size_t)a * P.m + a] += dd[a]; b[a] = -g[a]; }
      if (!chol_solve(A, b, P.m)) { radius /= decrease; decrease *= 2; continue; }
      double gd = 0, dDd = 0, dn = 0, xnorm = 0;
      for (int a = 0; a < P.m; a++) { gd += g[a] * b[a]; dDd += dd[a] * b[a] * b[a]; dn += b[a] * b[a]; }
      for (double v : x) xnorm += v * v;
      const double model = -0.5 * gd + 0.5 * dDd;
      P.plus(x, b.data(), xn); P.residual(xn, rp);
      double nc = 0; for (double v : rp) nc += 0.5 * v * v;
      const double rho = (std::isfinite(nc) && model > 0) ? (cost - nc) / model : -1;
      if (rho > 1e-3) {
        const double change = cost - nc;
        x = xn; r = rp; cost = nc; accepted = true;
        const double t3 = 2 * rho - 1; radius = std::min(radius / std::max(1.0 / 3.0, 1.0 - t3 * t3 * t3), 1e16); decrease = 2.0;
        if (std::fabs(change) / (cost + 1e-300) < 1e-6 || std::sqrt(dn) <= 1e-8 * (std::sqrt(xnorm) + 1e-8)) sum.converged = true;
      } else {
        radius /= decrease; decrease *= 2;
        if (std::sqrt(dn) <= 1e-8 * (std::sqrt(xnorm) + 1e-8) || radius < 1e-32) { sum.converged = true; accepted = true; }
      }
    }
    if (!accepted || sum.converged) { sum.converged = sum.converged || !accepted; break; }
  }
  sum.cost = cost;
  return sum;
}

// q (x y z w) <- q (*) exp(d): right perturbation with the exact exponential
void q_plus(const double* q, const double* d, double* o) {
  const double th = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  double dq[4];
  if (th < 1e-12) { dq[0] = 0.5 * d[0]; dq[1] = 0.5 * d[1]; dq[2] = 0.5 * d[2]; dq[3] = 1.0; }
  else { const double s = std::sin(0.5 * th) / th; dq[0] = s * d[0]; dq[1] = s * d[1]; dq[2] = s * d[2]; dq[3] = std::cos(0.5 * th); }
  qmul(q, dq, o); qnorm(o);
}
inline double clampd(double v, double lo, double hi) { return std::min(std::max(v, lo), hi); }

// ---- GlobalSFM (initial/initial_sfm.cpp) ----------------------------------------------------------------------------------------------------
struct SFMFeature { bool state = false; int id = 0; std::vector<std::pair<int, std::array<double, 2>>> observation; double position[3] = {0, 0, 0}; };
struct Pose34 { double m[12]; };   // row-major 3 x 4 [R | t], world -> camera

void make_pose(const double R[9], const double t[3], Pose34& P) { for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) P.m[4 * i + j] = R[3 * i + j]; P.m[4 * i + 3] = t[i]; } }

// initial_sfm.cpp:5-21 (cv::triangulatePoints uses the same design matrix)
void triangulatePoint(const Pose34& P0, const Pose34& P1, const double p0[2], const double p1[2], double X[3], double* Xh = nullptr) {
  double A[16], AtA[16], w[4], V[16];
  for (int c = 0; c < 4; c++) {
    A[c] = p0[0] * P0.m[8 + c] - P0.m[c]; A[4 + c] = p0[1] * P0.m[8 + c] - P0.m[4 + c];
    A[8 + c] = p1[0] * P1.m[8 + c] - P1.m[c]; A[12 + c] = p1[1] * P1.m[8 + c] - P1.m[4 + c];
  }
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { double a = 0; for (int k = 0; k < 4; k++) a += A[4 * k + i] * A[4 * k + j]; AtA[4 * i + j] = a; }
  jacobi_eig(AtA, 4, w, V);
  if (Xh) for (int i = 0; i < 4; i++) Xh[i] = V[4 * i];
  for (int i = 0; i < 3; i++) X[i] = V[4 * i] / V[12];
}

// solveFrameByPnP (:22-71): refinement of the supplied (R, P) over all triangulated points seen in frame i
bool solveFrameByPnP(double R[9], double Pt[3], int i, const std::vector<SFMFeature>& sfm_f) {
  std::vector<std::array<double, 5>> obs;                   // X Y Z u v
  for (const auto& f : sfm_f) {
    if (!f.state) continue;
    for (const auto& o : f.observation) if (o.first == i) { obs.push_back({f.position[0], f.position[1], f.position[2], o.second[0], o.second[1]}); break; }
  }
  if ((int)obs.size() < 10) return false;                    // < 15 only warns in the reference
  Vec x(7); R2q(R, x.data()); qnorm(x.data()); x[4] = Pt[0]; x[5] = Pt[1]; x[6] = Pt[2];
  Lsq P; P.m = 6; P.max_iter = 30;
  P.residual = [&](const Vec& s, Vec& r) {
    r.resize(2 * obs.size());
    for (size_t k = 0; k < obs.size(); k++) { double p[3]; qrot(s.data(), obs[k].data(), p); for (int a = 0; a < 3; a++) p[a] += s[4 + a]; r[2 * k] = p[0] / p[2] - obs[k][3]; r[2 * k + 1] = p[1] / p[2] - obs[k][4]; }
  };
  P.plus = [](const Vec& s, const double* d, Vec& sn) { sn.resize(7); q_plus(s.data(), d, sn.data()); for (int a = 0; a < 3; a++) sn[4 + a] = s[4 + a] + d[3 + a]; };
  const LsqSummary sm = lm_solve(P, x);
  if (!std::isfinite(sm.cost)) return false;
  q2R(x.data(), R); for (int a = 0; a < 3; a++) Pt[a] = x[4 + a];
  return true;
}

void triangulateTwoFrames(int f0, const Pose34& P0, int f1, const Pose34& P1, std::vector<SFMFeature>& sfm_f) {   // :72-106
  for (auto& f : sfm_f) {
    if (f.state) continue;
    bool has0 = false, has1 = false; double p0[2] = {0, 0}, p1[2] = {0, 0};
    for (const auto& o : f.observation) {
      if (o.first == f0) { p0[0] = o.second[0]; p0[1] = o.second[1]; has0 = true; }
      if (o.first == f1) { p1[0] = o.second[0]; p1[1] = o.second[1]; has1 = true; }
    }
    if (has0 && has1) { triangulatePoint(P0, P1, p0, p1, f.position); f.state = true; }
  }
}

// GlobalSFM::construct (:107-309).  q / T: camera -> c0 (frame l) rotation (x y z w) and position, as the reference returns them.
bool sfm_construct(int frame_num, std::vector<std::array<double, 4>>& q, std::vector<std::array<double, 3>>& T, int l, const double relative_R[9], const double relative_T[3],
                   std::vector<SFMFeature>& sfm_f, std::map<int, std::array<double, 3>>& tracked) {
  q.assign(frame_num, {0, 0, 0, 1}); T.assign(frame_num, {0, 0, 0});
  R2q(relative_R, q[frame_num - 1].data());
  for (int a = 0; a < 3; a++) T[frame_num - 1][a] = relative_T[a];
  std::vector<std::array<double, 9>> cR(frame_num); std::vector<std::array<double, 3>> cT(frame_num); std::vector<Pose34> Pose(frame_num);
  auto set_from_q = [&](int k) { double qi[4]; qinv(q[k].data(), qi); q2R(qi, cR[k].data()); double o[3]; mat3_vec(cR[k].data(), T[k].data(), o); for (int a = 0; a < 3; a++) cT[k][a] = -o[a]; make_pose(cR[k].data(), cT[k].data(), Pose[k]); };
  set_from_q(l); set_from_q(frame_num - 1);
  for (int i = l; i < frame_num - 1; i++) {
    if (i > l) {
      cR[i] = cR[i - 1]; cT[i] = cT[i - 1];
      if (!solveFrameByPnP(cR[i].data(), cT[i].data(), i, sfm_f)) return false;
      make_pose(cR[i].data(), cT[i].data(), Pose[i]);
    }
    triangulateTwoFrames(i, Pose[i], frame_num - 1, Pose[frame_num - 1], sfm_f);
  }
  for (int i = l + 1; i < frame_num - 1; i++) triangulateTwoFrames(l, Pose[l], i, Pose[i], sfm_f);
  for (int i = l - 1; i >= 0; i--) {
    cR[i] = cR[i + 1]; cT[i] = cT[i + 1];
    if (!solveFrameByPnP(cR[i].data(), cT[i].data(), i, sfm_f)) return false;
    make_pose(cR[i].data(), cT[i].data(), Pose[i]);
    triangulateTwoFrames(i, Pose[i], l, Pose[l], sfm_f);
  }
  for (auto& f : sfm_f) {
    if (f.state || f.observation.size() < 2) continue;
    const auto& o0 = f.observation.front(); const auto& o1 = f.observation.back();
    triangulatePoint(Pose[o0.first], Pose[o1.first], o0.second.data(), o1.second.data(), f.position);
    f.state = true;
  }
  // full bundle adjustment (:218-278): rotation of frame l and the translations of frames l and frame_num - 1 constant
  std::vector<int> rot_off(frame_num, -1), tr_off(frame_num, -1), pt_off(sfm_f.size(), -1);
  int m = 0;
  for (int i = 0; i < frame_num; i++) { if (i != l) { rot_off[i] = m; m += 3; } if (i != l && i != frame_num - 1) { tr_off[i] = m; m += 3; } }
  int npt = 0;
  for (size_t j = 0; j < sfm_f.size(); j++) if (sfm_f[j].state) { pt_off[j] = m; m += 3; npt++; }
  Vec x((size_t)7 * frame_num + 3 * sfm_f.size());
  for (int i = 0; i < frame_num; i++) { R2q(cR[i].data(), &x[7 * i]); qnorm(&x[7 * i]); for (int a = 0; a < 3; a++) x[7 * i + 4 + a] = cT[i][a]; }
  const size_t pbase = (size_t)7 * frame_num;
  for (size_t j = 0; j < sfm_f.size(); j++) for (int a = 0; a < 3; a++) x[pbase + 3 * j + a] = sfm_f[j].position[a];
  Lsq P; P.m = m; P.max_iter = 50;
  P.residual = [&](const Vec& s, Vec& r) {
    r.clear();
    for (size_t j = 0; j < sfm_f.size(); j++) {
      if (!sfm_f[j].state) continue;
      for (const auto& o : sfm_f[j].observation) {
        double p[3]; qrot(&s[7 * o.first], &s[pbase + 3 * j], p);
        for (int a = 0; a < 3; a++) p[a] += s[7 * o.first + 4 + a];
        r.push_back(p[0] / p[2] - o.second[0]); r.push_back(p[1] / p[2] - o.second[1]);
      }
    }
  };
  P.plus = [&](const Vec& s, const double* d, Vec& sn) {
    sn = s;
    for (int i = 0; i < frame_num; i++) {
      if (rot_off[i] >= 0) q_plus(&s[7 * i], d + rot_off[i], &sn[7 * i]);
      if (tr_off[i] >= 0) for (int a = 0; a < 3; a++) sn[7 * i + 4 + a] = s[7 * i + 4 + a] + d[tr_off[i] + a];
    }
    for (size_t j = 0; j < sfm_f.size(); j++) if (pt_off[j] >= 0) for (int a = 0; a < 3; a++) sn[pbase + 3 * j + a] = s[pbase + 3 * j + a] + d[pt_off[j] + a];
  };
  const LsqSummary sm = lm_solve(P, x);
  if (!(sm.converged || sm.cost < 5e-3)) return false;
  for (int i = 0; i < frame_num; i++) {
    double qi[4]; qinv(&x[7 * i], qi); for (int a = 0; a < 4; a++) q[i][a] = qi[a];
    double o[3]; qrot(qi, &x[7 * i + 4], o); for (int a = 0; a < 3; a++) T[i][a] = -o[a];
  }
  for (size_t j = 0; j < sfm_f.size(); j++) if (sfm_f[j].state) tracked[sfm_f[j].id] = {x[pbase + 3 * j], x[pbase + 3 * j + 1], x[pbase + 3 * j + 2]};
  return true;
}

// vendored cv::recoverPose (solve_5pts.cpp:24-190) on normalised points with the identity camera matrix
int recoverPose(const double E[9], const std::vector<std::array<double, 2>>& p1, const std::vector<std::array<double, 2>>& p2, double R[9], double t[3], std::vector<uint8_t>& mask) {
  double U[9], S[3], Vt[9];
  svd3(E, U, S, Vt);
  if (det3(U) < 0) for (double& v : U) v = -v;
  if (det3(Vt) < 0) for (double& v : Vt) v = -v;
  const double W[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1}, Wt[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1};
  double R1[9], R2[9], tmp[9], tt[3] = {U[2], U[5], U[8]};
  mat3_mul(U, W, tmp); mat3_mul(tmp, Vt, R1); mat3_mul(U, Wt, tmp); mat3_mul(tmp, Vt, R2);
  const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, z3[3] = {0, 0, 0}, tn[3] = {-tt[0], -tt[1], -tt[2]};
  Pose34 P0; make_pose(I3, z3, P0);
  const double* Rs[4] = {R1, R2, R1, R2}; const double* ts[4] = {tt, tt, tn, tn};
  const double dist = 50.0; int good[4] = {0, 0, 0, 0}; std::vector<uint8_t> m[4];
  for (int c = 0; c < 4; c++) {
    Pose34 Pc; make_pose(Rs[c], ts[c], Pc); m[c].assign(p1.size(), 0);
    for (size_t k = 0; k < p1.size(); k++) {
      double X[3], Xh[4];
      triangulatePoint(P0, Pc, p1[k].data(), p2[k].data(), X, Xh);
      bool ok = Xh[2] * Xh[3] > 0 && X[2] < dist;
      const double z2 = Pc.m[8] * X[0] + Pc.m[9] * X[1] + Pc.m[10] * X[2] + Pc.m[11];
      ok = ok && z2 > 0 && z2 < dist && (mask.empty() || mask[k]);
      m[c][k] = ok; good[c] += ok;
    }
  }
  int best;
  if (good[0] >= good[1] && good[0] >= good[2] && good[0] >= good[3]) best = 0;
  else if (good[1] >= good[0] && good[1] >= good[2] && good[1] >= good[3]) best = 1;
  else if (good[2] >= good[0] && good[2] >= good[1] && good[2] >= good[3]) best = 2;
  else best = 3;
  std::memcpy(R, Rs[best], sizeof(double) * 9); std::memcpy(t, ts[best], sizeof(double) * 3);
  mask = m[best];
  return good[best];
}

// Utility::g2R (utility/utility.cpp:3-13)
void g2R(const double g[3], double R0[9]) {
  const double n = std::sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
  const double v0[3] = {g[0] / n, g[1] / n, g[2] / n};
  const double c = v0[2];                                    // v1 = (0, 0, 1)
  double q[4];
  if (c < -1.0 + 1e-12) { q[0] = 1; q[1] = 0; q[2] = 0; q[3] = 0; }   // opposite vectors: any axis orthogonal to z
  else { const double ax[3] = {v0[1], -v0[0], 0.0}, s = std::sqrt((1 + c) * 2), inv = 1 / s; q[0] = ax[0] * inv; q[1] = ax[1] * inv; q[2] = 0; q[3] = 0.5 * s; }
  double Rq[9]; q2R(q, Rq);
  const double yaw = std::atan2(Rq[3], Rq[0]);              // R2ypr(R0).x() in radians
  const double cy = std::cos(-yaw), sy = std::sin(-yaw);
  const double Rz[9] = {cy, -sy, 0, sy, cy, 0, 0, 0, 1};      // ypr2R(-yaw, 0, 0)
  mat3_mul(Rz, Rq, R0);
}

}  // namespace

// estimator.cpp:873-901 + solve_5pts.cpp:193-230
static bool relative_pose(Estimator& e, vils_frontend*& fe, double relative_R[9], double relative_T[3], int& l) {
  const int W = e.WINDOW_SIZE;
  for (int i = 0; i < W; i++) {
    std::vector<std::array<double, 2>> ll, rr;
    for (const auto& it : e.feature) {                       // FeatureManager::getCorresponding (feature_manager.cpp:129-148)
      if (it.start_frame <= i && it.endFrame() >= W) {
        const auto& a = it.feature_per_frame[i - it.start_frame]; const auto& b = it.feature_per_frame[W - it.start_frame];
        ll.push_back({a.point[0], a.point[1]}); rr.push_back({b.point[0], b.point[1]});
      }
    }
    if (ll.size() <= 20) continue;
    double sum = 0; for (size_t k = 0; k < ll.size(); k++) sum += std::hypot(ll[k][0] - rr[k][0], ll[k][1] - rr[k][1]);
    if (sum / ll.size() * 460 <= 30) continue;
    // MotionEstimator::solveRelativeRT: cv::findFundamentalMat(ll, rr, FM_RANSAC, 0.3 / 460, 0.99, mask) on normalised coordinates.  The
    // device RANSAC works in pixel-like units: scale by the virtual focal length 460 (threshold 0.3), then E = K F K with K = diag(460, 460, 1).
    if (!fe) { e.last_status = vils_frontend_create(480, 640, 2048, e.device(), &fe); if (e.last_status != VILS_OK) return false; }
    const int n = (int)std::min<size_t>(ll.size(), 2048);
    std::vector<float> a(2 * (size_t)n), b(2 * (size_t)n); std::vector<uint8_t> mask(n); double F[9];
    for (int k = 0; k < n; k++) { a[2 * k] = (float)(460.0 * ll[k][0]); a[2 * k + 1] = (float)(460.0 * ll[k][1]); b[2 * k] = (float)(460.0 * rr[k][0]); b[2 * k + 1] = (float)(460.0 * rr[k][1]); }
    e.last_status = vils_reject_with_f(fe, a.data(), b.data(), n, 0.3, mask.data(), F);
    if (e.last_status != VILS_OK) return false;
    const double K[3] = {460.0, 460.0, 1.0}; double E[9];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) E[3 * r + c] = F[3 * r + c] * K[r] * K[c];
    ll.resize(n); rr.resize(n);
    double R[9], t[3];
    const int inliers = recoverPose(E, ll, rr, R, t, mask);
    double Rt[9], o[3]; mat3_T(R, Rt); mat3_vec(Rt, t, o);
    std::memcpy(relative_R, Rt, sizeof(Rt)); for (int k = 0; k < 3; k++) relative_T[k] = -o[k];   // Rotation = R^T, Translation = -R^T T
    if (inliers > 12) { l = i; return true; }
  }
  return false;
}

// initial_aligment.cpp:193-347 + :349-519 + estimator.cpp:768-871
static bool visual_initial_align(Estimator& e) {
  auto& frames = e.all_image_frame;
  const int F = (int)frames.size();
  if (F < 2) return false;
  std::vector<Estimator::ImageFrame*> fr; for (auto& kv : frames) fr.push_back(&kv.second);
  // pre-integration of every image interval at the biases it was started with (IntegrationBase of the ImageFrame), one device call
  std::vector<int32_t> off(1, 0); std::vector<double> dt, acc, gyr, acc0, gyr0, ba, bg; std::vector<int> of_frame;
  for (int k = 1; k < F; k++) {
    const auto& b = fr[k]->pre;
    if (b.dt.empty()) return false;
    dt.insert(dt.end(), b.dt.begin(), b.dt.end()); acc.insert(acc.end(), b.acc.begin(), b.acc.end()); gyr.insert(gyr.end(), b.gyr.begin(), b.gyr.end());
    for (int i = 0; i < 3; i++) { acc0.push_back(b.acc0[i]); gyr0.push_back(b.gyr0[i]); ba.push_back(b.ba[i]); bg.push_back(b.bg[i]); }
    off.push_back((int32_t)dt.size()); of_frame.push_back(k);
  }
  std::vector<vils_preint> pre(F - 1);
  e.last_status = vils_preintegrate(F - 1, off.data(), dt.data(), acc.data(), gyr.data(), acc0.data(), gyr0.data(), ba.data(), bg.data(), e.config().imu_noise, pre.data(), e.device());
  if (e.last_status != VILS_OK) return false;
  auto Jblk = [&](const vils_preint& p, int r0, int c0, double* M) { for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) M[3 * r + c] = p.jacobian[(c0 + c) * 15 + r0 + r]; };   // column-major 15 x 15
  const int O_P = 0, O_R = 3, O_V = 6, O_BA = 9, O_BG = 12;
  const bool est_ex = e.config().estimate_extrinsic != 0;
  // ---- Estimate_ric_td_bg (:193-347): x = [BGS (3F) | RIC_Q x y z w | Td (F)]
  double RIC[9]; q2R(e.ric, RIC);
  std::vector<std::array<double, 4>> qc(F);                  // frame R (camera -> c0) as quaternions
  for (int k = 0; k < F; k++) { R2q(fr[k]->R, qc[k].data()); qnorm(qc[k].data()); }
  std::vector<std::array<double, 3>> wlast(F - 1);           // gyr_1 == gyr_0 after the last push_back of the interval (integration_base.h:128)
  for (int k = 0; k < F - 1; k++) { const auto& b = fr[k + 1]->pre; const size_t n = b.dt.size(); for (int a = 0; a < 3; a++) wlast[k][a] = b.gyr[3 * (n - 1) + a]; }
  Vec x((size_t)3 * F + 4 + F, 0.0);
  { double q[4]; R2q(RIC, q); qnorm(q); for (int a = 0; a < 4; a++) x[3 * F + a] = q[a]; }
  Lsq P1; P1.m = 3 * F + (est_ex ? 3 : 0) + F; P1.max_iter = 50;
  const int ric_t = 3 * F, td_t = 3 * F + (est_ex ? 3 : 0);
  P1.residual = [&](const Vec& s, Vec& r) {
    r.resize(3 * (size_t)(F - 1));
    const double* rq = &s[3 * F];
    const double qbc[4] = {rq[0], rq[1], rq[2], rq[3]}, qcb[4] = {-rq[0], -rq[1], -rq[2], rq[3]};
    for (int i = 0; i < F - 1; i++) {
      const double* w = wlast[i].data(); const double tdi = s[3 * F + 4 + i], tdj = s[3 * F + 4 + i + 1];
      const double Ql[4] = {-0.5 * w[0] * tdj, -0.5 * w[1] * tdj, -0.5 * w[2] * tdj, 1.0}, Qr[4] = {0.5 * w[0] * tdi, 0.5 * w[1] * tdi, 0.5 * w[2] * tdi, 1.0};
      double qcj[4]; qinv(qc[i + 1].data(), qcj);            // Quaterniond(frame_j.R^T)
      double J[9]; Jblk(pre[i], O_R, O_BG, J);
      double jb[3]; mat3_vec(J, &s[3 * i], jb);
      const double Qjbg[4] = {0.5 * jb[0], 0.5 * jb[1], 0.5 * jb[2], 1.0};
      double a[4], b[4];
      qmul(Ql, qbc, a); qmul(a, qcj, b); qmul(b, qc[i].data(), a); qmul(a, qcb, b); qmul(b, Qr, a); qmul(a, pre[i].delta_q, b); qmul(b, Qjbg, a);
      r[3 * i] = 2 * a[0]; r[3 * i + 1] = 2 * a[1]; r[3 * i + 2] = 2 * a[2];
    }
  };
  P1.plus = [&](const Vec& s, const double* d, Vec& sn) {
    sn = s;
    for (int k = 0; k < 3 * F; k++) sn[k] = clampd(s[k] + d[k], -0.1, 0.1);
    if (est_ex) q_plus(&s[3 * F], d + ric_t, &sn[3 * F]);
    for (int k = 0; k < F; k++) sn[3 * F + 4 + k] = clampd(s[3 * F + 4 + k] + d[td_t + k], -0.1, 0.1);
  };
  const LsqSummary s1 = lm_solve(P1, x);
  if (s1.cost > 1e-5) return false;
  { double q[4] = {x[3 * F], x[3 * F + 1], x[3 * F + 2], x[3 * F + 3]}; qnorm(q); q2R(q, RIC); }
  double td_sum = 0;
  for (int j = 0; j < F - 1; j++) {
    const double tt = x[3 * F + 4 + j]; td_sum += tt;
    const double bgs[3] = {x[3 * j], x[3 * j + 1], x[3 * j + 2]};
    for (int a = 0; a < 3; a++) e.Bgs[j][a] = bgs[a];
    // init_refine_delta_pvq_bgs (integration_base.h:160-167)
    double M[9], v[3];
    Jblk(pre[j], O_R, O_BG, M); mat3_vec(M, bgs, v);
    { const double dq[4] = {0.5 * v[0], 0.5 * v[1], 0.5 * v[2], 1.0}; double qn[4]; qmul(pre[j].delta_q, dq, qn); for (int a = 0; a < 4; a++) pre[j].delta_q[a] = qn[a]; }
    Jblk(pre[j], O_V, O_BG, M); mat3_vec(M, bgs, v); for (int a = 0; a < 3; a++) pre[j].delta_v[a] += v[a];
    Jblk(pre[j], O_P, O_BG, M); mat3_vec(M, bgs, v); for (int a = 0; a < 3; a++) pre[j].delta_p[a] += v[a];
    // frame_i.R = R * RIC^T * Qr(gyr_0 of the next interval, Td_j)   (the last frame keeps its camera rotation: reference loop bound)
    const double* w = wlast[j].data();
    const double Qr[4] = {0.5 * w[0] * tt, 0.5 * w[1] * tt, 0.5 * w[2] * tt, 1.0};
    double RICt[9], qrci[4], a[4], b[4]; mat3_T(RIC, RICt); R2q(RICt, qrci);
    qmul(qc[j].data(), qrci, a); qmul(a, Qr, b); q2R(b, fr[j]->R);
  }
  {   // the newest frame (see the header: the reference leaves it in the camera frame)
    const double tt = x[3 * F + 4 + F - 1]; const double* w = wlast[F - 2].data();
    const double Qr[4] = {0.5 * w[0] * tt, 0.5 * w[1] * tt, 0.5 * w[2] * tt, 1.0};
    double RICt[9], qrci[4], a[4], b[4]; mat3_T(RIC, RICt); R2q(RICt, qrci);
    qmul(qc[F - 1].data(), qrci, a); qmul(a, Qr, b); q2R(b, fr[F - 1]->R);
  }
  // ---- Estimate_vel_g_s_tic (:349-519): x = [velocity (3F) | s (F) | bas (3F) | pbc (3) | g_normal (3)] — linear with box constraints
  const double gnorm = e.config().gravity[2];
  Vec y((size_t)3 * F + F + 3 * F + 3 + 3, 0.0);
  const int s_o = 3 * F, b_o = 4 * F, p_o = 7 * F, g_o = 7 * F + 3;
  for (int a = 0; a < 3; a++) { y[p_o + a] = e.tic[a]; y[g_o + a] = e.G_DIRECTION[a]; }
  Lsq P2; P2.m = 7 * F + (est_ex ? 3 : 0) + 3; P2.max_iter = 100;
  const int pbc_t = 7 * F, gn_t = 7 * F + (est_ex ? 3 : 0);
  P2.residual = [&](const Vec& s, Vec& r) {
    r.resize(6 * (size_t)(F - 1));
    for (int k = 0; k < F - 1; k++) {
      const vils_preint& pi = pre[k];
      const double dtk = pi.sum_dt;
      double Rbic0[9], tmp[9], JP[9], JV[9];
      mat3_T(fr[k]->R, Rbic0);                               // rbic0 = frame_i.R^T, rcobj = frame_j.R
      const double* Rcobj = fr[k + 1]->R;
      Jblk(pi, O_P, O_BA, JP); Jblk(pi, O_V, O_BA, JV);
      const double* Vi = &s[3 * k]; const double* Vj = &s[3 * (k + 1)]; const double* bas = &s[b_o + 3 * k]; const double* pbc = &s[p_o];
      const double Gc0[3] = {gnorm * s[g_o], gnorm * s[g_o + 1], gnorm * s[g_o + 2]};
      const double si = s[s_o + k], sj = s[s_o + k + 1];
      double a1[3], a2[3], a3[3], a4[3], dpc[3];
      mat3_vec(JP, bas, a1);
      mat3_mul(Rbic0, Rcobj, tmp); mat3_vec(tmp, pbc, a2);
      for (int a = 0; a < 3; a++) dpc[a] = sj * fr[k + 1]->T[a] - si * fr[k]->T[a];
      mat3_vec(Rbic0, dpc, a3); mat3_vec(Rbic0, Gc0, a4);
      for (int a = 0; a < 3; a++) r[6 * k + a] = pi.delta_p[a] + a1[a] - pbc[a] + a2[a] - a3[a] + Vi[a] * dtk - 0.5 * a4[a] * dtk * dtk;
      double b1[3], b2[3], b3[3], in[3], b4[3];
      mat3_vec(JV, bas, b1); mat3_vec(Rcobj, Vj, b2); mat3_vec(fr[k]->R, Vi, b3);   // Rc0bi = Rbic0^T = frame_i.R
      for (int a = 0; a < 3; a++) in[a] = b2[a] - b3[a] + Gc0[a] * dtk;
      mat3_vec(Rbic0, in, b4);
      for (int a = 0; a < 3; a++) r[6 * k + 3 + a] = pi.delta_v[a] + b1[a] - b4[a];
    }
  };
  // The problem as the reference poses it is rank deficient (7F + 6 unknowns against 6 (F - 1) equations: a lever arm, per-interval
  // accelerometer biases and per-frame scales can stand in for one another) and starts from zeros with s >= 0 active, so what ceres returns
  // depends on its internal path.  Here the SAME bounded problem is started from the closed-form solution of the well-posed alignment that the
  // reference file itself carries: LinearAlignment + RefineGravity (initial_aligment.cpp:143-191,69-141: velocities, gravity, ONE scale).
  {
    const int n_state = 3 * F + 3 + 1;
    auto solve_linear = [&](bool refine, const double g0[3], const double lxly[6], Vec& sol) -> bool {
      const int ng = refine ? 2 : 3, ns = 3 * F + ng + 1;
      Vec A((size_t)ns * ns, 0.0), b(ns, 0.0);
      for (int i = 0; i < F - 1; i++) {
        const vils_preint& pi = pre[i]; const double dtk = pi.sum_dt;
        double RiT[9], RiTRj[9]; mat3_T(fr[i]->R, RiT); mat3_mul(RiT, fr[i + 1]->R, RiTRj);
        const int nc = 6 + ng + 1;                          // tmp_A columns: v_i (3) v_j (3) g (ng) s (1)
        double tA[6 * 10] = {0}, tb[6] = {0};
        for (int a = 0; a < 3; a++) { tA[a * nc + a] = -dtk; tA[(3 + a) * nc + a] = -1.0; for (int c = 0; c < 3; c++) tA[(3 + a) * nc + 3 + c] = RiTRj[3 * a + c]; }
        for (int a = 0; a < 3; a++) for (int c = 0; c < ng; c++) {
          double gp = 0, gv = 0;                              // R_i^T dt^2 / 2 [I | lxly], R_i^T dt [I | lxly]
          for (int k = 0; k < 3; k++) { const double basis = refine ? lxly[2 * k + c] : (k == c ? 1.0 : 0.0); gp += RiT[3 * a + k] * basis; gv += RiT[3 * a + k] * basis; }
          tA[a * nc + 6 + c] = gp * dtk * dtk / 2; tA[(3 + a) * nc + 6 + c] = gv * dtk;
        }
        double dT[3], rdT[3], o1[3], o2[3] = {0, 0, 0}, o3[3] = {0, 0, 0};
        for (int a = 0; a < 3; a++) dT[a] = fr[i + 1]->T[a] - fr[i]->T[a];
        mat3_vec(RiT, dT, rdT); mat3_vec(RiTRj, e.tic, o1);
        if (refine) { mat3_vec(RiT, g0, o2); for (int a = 0; a < 3; a++) { o3[a] = o2[a] * dtk; o2[a] *= dtk * dtk / 2; } }
        for (int a = 0; a < 3; a++) { tA[a * nc + 6 + ng] = rdT[a] / 100.0; tb[a] = pi.delta_p[a] + o1[a] - e.tic[a] - o2[a]; tb[3 + a] = pi.delta_v[a] - o3[a]; }
        // r_A = tmp_A^T tmp_A scattered: v_i, v_j at 3 i, the tail (g, s) at the end
        auto col = [&](int c) { return c < 6 ? 3 * i + c : ns - (ng + 1) + (c - 6); };
        for (int c1 = 0; c1 < nc; c1++) {
          double rb = 0; for (int r = 0; r < 6; r++) rb += tA[r * nc + c1] * tb[r];
          b[col(c1)] += rb;
          for (int c2 = 0; c2 < nc; c2++) { double ra = 0; for (int r = 0; r < 6; r++) ra += tA[r * nc + c1] * tA[r * nc + c2]; A[(size_t)col(c1) * ns + col(c2)] += ra; }
        }
      }
      for (double& v : A) v *= 1000.0;
      for (double& v : b) v *= 1000.0;
      if (!chol_solve(A, b, ns)) return false;
      sol = b; return true;
    };
    Vec sol; const double z3[3] = {0, 0, 0}; double lx[6] = {0};
    bool seeded = solve_linear(false, z3, lx, sol);
    double gl[3] = {0, 0, 0}; double sc = 0;
    if (seeded) {
      for (int a = 0; a < 3; a++) gl[a] = sol[n_state - 4 + a];
      sc = sol[n_state - 1] / 100.0;
      const double gn = std::sqrt(gl[0] * gl[0] + gl[1] * gl[1] + gl[2] * gl[2]);
      seeded = std::fabs(gn - gnorm) <= 1.0 && sc >= 0;      // LinearAlignment's own acceptance test (:179-182)
      if (seeded) {
        double g0[3] = {gl[0] / gn * gnorm, gl[1] / gn * gnorm, gl[2] / gn * gnorm};
        for (int k = 0; k < 4; k++) {                        // RefineGravity (:69-141): 2-dof correction in the tangent plane of g0
          double a_[3] = {g0[0] / gnorm, g0[1] / gnorm, g0[2] / gnorm}, tmp[3] = {0, 0, 1};
          if (a_[0] == 0 && a_[1] == 0 && a_[2] == 1) { tmp[0] = 1; tmp[2] = 0; }
          const double dtp = a_[0] * tmp[0] + a_[1] * tmp[1] + a_[2] * tmp[2];
          double bb[3] = {tmp[0] - a_[0] * dtp, tmp[1] - a_[1] * dtp, tmp[2] - a_[2] * dtp};
          const double bn = std::sqrt(bb[0] * bb[0] + bb[1] * bb[1] + bb[2] * bb[2]); for (double& v : bb) v /= bn;
          const double cc[3] = {a_[1] * bb[2] - a_[2] * bb[1], a_[2] * bb[0] - a_[0] * bb[2], a_[0] * bb[1] - a_[1] * bb[0]};
          for (int r = 0; r < 3; r++) { lx[2 * r] = bb[r]; lx[2 * r + 1] = cc[r]; }
          Vec s2v;
          if (!solve_linear(true, g0, lx, s2v)) break;
          const double dg0 = s2v[3 * F], dg1 = s2v[3 * F + 1];
          double gn2 = 0; for (int r = 0; r < 3; r++) { g0[r] += lx[2 * r] * dg0 + lx[2 * r + 1] * dg1; gn2 += g0[r] * g0[r]; }
          gn2 = std::sqrt(gn2); for (int r = 0; r < 3; r++) g0[r] = g0[r] / gn2 * gnorm;
          sol = s2v; sc = s2v[3 * F + 2] / 100.0;
        }
        if (sc > 0) {
          for (int k = 0; k < 3 * F; k++) y[k] = sol[k];
          for (int k = 0; k < F; k++) y[s_o + k] = sc;
          for (int a = 0; a < 3; a++) y[g_o + a] = g0[a] / gnorm;
        }
        if (getenv("VILS_INIT_DEBUG")) fprintf(stderr, "[init] LinearAlignment + RefineGravity: scale %.4f, g %.3f %.3f %.3f\n", sc, g0[0], g0[1], g0[2]);
      }
    }
  }
  P2.plus = [&](const Vec& s, const double* d, Vec& sn) {
    sn = s;
    for (int k = 0; k < 3 * F; k++) sn[k] = s[k] + d[k];
    for (int k = 0; k < F; k++) sn[s_o + k] = std::max(0.0, s[s_o + k] + d[s_o + k]);
    for (int k = 0; k < 3 * F; k++) sn[b_o + k] = clampd(s[b_o + k] + d[b_o + k], -0.2, 0.2);
    if (est_ex) for (int a = 0; a < 3; a++) sn[p_o + a] = s[p_o + a] + d[pbc_t + a];
    for (int a = 0; a < 3; a++) sn[g_o + a] = s[g_o + a] + d[gn_t + a];
  };
  const LsqSummary s2 = lm_solve(P2, y);
  if (getenv("VILS_INIT_DEBUG")) {
    fprintf(stderr, "[init] ric/td/bg cost %.3e (%d it), vel/g/s cost %.3e (%d it)\n  s:", s1.cost, s1.iterations, s2.cost, s2.iterations);
    for (int k = 0; k < F; k++) fprintf(stderr, " %.4f", y[s_o + k]);
    fprintf(stderr, "\n  g_normal %.4f %.4f %.4f  pbc %.4f %.4f %.4f\n  |T|:", y[g_o], y[g_o + 1], y[g_o + 2], y[p_o], y[p_o + 1], y[p_o + 2]);
    for (int k = 0; k < F; k++) fprintf(stderr, " %.3f", std::sqrt(fr[k]->T[0] * fr[k]->T[0] + fr[k]->T[1] * fr[k]->T[1] + fr[k]->T[2] * fr[k]->T[2]));
    fprintf(stderr, "\n  bas0 %.4f %.4f %.4f  v0 %.3f %.3f %.3f  td0 %.5f bg0 %.5f %.5f %.5f\n", y[b_o], y[b_o + 1], y[b_o + 2], y[0], y[1], y[2], x[3 * F + 4], x[0], x[1], x[2]);
  }
  if (s2.cost > 5e-3) return false;
  { const double n = std::sqrt(y[g_o] * y[g_o] + y[g_o + 1] * y[g_o + 1] + y[g_o + 2] * y[g_o + 2]); if (!(n > 0)) return false; for (int a = 0; a < 3; a++) e.g[a] = gnorm * y[g_o + a] / n; }
  for (int a = 0; a < 3; a++) e.tic[a] = y[p_o + a];
  for (int j = 0; j < F && j <= e.WINDOW_SIZE; j++) for (int a = 0; a < 3; a++) e.Bas[j][a] = y[b_o + 3 * j + a];
  // ---- visualInitialAlign (estimator.cpp:787-868)
  for (int i = 0; i <= e.frame_count; i++) {
    auto it = frames.find(e.Headers[i].stamp);
    if (it == frames.end()) return false;
    const double* Ri = it->second.R; const double* Pi = it->second.T;
    double o[3]; mat3_vec(Ri, e.tic, o);
    for (int a = 0; a < 3; a++) e.Ps[i][a] = y[s_o + i] * Pi[a] - o[a];
    R2q(Ri, e.Qs[i].data()); qnorm(e.Qs[i].data());
    it->second.is_key_frame = true;
  }
  e.td = td_sum / e.frame_count;                            // see the header: the reference reads uninitialised memory here
  { double q[4]; R2q(RIC, q); qnorm(q); for (int a = 0; a < 4; a++) e.ric[a] = q[a]; }
  for (auto& f : e.feature) f.estimated_depth = f.lidar_depth_flag ? f.estimated_depth : -1.0;
  e.triangulate();                                           // f_manager.triangulate(Ps, &TIC, &RIC)
  if (e.last_status != VILS_OK) return false;
  int kv = -1;
  for (auto& kvp : frames) if (kvp.second.is_key_frame) { kv++; if (kv <= e.WINDOW_SIZE) { double o[3]; mat3_vec(kvp.second.R, &y[3 * kv], o); for (int a = 0; a < 3; a++) e.Vs[kv][a] = o[a]; } }
  double R0[9]; g2R(e.g, R0);
  { double gn[3]; mat3_vec(R0, e.g, gn); for (int a = 0; a < 3; a++) e.g[a] = gn[a]; }
  double q0[4]; R2q(R0, q0); qnorm(q0);
  for (int i = 0; i <= e.frame_count; i++) {
    double o[3], qn[4];
    mat3_vec(R0, e.Ps[i].data(), o); for (int a = 0; a < 3; a++) e.Ps[i][a] = o[a];
    mat3_vec(R0, e.Vs[i].data(), o); for (int a = 0; a < 3; a++) e.Vs[i][a] = o[a];
    qmul(q0, e.Qs[i].data(), qn); qnorm(qn); for (int a = 0; a < 4; a++) e.Qs[i][a] = qn[a];
  }
  return true;
}

// estimator.cpp:618-766
bool Estimator::initialStructure() {
  last_status = VILS_OK;
  // (the IMU-excitation check of :621-647 only logs in the reference)
  std::vector<SFMFeature> sfm_f;
  for (const auto& it : feature) {
    SFMFeature f; f.id = it.feature_id; int imu_j = it.start_frame - 1;
    for (const auto& fp : it.feature_per_frame) { imu_j++; f.observation.push_back({imu_j, {fp.point[0], fp.point[1]}}); }
    sfm_f.push_back(f);
  }
  double relative_R[9], relative_T[3]; int l = 0;
  if (!relative_pose(*this, init_fe_, relative_R, relative_T, l)) return false;
  std::vector<std::array<double, 4>> Q; std::vector<std::array<double, 3>> T; std::map<int, std::array<double, 3>> tracked;
  if (getenv("VILS_INIT_DEBUG")) fprintf(stderr, "[init] relativePose l = %d, T = %.3f %.3f %.3f\n", l, relative_T[0], relative_T[1], relative_T[2]);
  if (!sfm_construct(frame_count + 1, Q, T, l, relative_R, relative_T, sfm_f, tracked)) { marginalization_flag = MARGIN_OLD; return false; }
  // every image frame is a window frame while initialising (slideWindow erases the others, :1731-1747,:1781-1784): no PnP branch needed
  int i = 0;
  for (auto& kv : all_image_frame) {
    if (i > WINDOW_SIZE || kv.first != Headers[i].stamp) return false;
    kv.second.is_key_frame = true; q2R(Q[i].data(), kv.second.R); for (int a = 0; a < 3; a++) kv.second.T[a] = T[i][a];
    i++;
  }
  return visual_initial_align(*this);
}

}  // namespace vils
