// vils_host.cpp — see vils_host.h.  Host glue only: bookkeeping, packing the window for the C-ABI, state shuffling.
// Every number-crunching step (pre-integration, factor evaluation, solve, marginalization, KLT, deskew) is a library call.
#include "vils_host.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace vils {
namespace {

inline void qmul(const double a[4], const double b[4], double o[4]) {   // x y z w, Hamilton
  const double ax = a[0], ay = a[1], az = a[2], aw = a[3], bx = b[0], by = b[1], bz = b[2], bw = b[3];
  o[0] = aw * bx + ax * bw + ay * bz - az * by; o[1] = aw * by + ay * bw + az * bx - ax * bz;
  o[2] = aw * bz + az * bw + ax * by - ay * bx; o[3] = aw * bw - ax * bx - ay * by - az * bz;
}
inline void qrot(const double q[4], const double v[3], double o[3]) {
  const double ux = q[0], uy = q[1], uz = q[2], w = q[3];
  double tx = 2 * (uy * v[2] - uz * v[1]), ty = 2 * (uz * v[0] - ux * v[2]), tz = 2 * (ux * v[1] - uy * v[0]);
  o[0] = v[0] + w * tx + (uy * tz - uz * ty); o[1] = v[1] + w * ty + (uz * tx - ux * tz); o[2] = v[2] + w * tz + (ux * ty - uy * tx);
}
inline void qrot_inv(const double q[4], const double v[3], double o[3]) { const double c[4] = {-q[0], -q[1], -q[2], q[3]}; qrot(c, v, o); }
inline void qnorm(double q[4]) { const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]); for (int i = 0; i < 4; i++) q[i] /= n; }
inline void R2q(const double m[9], double q[4]) {   // Eigen Quaternion(Matrix3)
  double t = m[0] + m[4] + m[8];
  if (t > 0) { t = std::sqrt(t + 1.0); q[3] = 0.5 * t; t = 0.5 / t; q[0] = (m[7] - m[5]) * t; q[1] = (m[2] - m[6]) * t; q[2] = (m[3] - m[1]) * t; }
  else { int i = 0; if (m[4] > m[0]) i = 1; if (m[8] > m[i * 4]) i = 2; const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m[i * 4] - m[j * 4] - m[k * 4] + 1.0); q[i] = 0.5 * t; t = 0.5 / t; q[3] = (m[k * 3 + j] - m[j * 3 + k]) * t; q[j] = (m[j * 3 + i] + m[i * 3 + j]) * t; q[k] = (m[k * 3 + i] + m[i * 3 + k]) * t; }
}

}  // namespace

Estimator::Estimator(const vils_config& cfg, int window_size, int num_iterations) : WINDOW_SIZE(window_size), cfg_(cfg) {
  const int F = WINDOW_SIZE + 1;
  cfg_.max_kf = F;
  Ps.assign(F, {0, 0, 0}); Vs.assign(F, {0, 0, 0}); Bas.assign(F, {0, 0, 0}); Bgs.assign(F, {0, 0, 0}); Qs.assign(F, {0, 0, 0, 1});
  Headers.assign(F, Header()); imu_.assign(F, ImuBuf());
  for (int i = 0; i < 3; i++) g[i] = cfg.gravity[i];
  vils_default_solve_opts(&solve_opts);
  solve_opts.mode = VILS_MODE_LM; solve_opts.max_iters = num_iterations;   // ceres DOGLEG, max_num_iterations (estimator.cpp:1402-1404)
  last_status = vils_ba_create(&cfg_, 1, &ba_);
}
Estimator::~Estimator() { vils_ba_destroy(ba_); }

void Estimator::setParameter(const double r[9], const double t[3], double td_) {
  R2q(r, ric); qnorm(ric); for (int i = 0; i < 3; i++) tic[i] = t[i]; td = td_;
}

void Estimator::clearState() {
  const int F = WINDOW_SIZE + 1;
  Ps.assign(F, {0, 0, 0}); Vs.assign(F, {0, 0, 0}); Bas.assign(F, {0, 0, 0}); Bgs.assign(F, {0, 0, 0}); Qs.assign(F, {0, 0, 0, 1});
  imu_.assign(F, ImuBuf()); feature.clear(); frame_count = 0; first_imu_ = false; solver_flag = INITIAL; prior_n_ = 0;
  prior_J_.clear(); prior_r_.clear(); prior_x0_.clear(); prior_blk_.clear();
}

void Estimator::setFrameState(int k, const double P[3], const double Q[4], const double V[3], const double Ba[3], const double Bg[3]) {
  for (int i = 0; i < 3; i++) { Ps[k][i] = P[i]; Vs[k][i] = V[i]; Bas[k][i] = Ba[i]; Bgs[k][i] = Bg[i]; }
  for (int i = 0; i < 4; i++) Qs[k][i] = Q[i];
  solver_flag = NON_LINEAR;
}

// estimator.cpp:86-120.  The samples are buffered per interval (dt_buf etc.); the mid-point propagation of Rs/Ps/Vs that
// predicts the newest frame is kept here (three lines of host arithmetic), IntegrationBase::push_back is not: all intervals are
// pre-integrated on the GPU in one launch inside optimization().
void Estimator::processIMU(double dt, const double acc[3], const double gyr[3]) {
  if (!first_imu_) { first_imu_ = true; std::memcpy(acc_0_, acc, 24); std::memcpy(gyr_0_, gyr, 24); }
  ImuBuf& b = imu_[frame_count];
  if (!b.started) {   // new IntegrationBase{acc_0, gyr_0, Bas[frame_count], Bgs[frame_count]}
    b.started = true; std::memcpy(b.acc0, acc_0_, 24); std::memcpy(b.gyr0, gyr_0_, 24);
    for (int i = 0; i < 3; i++) { b.ba[i] = Bas[frame_count][i]; b.bg[i] = Bgs[frame_count][i]; }
  }
  if (frame_count != 0) {
    b.dt.push_back(dt); for (int i = 0; i < 3; i++) { b.acc.push_back(acc[i]); b.gyr.push_back(gyr[i]); }
    const int j = frame_count;
    double a0[3], un0[3], un1[3], w[3];
    for (int i = 0; i < 3; i++) a0[i] = acc_0_[i] - Bas[j][i];
    qrot(Qs[j].data(), a0, un0); for (int i = 0; i < 3; i++) un0[i] -= g[i];
    for (int i = 0; i < 3; i++) w[i] = 0.5 * (gyr_0_[i] + gyr[i]) - Bgs[j][i];
    const double dq[4] = {w[0] * dt / 2, w[1] * dt / 2, w[2] * dt / 2, 1.0};   // Utility::deltaQ
    double qn[4]; qmul(Qs[j].data(), dq, qn); qnorm(qn); for (int i = 0; i < 4; i++) Qs[j][i] = qn[i];
    double a1[3]; for (int i = 0; i < 3; i++) a1[i] = acc[i] - Bas[j][i];
    qrot(Qs[j].data(), a1, un1); for (int i = 0; i < 3; i++) un1[i] -= g[i];
    for (int i = 0; i < 3; i++) { const double ua = 0.5 * (un0[i] + un1[i]); Ps[j][i] += dt * Vs[j][i] + 0.5 * dt * dt * ua; Vs[j][i] += dt * ua; }
  }
  std::memcpy(acc_0_, acc, 24); std::memcpy(gyr_0_, gyr, 24);
}

double Estimator::compensatedParallax2(const FeaturePerId& it, int fc) const {
  const FeaturePerFrame& fi = it.feature_per_frame[fc - 2 - it.start_frame];
  const FeaturePerFrame& fj = it.feature_per_frame[fc - 1 - it.start_frame];
  const double du = fi.point[0] / fi.point[2] - fj.point[0], dv = fi.point[1] / fi.point[2] - fj.point[1];
  return std::sqrt(du * du + dv * dv);
}

bool Estimator::addFeatureCheckParallax(int fc, const ImageFeatures& image, double td_) {
  double parallax_sum = 0; int parallax_num = 0, last_track_num = 0;
  for (const auto& id_pts : image) {
    const Feature8& p = id_pts.second[0].second;
    FeaturePerFrame f{}; f.point[0] = p[0]; f.point[1] = p[1]; f.point[2] = p[2]; f.uv[0] = p[3]; f.uv[1] = p[4]; f.velocity[0] = p[5]; f.velocity[1] = p[6];
    f.depth = p[7]; f.cur_td = td_;
    const int id = id_pts.first;
    auto it = std::find_if(feature.begin(), feature.end(), [id](const FeaturePerId& x) { return x.feature_id == id; });
    if (it == feature.end()) {
      FeaturePerId n; n.feature_id = id; n.start_frame = fc;
      n.estimated_depth = f.depth > 0 ? f.depth : -1.0; n.lidar_depth_flag = f.depth > 0;   // FeaturePerId ctor (feature_manager.h:62-72)
      n.feature_per_frame.push_back(f); feature.push_back(n);
    } else {
      it->feature_per_frame.push_back(f); last_track_num++;
      if (f.depth > 0 && !it->lidar_depth_flag) { it->estimated_depth = f.depth; it->lidar_depth_flag = true; it->feature_per_frame[0].depth = f.depth; }
    }
  }
  if (fc < 2 || last_track_num < 20) return true;
  for (const auto& it : feature)
    if (it.start_frame <= fc - 2 && it.start_frame + (int)it.feature_per_frame.size() - 1 >= fc - 1) { parallax_sum += compensatedParallax2(it, fc); parallax_num++; }
  if (parallax_num == 0) return true;
  return parallax_sum / parallax_num >= MIN_PARALLAX;
}

void Estimator::processImage(const ImageFeatures& image, const Header& header) {
  marginalization_flag = addFeatureCheckParallax(frame_count, image, td) ? MARGIN_OLD : MARGIN_SECOND_NEW;   // :512-515
  Headers[frame_count] = header;
  if (solver_flag == INITIAL) {   // the visual-inertial bootstrap is out of scope: keep filling the window
    if (frame_count < WINDOW_SIZE) frame_count++;
    return;
  }
  if (frame_count < WINDOW_SIZE) { frame_count++; return; }
  if (TRIANGULATE) { triangulate(); if (last_status != VILS_OK) return; }   // solveOdometry (:903-914)
  optimization();
  if (last_status == VILS_OK) removeFailures();
  slideWindow();
}

// FeatureManager::triangulate (feature_manager.cpp:214-268): every feature that enters the solve and has no depth yet gets the DLT / SVD
// depth in its anchor camera, all of them in one batched device call.
void Estimator::triangulate() {
  std::vector<FeaturePerId*> todo; std::vector<int32_t> start, off{0}; std::vector<double> pts;
  for (auto& it : feature) {
    const int used_num = (int)it.feature_per_frame.size();
    if (!(used_num >= 2 && it.start_frame < WINDOW_SIZE - 2)) continue;
    if (it.estimated_depth > 0) continue;                                      // depth is available, skip (trust the first estimate)
    todo.push_back(&it); start.push_back(it.start_frame);
    for (const auto& fpf : it.feature_per_frame) { pts.push_back(fpf.point[0]); pts.push_back(fpf.point[1]); pts.push_back(fpf.point[2]); }
    off.push_back((int32_t)(pts.size() / 3));
  }
  last_status = VILS_OK;
  if (todo.empty()) return;
  const int F = WINDOW_SIZE + 1;
  auto q2R = [](const double* q, double* R) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w); R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w); R[7] = 2 * (y * z + x * w); R[8] = 1 - 2 * (x * x + y * y);
  };
  std::vector<double> P(3 * (size_t)F), R(9 * (size_t)F), depth(todo.size()); double Ric[9];
  for (int k = 0; k < F; k++) { for (int a = 0; a < 3; a++) P[3 * k + a] = Ps[k][a]; q2R(Qs[k].data(), &R[9 * k]); }
  q2R(ric, Ric);
  last_status = vils_triangulate((int)todo.size(), start.data(), off.data(), pts.data(), F, P.data(), R.data(), tic, Ric, INIT_DEPTH, depth.data(), cfg_.device);
  if (last_status != VILS_OK) return;
  for (size_t k = 0; k < todo.size(); k++) todo[k]->estimated_depth = depth[k];
}

void Estimator::removeFailures() {
  for (auto it = feature.begin(); it != feature.end();) it = (it->solve_flag == 2) ? feature.erase(it) : std::next(it);
}

// estimator.cpp:1124-1687
void Estimator::optimization() {
  const int F = WINDOW_SIZE + 1;
  // vector2double (:916-958)
  std::vector<double> pose(7 * F), sb(9 * F), ex(7);
  for (int k = 0; k < F; k++) {
    for (int i = 0; i < 3; i++) { pose[7 * k + i] = Ps[k][i]; sb[9 * k + i] = Vs[k][i]; sb[9 * k + 3 + i] = Bas[k][i]; sb[9 * k + 6 + i] = Bgs[k][i]; }
    for (int i = 0; i < 4; i++) pose[7 * k + 3 + i] = Qs[k][i];
  }
  for (int i = 0; i < 3; i++) ex[i] = tic[i];
  for (int i = 0; i < 4; i++) ex[3 + i] = ric[i];
  const double pose0_before[7] = {pose[0], pose[1], pose[2], pose[3], pose[4], pose[5], pose[6]};
  // IMU: re-integrate every interval on the GPU (IntegrationBase::repropagate semantics, integration_base.h:38-52)
  std::vector<int32_t> off(1, 0), imu_kf; std::vector<double> dt, acc, gyr, acc0, gyr0, ba, bg;
  for (int j = 1; j < F; j++) {
    const ImuBuf& b = imu_[j];
    if (b.dt.empty()) continue;
    dt.insert(dt.end(), b.dt.begin(), b.dt.end()); acc.insert(acc.end(), b.acc.begin(), b.acc.end()); gyr.insert(gyr.end(), b.gyr.begin(), b.gyr.end());
    for (int i = 0; i < 3; i++) { acc0.push_back(b.acc0[i]); gyr0.push_back(b.gyr0[i]); ba.push_back(b.ba[i]); bg.push_back(b.bg[i]); }
    off.push_back((int32_t)dt.size()); imu_kf.push_back(j - 1);
  }
  std::vector<vils_preint> pre(imu_kf.size());
  const double noise[4] = {0.02065, 0.00519, 0.00667, 0.00088056};   // ACC_N GYR_N ACC_W GYR_W (config/mynteye_leishen_indoor.yaml:81-86)
  if (!imu_kf.empty()) {
    last_status = vils_preintegrate((int)imu_kf.size(), off.data(), dt.data(), acc.data(), gyr.data(), acc0.data(), gyr0.data(), ba.data(), bg.data(), noise, pre.data(), cfg_.device);
    if (last_status != VILS_OK) return;
  }
  // projection factors (:1189-1242)
  std::vector<double> lam, pts_i, pts_j, vel_i, vel_j, td_i, td_j, row_i, row_j; std::vector<int32_t> kf_i, kf_j, feat; std::vector<uint8_t> dfix;
  std::vector<FeaturePerId*> used;
  for (auto& it : feature) {
    const int used_num = (int)it.feature_per_frame.size();
    if (!(used_num >= 2 && it.start_frame < WINDOW_SIZE - 2)) continue;
    const int fi = (int)lam.size();
    lam.push_back(it.estimated_depth > 0 ? 1.0 / it.estimated_depth : 1.0 / INIT_DEPTH);   // getDepthVector (feature_manager.cpp:195-212)
    dfix.push_back(it.lidar_depth_flag ? 1 : 0); used.push_back(&it);
    const FeaturePerFrame& f0 = it.feature_per_frame[0];
    int imu_j = it.start_frame - 1;
    for (const auto& fj : it.feature_per_frame) {
      imu_j++;
      if (imu_j == it.start_frame) continue;
      kf_i.push_back(it.start_frame); kf_j.push_back(imu_j); feat.push_back(fi);
      for (int a = 0; a < 3; a++) { pts_i.push_back(f0.point[a]); pts_j.push_back(fj.point[a]); }
      for (int a = 0; a < 2; a++) { vel_i.push_back(f0.velocity[a]); vel_j.push_back(fj.velocity[a]); }
      td_i.push_back(f0.cur_td); td_j.push_back(fj.cur_td); row_i.push_back(f0.uv[1]); row_j.push_back(fj.uv[1]);
    }
  }
  vils_window w{};
  w.n_kf = F; w.n_feat = (int)lam.size(); w.n_imu = (int)imu_kf.size(); w.n_proj = (int)kf_i.size();
  w.pose = pose.data(); w.speedbias = sb.data(); w.ex_pose = ex.data(); w.inv_depth = lam.data(); w.depth_fixed = dfix.data(); w.td = td;
  w.imu = pre.data(); w.imu_kf = imu_kf.data();
  w.pts_i = pts_i.data(); w.pts_j = pts_j.data(); w.vel_i = vel_i.data(); w.vel_j = vel_j.data(); w.td_i = td_i.data(); w.td_j = td_j.data();
  w.row_i = row_i.data(); w.row_j = row_j.data(); w.kf_i = kf_i.data(); w.kf_j = kf_j.data(); w.feat = feat.data();
  w.prior_n = prior_n_; w.prior_nblk = (int)prior_blk_.size(); w.prior_J = prior_J_.data(); w.prior_r = prior_r_.data(); w.prior_blk = prior_blk_.data(); w.prior_x0 = prior_x0_.data();
  last_n_proj = w.n_proj; last_n_feat = w.n_feat; last_prior_n = prior_n_;
  last_status = vils_ba_set_window(ba_, 0, &w);
  if (last_status != VILS_OK) return;
  last_status = vils_ba_solve(ba_, 1, &solve_opts);                      // ceres::Solve (:1400-1414)
  if (last_status != VILS_OK) return;
  double tdo = td;
  const int st = vils_ba_get_state(ba_, 0, pose.data(), sb.data(), ex.data(), lam.data(), &tdo, &last_summary);
  last_status = st;
  if (st != VILS_OK) return;                                            // state untouched: failureDetection()/reboot is the caller's job
  // double2vector (:960-1074): yaw / position re-anchoring, then unpack
  vils_double2vector(F, pose0_before, pose.data(), sb.data());
  for (int k = 0; k < F; k++) {
    for (int i = 0; i < 3; i++) { Ps[k][i] = pose[7 * k + i]; Vs[k][i] = sb[9 * k + i]; Bas[k][i] = sb[9 * k + 3 + i]; Bgs[k][i] = sb[9 * k + 6 + i]; }
    for (int i = 0; i < 4; i++) Qs[k][i] = pose[7 * k + 3 + i];
  }
  for (int i = 0; i < 3; i++) tic[i] = ex[i];
  for (int i = 0; i < 4; i++) ric[i] = ex[3 + i];
  td = tdo;
  for (size_t f = 0; f < used.size(); f++) {                            // setDepth (feature_manager.cpp:150-169)
    used[f]->estimated_depth = 1.0 / lam[f];
    used[f]->solve_flag = used[f]->estimated_depth < 0 ? 2 : 1;
  }
  // marginalization (:1483-1684): the new prior, block ids already re-addressed to the slid window
  const int cap = 15 * F + 7;
  std::vector<double> J((size_t)cap * cap), r(cap), x0((size_t)(2 * F + 2) * 9); std::vector<int32_t> blk(2 * F + 2);
  vils_prior_out po{}; po.capacity_n = cap; po.J = J.data(); po.r = r.data(); po.blk = blk.data(); po.x0 = x0.data();
  const int ms = vils_ba_marginalize(ba_, 0, marginalization_flag == MARGIN_OLD ? VILS_MARGIN_OLD : VILS_MARGIN_SECOND_NEW, &po);
  if (ms == VILS_OK && po.n > 0) {
    prior_n_ = po.n; prior_J_.assign(J.begin(), J.begin() + (size_t)po.n * po.n); prior_r_.assign(r.begin(), r.begin() + po.n);
    prior_blk_.assign(blk.begin(), blk.begin() + po.nblk);
    int gs = 0; for (int b = 0; b < po.nblk; b++) { const int t = VILS_BLK_TYPE(blk[b]); gs += (t == VILS_BLK_POSE || t == VILS_BLK_EXPOSE) ? 7 : t == VILS_BLK_SPEEDBIAS ? 9 : 1; }
    prior_x0_.assign(x0.begin(), x0.begin() + gs);
  } else if (ms == VILS_OK && marginalization_flag == MARGIN_OLD) {
    prior_n_ = 0; prior_J_.clear(); prior_r_.clear(); prior_blk_.clear(); prior_x0_.clear();
  }
}

void Estimator::removeBackShiftDepth() {
  // handled inside slideWindow (needs the marginalised and the new first camera pose)
}

void Estimator::removeFront(int fc) {
  for (auto it = feature.begin(); it != feature.end();) {
    if (it->start_frame == fc) { it->start_frame--; ++it; continue; }
    const int j = WINDOW_SIZE - 1 - it->start_frame;
    if (it->endFrame() < fc - 1) { ++it; continue; }
    it->feature_per_frame.erase(it->feature_per_frame.begin() + j);
    it = it->feature_per_frame.empty() ? feature.erase(it) : std::next(it);
  }
}

// estimator.cpp:1689-1814
void Estimator::slideWindow() {
  if (frame_count != WINDOW_SIZE) return;
  const int Wn = WINDOW_SIZE;
  if (marginalization_flag == MARGIN_OLD) {
    double back_Q0[4], back_P0[3];
    for (int i = 0; i < 4; i++) back_Q0[i] = Qs[0][i];
    for (int i = 0; i < 3; i++) back_P0[i] = Ps[0][i];
    for (int i = 0; i < Wn; i++) {
      std::swap(Qs[i], Qs[i + 1]); std::swap(imu_[i], imu_[i + 1]); Headers[i] = Headers[i + 1];
      std::swap(Ps[i], Ps[i + 1]); std::swap(Vs[i], Vs[i + 1]); std::swap(Bas[i], Bas[i + 1]); std::swap(Bgs[i], Bgs[i + 1]);
    }
    Headers[Wn] = Headers[Wn - 1]; Ps[Wn] = Ps[Wn - 1]; Vs[Wn] = Vs[Wn - 1]; Qs[Wn] = Qs[Wn - 1]; Bas[Wn] = Bas[Wn - 1]; Bgs[Wn] = Bgs[Wn - 1];
    imu_[Wn] = ImuBuf();
    // slideWindowOld + removeBackShiftDepth (feature_manager.cpp:286-344)
    double Q0c[4], Q1c[4], P0c[3], P1c[3], t0[3], t1[3];
    qmul(back_Q0, ric, Q0c); qmul(Qs[0].data(), ric, Q1c);
    qrot(back_Q0, tic, t0); qrot(Qs[0].data(), tic, t1);
    for (int i = 0; i < 3; i++) { P0c[i] = back_P0[i] + t0[i]; P1c[i] = Ps[0][i] + t1[i]; }
    for (auto it = feature.begin(); it != feature.end();) {
      if (it->start_frame != 0) { it->start_frame--; ++it; continue; }
      const FeaturePerFrame f0 = it->feature_per_frame[0];
      double depth = -1;
      if (f0.depth > 0) depth = f0.depth; else if (it->estimated_depth > 0) depth = it->estimated_depth;
      it->feature_per_frame.erase(it->feature_per_frame.begin());
      if (it->feature_per_frame.size() < 2) { it = feature.erase(it); continue; }
      const double pi[3] = {f0.point[0] * depth, f0.point[1] * depth, f0.point[2] * depth};
      double wp[3], d[3], pj[3];
      qrot(Q0c, pi, wp); for (int i = 0; i < 3; i++) d[i] = wp[i] + P0c[i] - P1c[i];
      qrot_inv(Q1c, d, pj);
      if (it->feature_per_frame[0].depth > 0) { it->estimated_depth = it->feature_per_frame[0].depth; it->lidar_depth_flag = true; }
      else if (pj[2] > 0) { it->estimated_depth = pj[2]; it->lidar_depth_flag = false; }
      else { it->estimated_depth = INIT_DEPTH; it->lidar_depth_flag = false; }
      ++it;
    }
  } else {
    ImuBuf& dst = imu_[frame_count - 1]; const ImuBuf& src = imu_[frame_count];
    dst.dt.insert(dst.dt.end(), src.dt.begin(), src.dt.end()); dst.acc.insert(dst.acc.end(), src.acc.begin(), src.acc.end()); dst.gyr.insert(dst.gyr.end(), src.gyr.begin(), src.gyr.end());
    Headers[frame_count - 1] = Headers[frame_count]; Ps[frame_count - 1] = Ps[frame_count]; Vs[frame_count - 1] = Vs[frame_count];
    Qs[frame_count - 1] = Qs[frame_count]; Bas[frame_count - 1] = Bas[frame_count]; Bgs[frame_count - 1] = Bgs[frame_count];
    imu_[Wn] = ImuBuf();
    removeFront(frame_count);
  }
}

// ---- FeatureTracker ------------------------------------------------------------------------------------------------
FeatureTracker::FeatureTracker(int rows, int cols, int max_cnt, int device) : rows_(rows), cols_(cols), max_cnt_(max_cnt) {
  last_status = vils_klt_create(rows, cols, std::max(4 * max_cnt, 64), 21, 3, device, &klt_);
  if (last_status == VILS_OK) last_status = vils_frontend_create(rows, cols, std::max(4 * max_cnt, 64), device, &fe_);
  cur_img_.resize((size_t)rows * cols);
}
FeatureTracker::~FeatureTracker() { vils_klt_destroy(klt_); vils_frontend_destroy(fe_); }
bool FeatureTracker::inBorder(float x, float y) const {
  const int B = 1; const int ix = (int)std::lround(x), iy = (int)std::lround(y);   // BORDER_SIZE = 1, cvRound
  return B <= ix && ix < cols_ - B && B <= iy && iy < rows_ - B;
}
void FeatureTracker::addPoints(const float* xy, int n) {
  for (int k = 0; k < n && (int)cur_pts.size() < max_cnt_; k++) { cur_pts.push_back({xy[2 * k], xy[2 * k + 1]}); ids.push_back(n_id_++); track_cnt.push_back(1); }
}
bool FeatureTracker::updateID(unsigned int i) {
  if (i >= ids.size()) return false;
  if (ids[i] == -1) ids[i] = n_id_++;
  return true;
}
void FeatureTracker::readImage(const uint8_t* img, int stride, double t) {
  prev_time = cur_time; cur_time = t;
  std::vector<uint8_t> forw_img((size_t)rows_ * cols_);                      // compact copy (cv::Mat::step may exceed cols)
  if (EQUALIZE) {                                                             // cv::createCLAHE(3.0, Size(8, 8))->apply (:87-93)
    last_status = vils_clahe(fe_, img, stride, 3.0, 8, 8, forw_img.data(), cols_);
    if (last_status != VILS_OK) return;
  } else {
    for (int y = 0; y < rows_; y++) std::memcpy(forw_img.data() + (size_t)y * cols_, img + (size_t)y * stride, cols_);
  }
  std::vector<std::array<float, 2>> forw_pts;
  if (has_img_ && !cur_pts.empty()) {
    const int n = (int)cur_pts.size();
    std::vector<std::array<float, 2>> forw(n); std::vector<uint8_t> status(n); std::vector<float> err(n);
    last_status = vils_klt_track(klt_, cur_img_.data(), forw_img.data(), cols_, &cur_pts[0][0], n, &forw[0][0], status.data(), err.data());   // :113
    if (last_status != VILS_OK) return;
    for (int i = 0; i < n; i++) if (status[i] && !inBorder(forw[i][0], forw[i][1])) status[i] = 0;        // :115-117
    prev_pts = cur_pts;
    size_t j = 0;                                                                                          // reduceVector (:21-34)
    for (int i = 0; i < n; i++) if (status[i]) { prev_pts[j] = prev_pts[i]; forw[j] = forw[i]; ids[j] = ids[i]; track_cnt[j] = track_cnt[i] + 1; j++; }   // + n++ (:127-128)
    prev_pts.resize(j); forw.resize(j); ids.resize(j); track_cnt.resize(j);
    forw_pts = forw;
  } else {
    for (auto& c : track_cnt) c++;                                            // :127-128 also runs when nothing was tracked
    forw_pts = cur_pts;
  }
  if (PUB_THIS_FRAME) {
    if (REJECT_WITH_F && has_img_) { rejectWithF(forw_pts); if (last_status != VILS_OK) return; }   // :131
    // setMask (:36-69): long-tracked points first, MIN_DIST discs
    const int n = (int)forw_pts.size();
    std::vector<int32_t> keep(std::max(n, 1)), cnt32(track_cnt.begin(), track_cnt.end()); int32_t nk = 0;
    last_status = vils_set_mask(fe_, n ? &forw_pts[0][0] : nullptr, cnt32.data(), n, MIN_DIST, keep.data(), &nk);
    if (last_status != VILS_OK) return;
    std::vector<std::array<float, 2>> fp(nk); std::vector<int> id2(nk), tc2(nk);
    for (int k = 0; k < nk; k++) { fp[k] = forw_pts[keep[k]]; id2[k] = ids[keep[k]]; tc2[k] = track_cnt[keep[k]]; }
    forw_pts.swap(fp); ids.swap(id2); track_cnt.swap(tc2);
    const int n_max_cnt = max_cnt_ - (int)forw_pts.size();                    // :139-151
    if (n_max_cnt > 0) {
      std::vector<float> n_pts(2 * (size_t)n_max_cnt); int32_t nn = 0;
      last_status = vils_good_features(fe_, forw_img.data(), cols_, n_max_cnt, 0.01, (double)MIN_DIST, 1, n_pts.data(), &nn);
      if (last_status != VILS_OK) return;
      for (int k = 0; k < nn; k++) { forw_pts.push_back({n_pts[2 * k], n_pts[2 * k + 1]}); ids.push_back(-1); track_cnt.push_back(1); }   // addPoints (:71-79)
    }
  }
  cur_pts = forw_pts;
  cur_img_.swap(forw_img);                                                   // prev = cur, cur = forw (:160-164)
  has_img_ = true;
  undistortedPoints();
}
void FeatureTracker::rejectWithF(std::vector<std::array<float, 2>>& forw_pts) {
  const int n = (int)forw_pts.size();
  if (n < 8 || (int)prev_pts.size() != n) return;                            // :171 (prev_pts here = the reference's cur_pts after reduceVector)
  std::vector<double> r1(3 * (size_t)n), r2(3 * (size_t)n);
  last_status = vils_lift_projective(fe_, cam, &prev_pts[0][0], n, r1.data());
  if (last_status == VILS_OK) last_status = vils_lift_projective(fe_, cam, &forw_pts[0][0], n, r2.data());
  if (last_status != VILS_OK) return;
  std::vector<std::array<float, 2>> u1(n), u2(n);
  for (int i = 0; i < n; i++) {                                               // :176-188 virtual pinhole with FOCAL_LENGTH
    u1[i] = {(float)(FOCAL_LENGTH * r1[3 * i] / r1[3 * i + 2] + cols_ / 2.0), (float)(FOCAL_LENGTH * r1[3 * i + 1] / r1[3 * i + 2] + rows_ / 2.0)};
    u2[i] = {(float)(FOCAL_LENGTH * r2[3 * i] / r2[3 * i + 2] + cols_ / 2.0), (float)(FOCAL_LENGTH * r2[3 * i + 1] / r2[3 * i + 2] + rows_ / 2.0)};
  }
  std::vector<uint8_t> status(n);
  last_status = vils_reject_with_f(fe_, &u1[0][0], &u2[0][0], n, F_THRESHOLD, status.data(), nullptr);
  if (last_status != VILS_OK) return;
  size_t j = 0;                                                               // reduceVector x6 (:193-198)
  for (int i = 0; i < n; i++) if (status[i]) { prev_pts[j] = prev_pts[i]; forw_pts[j] = forw_pts[i]; ids[j] = ids[i]; track_cnt[j] = track_cnt[i]; j++; }
  prev_pts.resize(j); forw_pts.resize(j); ids.resize(j); track_cnt.resize(j);
}
void FeatureTracker::undistortedPoints() {
  const int n = (int)cur_pts.size();
  cur_un_pts.assign(n, {0.f, 0.f}); pts_velocity.assign(n, {0.f, 0.f});
  std::map<int, std::array<float, 2>> cur_map;
  if (n) {
    std::vector<double> rays(3 * (size_t)n);
    last_status = vils_lift_projective(fe_, cam, &cur_pts[0][0], n, rays.data());   // m_camera->liftProjective (:266-268)
    if (last_status != VILS_OK) return;
    for (int i = 0; i < n; i++) {
      cur_un_pts[i] = {(float)(rays[3 * i] / rays[3 * i + 2]), (float)(rays[3 * i + 1] / rays[3 * i + 2])};
      cur_map.insert({ids[i], cur_un_pts[i]});                                 // std::map::insert keeps the FIRST entry of a duplicated id (-1)
    }
  }
  if (!prev_un_pts_map_.empty()) {                                            // :274-299
    const double dt = cur_time - prev_time;
    for (int i = 0; i < n; i++) {
      if (ids[i] == -1) continue;
      auto it = prev_un_pts_map_.find(ids[i]);
      if (it != prev_un_pts_map_.end()) pts_velocity[i] = {(float)((cur_un_pts[i][0] - it->second[0]) / dt), (float)((cur_un_pts[i][1] - it->second[1]) / dt)};
    }
  }
  prev_un_pts_map_ = cur_map;
}

// ---- FastVGICP (estimator.cpp:269-297) --------------------------------------------------------------------------------------------
FastVGICP::FastVGICP(int device) : device_(device) {
  vils_vgicp_default_opts(&opts_);
  std::memset(&res_, 0, sizeof(res_)); res_.fitness = -1.0;
  for (int k = 0; k < 16; k++) final_[k] = (k % 5 == 0) ? 1.0f : 0.0f;
}
static void pack_xyzi(std::vector<float>& dst, const float* xyzi, int n, int stride) {
  dst.resize((size_t)4 * std::max(n, 0));
  for (int i = 0; i < n; i++) { const float* p = xyzi + (size_t)stride * i; dst[4 * i] = p[0]; dst[4 * i + 1] = p[1]; dst[4 * i + 2] = p[2]; dst[4 * i + 3] = stride > 4 ? p[4] : (stride > 3 ? p[3] : 0.0f); }
}
void FastVGICP::setInputSource(const float* xyzi, int n, int stride) { pack_xyzi(src_, xyzi, n, stride); }
void FastVGICP::setInputTarget(const float* xyzi, int n, int stride) { pack_xyzi(tgt_, xyzi, n, stride); }
int FastVGICP::align(float* aligned, const float* guess) {
  double g[16];
  if (guess) for (int k = 0; k < 16; k++) g[k] = (double)guess[k];     // Eigen::Isometry3d(guess.cast<double>()), lsq_registration_impl.hpp:53
  last_status = vils_vgicp_align(src_.data(), (int)(src_.size() / 4), tgt_.data(), (int)(tgt_.size() / 4), guess ? g : nullptr, &opts_, &res_, device_);
  if (last_status != VILS_OK) return last_status;
  for (int k = 0; k < 16; k++) final_[k] = (float)res_.T[k];            // final_transformation_ = x0.cast<float>().matrix()
  if (aligned) {
    const int n = (int)(src_.size() / 4);
    for (int i = 0; i < n; i++) {
      const float x = src_[4 * i], y = src_[4 * i + 1], z = src_[4 * i + 2];
      for (int r = 0; r < 3; r++) aligned[4 * i + r] = ((final_[4 * r] * x + final_[4 * r + 1] * y) + final_[4 * r + 2] * z) + final_[4 * r + 3];
      aligned[4 * i + 3] = src_[4 * i + 3];
    }
  }
  return VILS_OK;
}

int TransformToEnd(float* xyzi, int n, const float q[4], const float t[3], float time_factor, double min_r, double max_r, int device) {
  return vils_deskew(xyzi, n, 8, q, t, time_factor, (float)min_r, (float)max_r, device);
}

}  // namespace vils

// ---- C wrapper for driving the C++ host classes from tests (ctypes) ---------------------------------------------------
extern "C" {
void* vh_estimator_create(const vils_config* cfg, int window_size, int iters, int mode, double mu) {
  auto* e = new vils::Estimator(*cfg, window_size, iters);
  e->solve_opts.mode = mode; e->solve_opts.mu = mu;
  return e;
}
void vh_estimator_destroy(void* p) { delete static_cast<vils::Estimator*>(p); }
void vh_set_parameter(void* p, const double* ric9, const double* tic3, double td) { static_cast<vils::Estimator*>(p)->setParameter(ric9, tic3, td); }
void vh_set_frame_state(void* p, int k, const double* P, const double* Q, const double* V, const double* Ba, const double* Bg) { static_cast<vils::Estimator*>(p)->setFrameState(k, P, Q, V, Ba, Bg); }
void vh_process_imu(void* p, double dt, const double* acc, const double* gyr) { static_cast<vils::Estimator*>(p)->processIMU(dt, acc, gyr); }
void vh_process_image(void* p, int n, const int* ids, const double* feat8, double stamp) {
  vils::ImageFeatures im;
  for (int k = 0; k < n; k++) { vils::Feature8 f; for (int a = 0; a < 8; a++) f[a] = feat8[8 * k + a]; im[ids[k]].emplace_back(0, f); }
  vils::Header h; h.stamp = stamp;
  static_cast<vils::Estimator*>(p)->processImage(im, h);
}
int vh_frame_count(void* p) { return static_cast<vils::Estimator*>(p)->frame_count; }
int vh_last_status(void* p) { return static_cast<vils::Estimator*>(p)->last_status; }
void vh_get_frame(void* p, int k, double* P, double* Q, double* V, double* Ba, double* Bg) {
  auto* e = static_cast<vils::Estimator*>(p);
  for (int i = 0; i < 3; i++) { P[i] = e->Ps[k][i]; V[i] = e->Vs[k][i]; Ba[i] = e->Bas[k][i]; Bg[i] = e->Bgs[k][i]; }
  for (int i = 0; i < 4; i++) Q[i] = e->Qs[k][i];
}
void vh_get_info(void* p, double* out /* cost0 cost1 iters n_proj n_feat prior_n td margin_flag n_features */) {
  auto* e = static_cast<vils::Estimator*>(p);
  out[0] = e->last_summary.cost_initial; out[1] = e->last_summary.cost_final; out[2] = e->last_summary.iterations; out[3] = e->last_n_proj;
  out[4] = e->last_n_feat; out[5] = e->last_prior_n; out[6] = e->td; out[7] = e->marginalization_flag; out[8] = (double)e->feature.size();
}
void* vh_tracker_create(int rows, int cols, int max_cnt, int device) { return new vils::FeatureTracker(rows, cols, max_cnt, device); }
void vh_tracker_destroy(void* p) { delete static_cast<vils::FeatureTracker*>(p); }
void vh_tracker_add(void* p, const float* xy, int n) { static_cast<vils::FeatureTracker*>(p)->addPoints(xy, n); }
int vh_tracker_read(void* p, const uint8_t* img, int stride, double t) { auto* f = static_cast<vils::FeatureTracker*>(p); f->readImage(img, stride, t); return f->last_status; }
int vh_tracker_get(void* p, float* xy, int* ids, int* cnt, int cap) {
  auto* f = static_cast<vils::FeatureTracker*>(p); const int n = std::min<int>(cap, (int)f->cur_pts.size());
  for (int k = 0; k < n; k++) { xy[2 * k] = f->cur_pts[k][0]; xy[2 * k + 1] = f->cur_pts[k][1]; ids[k] = f->ids[k]; cnt[k] = f->track_cnt[k]; }
  return n;
}
void vh_tracker_config(void* p, int equalize, int pub, int min_dist, const double* cam) {
  auto* f = static_cast<vils::FeatureTracker*>(p); f->EQUALIZE = equalize != 0; f->PUB_THIS_FRAME = (pub & 1) != 0; f->REJECT_WITH_F = (pub & 2) != 0; f->MIN_DIST = min_dist;
  if (cam) for (int k = 0; k < 8; k++) f->cam[k] = cam[k];
}
void vh_tracker_update_ids(void* p) { auto* f = static_cast<vils::FeatureTracker*>(p); for (unsigned int i = 0; f->updateID(i); i++) {} }   // feature_tracker_node.cpp:120-128
int vh_tracker_get_un(void* p, float* un_xy, float* vel, int cap) {
  auto* f = static_cast<vils::FeatureTracker*>(p); const int n = std::min<int>(cap, (int)f->cur_un_pts.size());
  for (int k = 0; k < n; k++) { un_xy[2 * k] = f->cur_un_pts[k][0]; un_xy[2 * k + 1] = f->cur_un_pts[k][1]; vel[2 * k] = f->pts_velocity[k][0]; vel[2 * k + 1] = f->pts_velocity[k][1]; }
  return n;
}
int vh_transform_to_end(float* xyzi, int n, const float* q, const float* t, float tf, double mn, double mx) { return vils::TransformToEnd(xyzi, n, q, t, tf, mn, mx, 0); }
// FastVGICP through a flat interface (tests): returns the status; out = T(16 floats) fitness converged iterations
int vh_vgicp_align(const float* src, int n_src, const float* tgt, int n_tgt, int stride, double resolution, const float* guess, float* T16, double* info3) {
  vils::FastVGICP g; g.setResolution(resolution); g.setNumThreads(4);
  g.setInputSource(src, n_src, stride); g.setInputTarget(tgt, n_tgt, stride);
  const int st = g.align(nullptr, guess);
  if (st != VILS_OK) return st;
  for (int k = 0; k < 16; k++) T16[k] = g.getFinalTransformation()[k];
  info3[0] = g.getFitnessScore(); info3[1] = g.hasConverged() ? 1.0 : 0.0; info3[2] = g.last_result().iterations;
  return VILS_OK;
}
}
