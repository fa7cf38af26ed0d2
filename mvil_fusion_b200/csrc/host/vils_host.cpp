// vils_host.cpp — see vils_host.h.  Host glue only: bookkeeping, packing the window for the C-ABI, state shuffling.
// Every number-crunching step (pre-integration, factor evaluation, solve, marginalization, KLT, deskew) is a library call.
#include "vils_host.h"
#include "vils_hostmath.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace vils {
using namespace hm;
namespace {

// pcl::ApproximateVoxelGrid<PointXYZI> [upstream PCL 1.8 filters/impl/approximate_voxel_grid.hpp, restated from memory; PCL is not in the
// reference tree]: a 512-entry hash history keyed by (ix * 7171 + iy * 3079 + iz * 4231) & 511 over floor(p / leaf); a colliding voxel flushes the
// resident centroid (x, y, z, intensity averaged in float, downsample_all_data = true) to the output; everything left is flushed at the end.
void approximate_voxel_grid(const float* xyzi, int n, int stride, float leaf, std::vector<float>& out) {
  struct He { int ix = 0, iy = 0, iz = 0, count = 0; float c[4] = {0, 0, 0, 0}; };
  const int hist = 512; std::vector<He> h(hist);
  out.clear(); out.reserve((size_t)4 * n);
  const float inv = 1.0f / leaf;
  auto flush = [&](He& e) { const float k = (float)e.count; for (int a = 0; a < 4; a++) out.push_back(e.c[a] / k); };
  for (int i = 0; i < n; i++) {
    const float* p = xyzi + (size_t)stride * i; const float inten = stride > 4 ? p[4] : p[3];
    const int ix = (int)std::floor(p[0] * inv), iy = (int)std::floor(p[1] * inv), iz = (int)std::floor(p[2] * inv);
    He& e = h[(unsigned int)(ix * 7171 + iy * 3079 + iz * 4231) & (hist - 1)];
    if (e.count && (ix != e.ix || iy != e.iy || iz != e.iz)) { flush(e); e.count = 0; e.c[0] = e.c[1] = e.c[2] = e.c[3] = 0; }
    e.ix = ix; e.iy = iy; e.iz = iz; e.count++;
    e.c[0] += p[0]; e.c[1] += p[1]; e.c[2] += p[2]; e.c[3] += inten;
  }
  for (auto& e : h) if (e.count) flush(e);
}

}  // namespace

Estimator::Estimator(const vils_config& cfg, int window_size, int num_iterations) : WINDOW_SIZE(window_size), cfg_(cfg) {
  const int F = WINDOW_SIZE + 1;
  cfg_.max_kf = F;
  Ps.assign(F, {0, 0, 0}); Vs.assign(F, {0, 0, 0}); Bas.assign(F, {0, 0, 0}); Bgs.assign(F, {0, 0, 0}); Qs.assign(F, {0, 0, 0, 1});
  Headers.assign(F, Header()); imu_.assign(F, ImuBuf());
  for (int i = 0; i < 3; i++) g[i] = cfg.gravity[i];
  vils_default_solve_opts(&solve_opts);
  // ceres DENSE_SCHUR + DOGLEG, max_num_iterations = NUM_ITERATIONS, max_solver_time_in_seconds = SOLVER_TIME (estimator.cpp:1402-1411)
  solve_opts.mode = VILS_MODE_DOGLEG; solve_opts.max_iters = num_iterations; solve_opts.max_solver_time = SOLVER_TIME;
  for (int i = 0; i < 9; i++) RLB[i] = cfg.rlb[i];
  for (int i = 0; i < 3; i++) TLB[i] = cfg.tlb[i];
  { double Rt[9]; mat3_T(RLB, Rt); mat3_vec(Rt, TLB, TBL); for (int i = 0; i < 3; i++) TBL[i] = -TBL[i]; }   // TBL = -RLB^T TLB (estimator.cpp:451)
  last_status = vils_ba_create(&cfg_, 1, &ba_);
}
Estimator::~Estimator() { vils_ba_destroy(ba_); vils_frontend_destroy(init_fe_); }

void Estimator::setParameter(const double r[9], const double t[3], double td_) {
  R2q(r, ric); qnorm(ric); for (int i = 0; i < 3; i++) tic[i] = t[i]; td = td_;
  if (r != ric_cfg_) { std::memcpy(ric_cfg_, r, sizeof(ric_cfg_)); std::memcpy(tic_cfg_, t, sizeof(tic_cfg_)); td_cfg_ = td_; }
}

void Estimator::clearState() {
  const int F = WINDOW_SIZE + 1;
  Ps.assign(F, {0, 0, 0}); Vs.assign(F, {0, 0, 0}); Bas.assign(F, {0, 0, 0}); Bgs.assign(F, {0, 0, 0}); Qs.assign(F, {0, 0, 0, 1});
  imu_.assign(F, ImuBuf()); feature.clear(); frame_count = 0; first_imu_ = false; solver_flag = INITIAL; prior_n_ = 0;
  prior_J_.clear(); prior_r_.clear(); prior_x0_.clear(); prior_blk_.clear();
  all_image_frame.clear(); tmp_pre_ = ImuBuf(); initial_timestamp = 0; td = 0;
  // the LiDAR side of clearState (estimator.cpp:66-83): queues, frames, counters
  LidarICPConstraints.clear(); LidarLPSConstraints.clear(); all_lidar_frame.clear(); lidar_count = 0; lidar_count_ = 0; first_zv_ = true; LPS_call_ = false;
  lp_plane_.clear(); lp_edge_.clear(); lp_plane_kf_.clear(); lp_edge_kf_.clear();
}

void Estimator::setLPS(const double q[4], const double t[3], double time) {
  if (!ADD_LPS) return;                                     // estimator_node.cpp:560
  for (int i = 0; i < 4; i++) LPS_q_[i] = q[i];
  for (int i = 0; i < 3; i++) LPS_t_[i] = t[i];
  LPS_time_ = time; LPS_call_ = true;
}

void Estimator::setLidarPointFactors(int n_plane, const double* pl, const int32_t* pkf, int n_edge, const double* ed, const int32_t* ekf) {
  lp_plane_.assign(pl, pl + (size_t)7 * std::max(n_plane, 0)); lp_plane_kf_.assign(pkf, pkf + std::max(n_plane, 0));
  lp_edge_.assign(ed, ed + (size_t)9 * std::max(n_edge, 0)); lp_edge_kf_.assign(ekf, ekf + std::max(n_edge, 0));
}

// lidar_backend.cpp:3-36: the two window frames that bracket time tl
bool Estimator::FindNearest2ID(double tl, int& id_a, int& id_b) const {
  std::vector<double> header;
  for (int i = 0; i < WINDOW_SIZE + 1; i++) header.push_back(Headers[i].stamp);
  header.push_back(tl);
  std::sort(header.begin(), header.end());
  auto it = std::find(header.begin(), header.end(), tl);
  if (it == header.end()) return false;
  const int index = (int)(it - header.begin());
  id_a = index - 1; id_b = index;
  return !(id_b > WINDOW_SIZE || id_a < 0);
}

// lidar_backend.cpp:38-97, including its fall-throughs (ids that are not found keep their initial 0, b == c shifts the first pair back)
bool Estimator::FindWindowsID(double ta, double tb, double tc, double td_, int& id_a, int& id_b, int& id_c, int& id_d) const {
  if (Headers[0].stamp > ta || Headers[WINDOW_SIZE].stamp < td_ || (tb - ta) > 0.5) return false;
  auto find = [&](double t, int& id) { for (int i = 0; i < WINDOW_SIZE + 1; i++) if (Headers[i].stamp == t) { id = i; return; } };
  find(ta, id_a); find(tb, id_b); find(tc, id_c); find(td_, id_d);
  if (id_b == id_c) { id_a--; id_b--; }
  return id_b > id_a && id_d > id_c && id_a >= 0 && id_a != id_c;
}

namespace {
// lidar_frontend.cpp:941-987
void Predict_r(const LidarFrame& f, double R[9]) {
  const double t = (f.time - f.vioData.ti) / (f.vioData.tj - f.vioData.ti);
  double q[4];
  if (t > 0) qslerp(f.vioData.Qwbi, f.vioData.Qwbj, t, q); else std::memcpy(q, f.vioData.Qwbi, sizeof(q));
  q2R(q, R);
}
void Predict_t(const LidarFrame& f, double P[3]) {
  const double dt = f.time - f.vioData.ti;
  for (int i = 0; i < 3; i++) {
    if (dt >= 0) { const double a = (f.vioData.Vbj[i] - f.vioData.Vbi[i]) / (f.vioData.tj - f.vioData.ti); P[i] = f.vioData.Pwbi[i] + f.vioData.Vbi[i] * dt + 0.5 * a * dt * dt; }
    else P[i] = f.vioData.Pwbi[i];
  }
}
// lidar_frontend.cpp:921-939: relative body motion between two LiDAR stamps predicted from the VIO states; also leaves the predicted LiDAR poses in the frames
void PredictRelative_rt(LidarFrame& fi, LidarFrame& fj, const double RBL[9], const double TBL[3], double Lij[16]) {
  double Rbi[9], Rbj[9], Tbi[3], Tbj[3], RbiT[9], Rij[9], d[3], Tij[3];
  Predict_r(fj, Rbj); Predict_r(fi, Rbi); Predict_t(fj, Tbj); Predict_t(fi, Tbi);
  mat3_T(Rbi, RbiT); mat3_mul(RbiT, Rbj, Rij);
  for (int i = 0; i < 3; i++) d[i] = Tbj[i] - Tbi[i];
  mat3_vec(RbiT, d, Tij);
  rigid4(Rij, Tij, Lij);
  double o[3];
  mat3_mul(Rbi, RBL, fi.lidar_R); mat3_mul(Rbj, RBL, fj.lidar_R);
  mat3_vec(Rbi, TBL, o); for (int i = 0; i < 3; i++) fi.lidar_T[i] = Tbi[i] + o[i];
  mat3_vec(Rbj, TBL, o); for (int i = 0; i < 3; i++) fj.lidar_T[i] = Tbj[i] + o[i];
}
}  // namespace

// estimator.cpp:122-504
void Estimator::processLidar(float* xyzi, int n, int stride, double cloud_time, double time_) {
  current_lidar = LidarFrame(); current_lidar.frameID = lidar_count_; current_lidar.time = cloud_time; current_lidar_points = n;
  last_status = VILS_OK;
  const double tm = cloud_time;
  int idl = 0, idr = 0;
  if (solver_flag != INITIAL && n > 0 && FindNearest2ID(tm, idl, idr)) {
    auto iterj = all_image_frame.find(Headers[idr].stamp), iteri = all_image_frame.find(Headers[idl].stamp);
    current_lidar.keylidar = iterj != all_image_frame.end() && iteri != all_image_frame.end();
    if (current_lidar.keylidar) {
      VIOData& v = current_lidar.vioData;
      current_lidar.next_image_t = iterj->second.t; v.tj = iterj->first + time_;
      current_lidar.last_image_t = iteri->second.t; v.ti = iteri->first + time_; v.dt = time_;
      for (int i = 0; i < 4; i++) { v.Qwbi[i] = Qs[idl][i]; v.Qwbj[i] = Qs[idr][i]; }
      for (int i = 0; i < 3; i++) { v.Pwbi[i] = Ps[idl][i]; v.Vbi[i] = Vs[idl][i]; v.Pwbj[i] = Ps[idr][i]; v.Vbj[i] = Vs[idr][i]; }
      double trans_lb[16], trans_lb_inv[16];
      rigid4(RLB, TLB, trans_lb); rigid4_inv(trans_lb, trans_lb_inv);
      // step 1: distortion adjust (only once the LiDAR extrinsic is in place), VERS2 of :190-237
      if (!lidar_init_flag) {
        const float time_factor = (float)(1.0 / LidarTimeStep);
        const double ta = v.ti, tb = v.tj, tls = tm - 0.5 * LidarTimeStep, tle = tm + 0.5 * LidarTimeStep;   // LiDAR stamp sits in the middle of the sweep
        const double ss = (tls - ta) / (tb - ta), se = (tle - ta) / (tb - ta);
        double qls[4], qle[4], qlei[4], qr[4], Rb[9], Rwbj[9], RwbjT[9], dP[3], tb3[3];
        qslerp(v.Qwbi, v.Qwbj, ss, qls); qslerp(v.Qwbi, v.Qwbj, se, qle);
        qinv(qle, qlei); qmul(qlei, qls, qr); q2R(qr, Rb);
        q2R(v.Qwbj, Rwbj); mat3_T(Rwbj, RwbjT);
        for (int i = 0; i < 3; i++) dP[i] = v.Pwbi[i] - v.Pwbj[i];
        mat3_vec(RwbjT, dP, tb3); for (int i = 0; i < 3; i++) tb3[i] *= LidarTimeStep / (tb - ta);
        double trans_b[16], tmp[16], trans_l[16];
        rigid4(Rb, tb3, trans_b); mat4_mul(trans_lb, trans_b, tmp); mat4_mul(tmp, trans_lb_inv, trans_l);
        float Rf[9]; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Rf[3 * i + j] = (float)trans_l[4 * i + j];
        double Rd[9], qd[4]; for (int i = 0; i < 9; i++) Rd[i] = Rf[i];
        R2q(Rd, qd);                                            // Eigen::Quaternionf(Matrix3f): same branch structure, evaluated here in double on the float entries
        const float qf[4] = {(float)qd[0], (float)qd[1], (float)qd[2], (float)qd[3]};
        const float tf[3] = {(float)trans_l[3], (float)trans_l[7], (float)trans_l[11]};
        last_status = vils_deskew(xyzi, n, stride, qf, tf, time_factor, (float)MinDistance, (float)MaxDistance, cfg_.device);   // TransformToEnd (:233)
        if (last_status != VILS_OK) return;
        int kept = 0;                                           // pcl::removeNaNFromPointCloud (:236)
        for (int i = 0; i < n; i++) {
          const float* p = xyzi + (size_t)stride * i;
          if (std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2])) { if (kept != i) std::memmove(xyzi + (size_t)stride * kept, p, sizeof(float) * stride); kept++; }
        }
        n = kept; current_lidar_points = n;
      }
      // step 2: downsampling (:241-247)
      approximate_voxel_grid(xyzi, n, stride, (float)LeafSize, current_lidar.cloud);
      // step 3: fast-gicp against the previous key LiDAR frame
      all_lidar_frame[tm] = current_lidar;
      if (all_lidar_frame.size() > 1) {
        if (all_lidar_frame.size() > 2) all_lidar_frame.erase(all_lidar_frame.begin());
        auto lframej = std::prev(all_lidar_frame.end()); auto lframei = std::prev(lframej);
        FastVGICP gicp(cfg_.device);
        gicp.setResolution(0.5);
        gicp.setInputSource(lframej->second.cloud.data(), (int)(lframej->second.cloud.size() / 4), 4);
        gicp.setInputTarget(lframei->second.cloud.data(), (int)(lframei->second.cloud.size() / 4), 4);
        double init_guss[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        if (!lidar_init_flag) {
          double RBL[9], Lij[16], tem_r[9], tem_t[3], A[9], o1[3], o2[3];
          mat3_T(RLB, RBL);
          PredictRelative_rt(lframei->second, lframej->second, RBL, TBL, Lij);
          for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) tem_r[3 * i + j] = Lij[4 * i + j]; tem_t[i] = Lij[4 * i + 3]; }
          mat3_mul(RLB, tem_r, A);                               // init_guss = [RLB tem_r RLB^T | RLB tem_r TBL + TLB + RLB tem_t] (:283-285)
          double Rg[9]; mat3_mul(A, RBL, Rg);
          mat3_vec(A, TBL, o1); mat3_vec(RLB, tem_t, o2);
          double tg[3]; for (int i = 0; i < 3; i++) tg[i] = o1[i] + TLB[i] + o2[i];
          rigid4(Rg, tg, init_guss);
          float guss[16]; for (int i = 0; i < 16; i++) guss[i] = (float)init_guss[i];
          last_status = gicp.align(nullptr, guss);
          std::memcpy(current_lidar.lidar_R, lframej->second.lidar_R, sizeof(current_lidar.lidar_R));
          std::memcpy(current_lidar.lidar_T, lframej->second.lidar_T, sizeof(current_lidar.lidar_T));
        } else last_status = gicp.align(nullptr, nullptr);
        if (last_status != VILS_OK) return;
        const double fitness_core = gicp.getFitnessScore();     // step 4
        double T[16]; for (int i = 0; i < 16; i++) T[i] = (double)gicp.getFinalTransformation()[i];
        const double Tij[3] = {T[3], T[7], T[11]};
        // step 7: classify and queue the LiDAR constraint (:322-436)
        if (!lidar_init_flag) {
          const double tem_T = std::fabs(init_guss[3] - Tij[0]) + std::fabs(init_guss[7] - Tij[1]) + std::fabs(init_guss[11] - Tij[2]);
          LidarICPConstraint c; c.constraint_mode = 0;
          double Rg[9]; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Rg[3 * i + j] = init_guss[4 * i + j];
          const double temYaw = yaw_deg(Rg);
          if (fitness_core < 1.0 && tem_T > 0.1) c.constraint_mode = 3;
          else if (fitness_core < 1.0 && tem_T <= 0.1) c.constraint_mode = 2;
          else if (fitness_core > 1.0) c.constraint_mode = 1;
          if (std::fabs(Tij[0]) + std::fabs(Tij[1]) + std::fabs(Tij[2]) < 0.01) c.constraint_mode = std::fabs(temYaw) < 0.5 ? 4 : 5;   // zero velocity / pure rotation
          current_lidar.mode = c.constraint_mode;
          if (!ADD_LIDAR_ICP) c.constraint_mode = 0;
          c.lidar_ta = lframei->second.last_image_t; c.lidar_tb = lframei->second.next_image_t;
          c.lidar_tc = lframej->second.last_image_t; c.lidar_td = lframej->second.next_image_t;
          c.lidar_ti = lframei->first; c.lidar_tj = lframej->first;
          if (c.constraint_mode == 4) {
            c.lidar_sqrt_info00 = 1e12;                         // lidar_trans stays the identity
            if (first_zv_) {
              std::memcpy(tem_zv_r_, lframei->second.lidar_R, sizeof(tem_zv_r_)); std::memcpy(tem_zv_t_, lframei->second.lidar_T, sizeof(tem_zv_t_));
              first_zv_ = false;
              while (LidarICPConstraints.size() > 1) LidarICPConstraints.pop_front();
            }
            std::memcpy(current_lidar.lidar_R, tem_zv_r_, sizeof(tem_zv_r_)); std::memcpy(current_lidar.lidar_T, tem_zv_t_, sizeof(tem_zv_t_));
          } else if (c.constraint_mode == 3) {
            double tmp[16]; mat4_mul(trans_lb_inv, T, tmp); mat4_mul(tmp, trans_lb, c.lidar_trans);   // EX_LB^-1 T EX_LB (:415)
            c.lidar_sqrt_info00 = 1.0 / fitness_core * 100.0;   // :417
            if (!first_zv_ && LidarICPConstraints.size() == 1) { LidarICPConstraints.pop_front(); first_zv_ = true; }   // start moving again
          }
          LidarICPConstraints.push_back(c);
        }
        // step 8: LiDAR-IMU initialisation: the reference adopts the configured ("gt") extrinsic after 15 key LiDAR frames (:439-497, USE_ES off)
        if (solver_flag != INITIAL && lidar_init_flag && current_lidar.frameID > 15) {
          for (int i = 0; i < 9; i++) RLB[i] = cfg_.rlb[i];
          for (int i = 0; i < 3; i++) TLB[i] = cfg_.tlb[i];
          { double Rt[9]; mat3_T(RLB, Rt); mat3_vec(Rt, TLB, TBL); for (int i = 0; i < 3; i++) TBL[i] = -TBL[i]; }
          lidar_init_flag = false;
          double Ri[9], ti[3], RBL[9], o[3];
          Predict_r(current_lidar, Ri); Predict_t(current_lidar, ti); mat3_T(RLB, RBL);
          mat3_vec(Ri, TBL, o); for (int i = 0; i < 3; i++) current_lidar.lidar_T[i] = ti[i] + o[i];
          mat3_mul(Ri, RBL, current_lidar.lidar_R);
        }
      }
      lidar_count_++;
    }
  }
  lidar_count++;
}

// estimator.cpp:1076-1122 (the commented-out early returns stay out)
bool Estimator::failureDetection() {
  const int Wn = WINDOW_SIZE;
  auto norm3 = [](const std::array<double, 3>& v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); };
  if (norm3(Bas[Wn]) > 2.5) return true;
  if (norm3(Bgs[Wn]) > 1.0) return true;
  const double d[3] = {Ps[Wn][0] - last_P[0], Ps[Wn][1] - last_P[1], Ps[Wn][2] - last_P[2]};
  if (std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) > 10) return true;
  if (std::fabs(d[2]) > 1) return true;
  return false;
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
    if (!tmp_pre_.started) { tmp_pre_.started = true; std::memcpy(tmp_pre_.acc0, acc_0_, 24); std::memcpy(tmp_pre_.gyr0, gyr_0_, 24); for (int i = 0; i < 3; i++) { tmp_pre_.ba[i] = Bas[frame_count][i]; tmp_pre_.bg[i] = Bgs[frame_count][i]; } }
    tmp_pre_.dt.push_back(dt); for (int i = 0; i < 3; i++) { tmp_pre_.acc.push_back(acc[i]); tmp_pre_.gyr.push_back(gyr[i]); }   // tmp_pre_integration->push_back (:104)
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
  { ImageFrame fr; fr.t = header.stamp; fr.points = image; fr.pre = tmp_pre_; all_image_frame[header.stamp] = fr; }   // :523-525
  tmp_pre_ = ImuBuf();                                        // tmp_pre_integration = new IntegrationBase{acc_0, gyr_0, Bas[frame_count], Bgs[frame_count]}
  tmp_pre_.started = true; std::memcpy(tmp_pre_.acc0, acc_0_, 24); std::memcpy(tmp_pre_.gyr0, gyr_0_, 24);
  for (int i = 0; i < 3; i++) { tmp_pre_.ba[i] = Bas[frame_count][i]; tmp_pre_.bg[i] = Bgs[frame_count][i]; }
  auto remember = [&] { for (int i = 0; i < 3; i++) { last_P[i] = Ps[WINDOW_SIZE][i]; last_P0[i] = Ps[0][i]; } for (int i = 0; i < 4; i++) { last_Q[i] = Qs[WINDOW_SIZE][i]; last_Q0[i] = Qs[0][i]; } };
  if (frame_count < WINDOW_SIZE) { frame_count++; return; }  // the window is still filling (:578-579; also the setFrameState bootstrap)
  if (solver_flag == INITIAL) {                               // :553-580
    bool result = false;
    if (header.stamp - initial_timestamp > 0.1) { result = initialStructure(); initial_timestamp = header.stamp; }
    if (result) {
      solver_flag = NON_LINEAR;
      if (TRIANGULATE) triangulate();
      if (last_status == VILS_OK) optimization();             // solveOdometry
      slideWindow();
      removeFailures();
      remember();
    } else slideWindow();
    return;
  }
  if (TRIANGULATE) { triangulate(); if (last_status != VILS_OK) return; }   // solveOdometry (:903-914)
  optimization();
  if (last_status == VILS_OK && failureDetection()) {         // :588-597: reboot
    failure_occur = 1;
    clearState(); setParameter(ric_cfg_, tic_cfg_, td_cfg_);   // back to the configured RIC / TIC / TD (estimator.cpp:21-33)
    return;
  }
  slideWindow();
  if (last_status == VILS_OK) removeFailures();
  remember();
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
  const double* noise = cfg_.imu_noise;                              // ACC_N GYR_N ACC_W GYR_W (parameters.cpp:96-99)
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
  // LPS absolute-rotation constraints (:1283-1326): a new mapper pose arrived -> queue it (moved to the body frame) and add every queued
  // constraint whose bracketing frames are < 0.2 s apart
  std::vector<vils_lps> lps;
  while (LidarLPSConstraints.size() > 7) LidarLPSConstraints.pop_front();
  if (LPS_call_) {
    LidarLPSConstraint cur; double Rq[9], o[3], qrlb[4];
    q2R(LPS_q_, Rq); mat3_vec(Rq, TLB, o);
    for (int i = 0; i < 3; i++) cur.LPSt[i] = o[i] + LPS_t_[i];                       // LPS_t = LPS_q TLB + LPS_t
    R2q(RLB, qrlb); qmul(LPS_q_, qrlb, cur.LPSq);                                       // LPS_q = LPS_q * Quaterniond(RLB)
    cur.lidar_t = LPS_time_;
    LidarLPSConstraints.push_back(cur);
    LPS_call_ = false;
    for (const auto& c : LidarLPSConstraints) {
      int id_l = 0, id_r = 0;
      if (!FindNearest2ID(c.lidar_t, id_l, id_r)) continue;
      if (Headers[id_r].stamp - Headers[id_l].stamp >= 0.2) continue;
      vils_lps q{}; q.tl = Headers[id_l].stamp; q.tr = Headers[id_r].stamp; q.tk = c.lidar_t;
      for (int i = 0; i < 4; i++) q.q[i] = c.LPSq[i];
      q.kf[0] = id_l; q.kf[1] = id_r;
      lps.push_back(q);
    }
  }
  // LiDAR ICP relative constraints (:1345-1398): mode 3 inside the window -> LidarICPConstraint_b; mode 4 (zero velocity) -> frame WINDOW_SIZE-1
  // gets V = 0 and its pose / speed-bias blocks are held constant
  std::vector<vils_icp> icp; std::vector<uint8_t> kffix(F, 0); int nfixed = 0;
  while (LidarICPConstraints.size() > 5) LidarICPConstraints.pop_front();
  for (const auto& c : LidarICPConstraints) {
    int a = 0, b = 0, cc = 0, d = 0;
    if (c.constraint_mode == 4) {
      for (int i = 0; i < 3; i++) sb[9 * (WINDOW_SIZE - 1) + i] = 0.0;
      if (!kffix[WINDOW_SIZE - 1]) nfixed++;
      kffix[WINDOW_SIZE - 1] = 1;
    } else if (c.constraint_mode == 3 && FindWindowsID(c.lidar_ta, c.lidar_tb, c.lidar_tc, c.lidar_td, a, b, cc, d)) {
      if (a == b || a == cc || a == d || b == cc || b == d || cc == d || b > WINDOW_SIZE || d > WINDOW_SIZE) continue;   // ceres rejects duplicate blocks
      vils_icp q{}; q.ta = c.lidar_ta; q.tb = c.lidar_tb; q.tc = c.lidar_tc; q.td = c.lidar_td; q.ti = c.lidar_ti; q.tj = c.lidar_tj;
      q.trans_t[0] = c.lidar_trans[3]; q.trans_t[1] = c.lidar_trans[7]; q.trans_t[2] = c.lidar_trans[11]; q.sqrt_info = c.lidar_sqrt_info00;
      q.kf[0] = a; q.kf[1] = b; q.kf[2] = cc; q.kf[3] = d;
      icp.push_back(q);
    }
  }
  // scan-to-map point factors handed over through setLidarPointFactors (SoA split of the 7 / 9 doubles per row)
  const int npl = (int)lp_plane_kf_.size(), ned = (int)lp_edge_kf_.size();
  std::vector<double> pl_p(3 * (size_t)npl), pl_n(3 * (size_t)npl), pl_d(npl), ed_p(3 * (size_t)ned), ed_a(3 * (size_t)ned), ed_b(3 * (size_t)ned);
  for (int k = 0; k < npl; k++) { for (int a = 0; a < 3; a++) { pl_p[3 * k + a] = lp_plane_[7 * (size_t)k + a]; pl_n[3 * k + a] = lp_plane_[7 * (size_t)k + 3 + a]; } pl_d[k] = lp_plane_[7 * (size_t)k + 6]; }
  for (int k = 0; k < ned; k++) for (int a = 0; a < 3; a++) { ed_p[3 * k + a] = lp_edge_[9 * (size_t)k + a]; ed_a[3 * k + a] = lp_edge_[9 * (size_t)k + 3 + a]; ed_b[3 * k + a] = lp_edge_[9 * (size_t)k + 6 + a]; }
  vils_window w{};
  w.n_kf = F; w.n_feat = (int)lam.size(); w.n_imu = (int)imu_kf.size(); w.n_proj = (int)kf_i.size();
  w.pose = pose.data(); w.speedbias = sb.data(); w.ex_pose = ex.data(); w.inv_depth = lam.data(); w.depth_fixed = dfix.data(); w.kf_fixed = kffix.data(); w.td = td;
  w.imu = pre.data(); w.imu_kf = imu_kf.data();
  w.pts_i = pts_i.data(); w.pts_j = pts_j.data(); w.vel_i = vel_i.data(); w.vel_j = vel_j.data(); w.td_i = td_i.data(); w.td_j = td_j.data();
  w.row_i = row_i.data(); w.row_j = row_j.data(); w.kf_i = kf_i.data(); w.kf_j = kf_j.data(); w.feat = feat.data();
  w.n_plane = npl; w.plane_p = pl_p.data(); w.plane_n = pl_n.data(); w.plane_d = pl_d.data(); w.plane_kf = lp_plane_kf_.data();
  w.n_edge = ned; w.edge_p = ed_p.data(); w.edge_a = ed_a.data(); w.edge_b = ed_b.data(); w.edge_kf = lp_edge_kf_.data();
  w.n_icp = (int)icp.size(); w.icp = icp.data(); w.n_lps = (int)lps.size(); w.lps = lps.data();
  w.prior_n = prior_n_; w.prior_nblk = (int)prior_blk_.size(); w.prior_J = prior_J_.data(); w.prior_r = prior_r_.data(); w.prior_blk = prior_blk_.data(); w.prior_x0 = prior_x0_.data();
  last_n_icp = w.n_icp; last_n_lps = w.n_lps; last_n_fixed = nfixed; last_n_plane = npl; last_n_edge = ned;
  lp_plane_.clear(); lp_plane_kf_.clear(); lp_edge_.clear(); lp_edge_kf_.clear();
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
  // marginalization (:1483-1684): the new prior, block ids already re-addressed to the slid window.  The reference re-packs the RE-ANCHORED
  // state (vector2double, :1487) before it builds MarginalizationInfo, so the prior is linearised — and its x0 snapshots are taken — there.
  last_status = vils_ba_put_state(ba_, 0, pose.data(), sb.data(), ex.data(), lam.data(), tdo);
  if (last_status != VILS_OK) return;
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

// feature_manager.cpp:346-362
void Estimator::removeBack() {
  for (auto it = feature.begin(); it != feature.end();) {
    if (it->start_frame != 0) { it->start_frame--; ++it; continue; }
    it->feature_per_frame.erase(it->feature_per_frame.begin());
    it = it->feature_per_frame.empty() ? feature.erase(it) : std::next(it);
  }
}

#ifndef VILS_HOST_HAS_INITIAL
bool Estimator::initialStructure() { return false; }
#endif

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
    const double t_0 = Headers[0].stamp;
    double back_Q0[4], back_P0[3];
    for (int i = 0; i < 4; i++) back_Q0[i] = Qs[0][i];
    for (int i = 0; i < 3; i++) back_P0[i] = Ps[0][i];
    for (int i = 0; i < Wn; i++) {
      std::swap(Qs[i], Qs[i + 1]); std::swap(imu_[i], imu_[i + 1]); Headers[i] = Headers[i + 1];
      std::swap(Ps[i], Ps[i + 1]); std::swap(Vs[i], Vs[i + 1]); std::swap(Bas[i], Bas[i + 1]); std::swap(Bgs[i], Bgs[i + 1]);
    }
    Headers[Wn] = Headers[Wn - 1]; Ps[Wn] = Ps[Wn - 1]; Vs[Wn] = Vs[Wn - 1]; Qs[Wn] = Qs[Wn - 1]; Bas[Wn] = Bas[Wn - 1]; Bgs[Wn] = Bgs[Wn - 1];
    imu_[Wn] = ImuBuf();
    { auto it0 = all_image_frame.find(t_0); if (it0 != all_image_frame.end()) all_image_frame.erase(all_image_frame.begin(), std::next(it0)); }   // :1731-1747
    if (solver_flag != NON_LINEAR) { removeBack(); return; }   // slideWindowOld without depth shifting while initialising (:1806-1807)
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
    const double t_second_new = Headers[Wn - 1].stamp;
    ImuBuf& dst = imu_[frame_count - 1]; const ImuBuf& src = imu_[frame_count];
    dst.dt.insert(dst.dt.end(), src.dt.begin(), src.dt.end()); dst.acc.insert(dst.acc.end(), src.acc.begin(), src.acc.end()); dst.gyr.insert(dst.gyr.end(), src.gyr.begin(), src.gyr.end());
    Headers[frame_count - 1] = Headers[frame_count]; Ps[frame_count - 1] = Ps[frame_count]; Vs[frame_count - 1] = Vs[frame_count];
    Qs[frame_count - 1] = Qs[frame_count]; Bas[frame_count - 1] = Bas[frame_count]; Bgs[frame_count - 1] = Bgs[frame_count];
    imu_[Wn] = ImuBuf();
    if (solver_flag == INITIAL) all_image_frame.erase(t_second_new);   // :1781-1784
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
  const int B = 1; const int ix = (int)std::lrint(x), iy = (int)std::lrint(y);   // BORDER_SIZE = 1; cvRound = round half to even (lrint in the default rounding mode)
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
  // The frame is uploaded once; CLAHE (EQUALIZE, cv::createCLAHE(3.0, Size(8, 8))->apply, :87-93), the LK pyramids and the corner detector all
  // work on the device-resident image, and its pyramid is kept as the next call's cur_img (:160-164): no image ever comes back.
  last_status = vils_frontend_load(fe_, img, stride, EQUALIZE ? 1 : 0, 3.0, 8, 8);
  if (last_status != VILS_OK) return;
  const uint8_t* forw_dev = nullptr; int32_t pitch = 0; void* ready = nullptr;
  last_status = vils_frontend_current(fe_, &forw_dev, &pitch, &ready);
  if (last_status != VILS_OK) return;
  std::vector<std::array<float, 2>> forw_pts;
  const int n0 = (int)cur_pts.size();
  std::vector<std::array<float, 2>> forw(std::max(n0, 1)); std::vector<uint8_t> status(std::max(n0, 1)); std::vector<float> err(std::max(n0, 1));
  last_status = vils_klt_advance(klt_, forw_dev, pitch, ready, n0 ? &cur_pts[0][0] : nullptr, n0, &forw[0][0], status.data(), err.data());   // :113
  if (last_status != VILS_OK) return;
  if (has_img_ && n0 > 0) {
    const int n = n0;
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
      last_status = vils_good_features_resident(fe_, n_max_cnt, 0.01, (double)MIN_DIST, 1, n_pts.data(), &nn);
      if (last_status != VILS_OK) return;
      for (int k = 0; k < nn; k++) { forw_pts.push_back({n_pts[2 * k], n_pts[2 * k + 1]}); ids.push_back(-1); track_cnt.push_back(1); }   // addPoints (:71-79)
    }
  }
  cur_pts = forw_pts;                                                        // prev = cur, cur = forw (:160-164): the images swapped roles on the device
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

void approximate_voxel_grid_public(const float* xyzi, int n, int stride, float leaf, std::vector<float>& out) { approximate_voxel_grid(xyzi, n, stride, leaf, out); }

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
// ---- LiDAR side of the Estimator
int vh_process_lidar(void* p, float* xyzi, int n, int stride, double cloud_time, double time_offset) {
  auto* e = static_cast<vils::Estimator*>(p); e->processLidar(xyzi, n, stride, cloud_time, time_offset); return e->last_status;
}
void vh_set_lidar_init_flag(void* p, int flag) { static_cast<vils::Estimator*>(p)->lidar_init_flag = flag != 0; }
void vh_set_lps(void* p, const double* q, const double* t, double time) { static_cast<vils::Estimator*>(p)->setLPS(q, t, time); }
void vh_set_lidar_point_factors(void* p, int npl, const double* pl7, const int32_t* pkf, int ned, const double* ed9, const int32_t* ekf) {
  static_cast<vils::Estimator*>(p)->setLidarPointFactors(npl, pl7, pkf, ned, ed9, ekf);
}
// out: keylidar points mode n_icp_queue lidar_count lidar_count_ last_n_icp last_n_lps last_n_fixed last_n_plane last_n_edge failure_occur solver_flag
void vh_get_lidar_info(void* p, double* out) {
  auto* e = static_cast<vils::Estimator*>(p);
  out[0] = e->current_lidar.keylidar; out[1] = e->current_lidar_points; out[2] = e->current_lidar.mode; out[3] = (double)e->LidarICPConstraints.size();
  out[4] = e->lidar_count; out[5] = e->lidar_count_; out[6] = e->last_n_icp; out[7] = e->last_n_lps; out[8] = e->last_n_fixed; out[9] = e->last_n_plane;
  out[10] = e->last_n_edge; out[11] = e->failure_occur; out[12] = e->solver_flag;
}
// constraint k of the ICP queue: mode ta tb tc td ti tj sqrt_info trans(16)
int vh_get_icp(void* p, int k, double* out) {
  auto* e = static_cast<vils::Estimator*>(p);
  if (k < 0 || k >= (int)e->LidarICPConstraints.size()) return 1;
  const auto& c = e->LidarICPConstraints[k];
  out[0] = c.constraint_mode; out[1] = c.lidar_ta; out[2] = c.lidar_tb; out[3] = c.lidar_tc; out[4] = c.lidar_td; out[5] = c.lidar_ti; out[6] = c.lidar_tj; out[7] = c.lidar_sqrt_info00;
  for (int i = 0; i < 16; i++) out[8 + i] = c.lidar_trans[i];
  return 0;
}
int vh_failure_detection(void* p) { return static_cast<vils::Estimator*>(p)->failureDetection() ? 1 : 0; }
void vh_set_solver(void* p, int mode, int iters, double mu, double max_time) {
  auto* e = static_cast<vils::Estimator*>(p); e->solve_opts.mode = mode; e->solve_opts.max_iters = iters; e->solve_opts.mu = mu; e->solve_opts.max_solver_time = max_time;
}
void vh_get_header_stamps(void* p, double* out) { auto* e = static_cast<vils::Estimator*>(p); for (int i = 0; i <= e->WINDOW_SIZE; i++) out[i] = e->Headers[i].stamp; }
void vh_voxel_filter(const float* xyzi, int n, int stride, float leaf, float* out, int* n_out) {
  std::vector<float> o; vils::approximate_voxel_grid_public(xyzi, n, stride, leaf, o); *n_out = (int)(o.size() / 4); std::memcpy(out, o.data(), sizeof(float) * o.size());
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
