// vils_host.h — C++ host side above the C-ABI, mirroring the reference's class API for the hot path so that the ROS nodes
// (estimator_node.cpp / feature_tracker_node.cpp) can keep calling the same methods:
//   vils::Estimator::{processIMU, processImage, optimization, slideWindow}   <- vils_estimator/src/estimator.h:37-50,141
//   vils::FeatureTracker::readImage                                          <- feature_tracker_/src/feature_tracker.h:33
//   vils::TransformToEnd                                                     <- vils_estimator/src/lidar_frontend.h:287
// ROS / Eigen / OpenCV types are replaced by plain structs (Header{stamp}, std::array, raw image pointers).  All heavy
// arithmetic goes through libvils_b200.so (include/vils_cabi.h); there is no CPU implementation of it here.
// Scope: processIMU / processImage (INITIAL bootstrap through initialStructure, NON_LINEAR steady state with failureDetection and
// reboot) / optimization (IMU + projection + LiDAR ICP / LPS / point factors + prior) / slideWindow / processLidar (deskew ->
// voxel filter -> VGICP -> constraint classification).
#pragma once
#include <array>
#include <cstdint>
#include <deque>
#include <map>
#include <utility>
#include <vector>

#include "../../../include/vils_cabi.h"

namespace vils {

struct Header { double stamp = 0.0; };                                        // std_msgs::Header::stamp.toSec()
typedef std::array<double, 8> Feature8;                                       // x y z u v vx vy depth (estimator_node.cpp:485-503)
typedef std::map<int, std::vector<std::pair<int, Feature8>>> ImageFeatures;   // feature id -> [(camera id, xyz_uv_velocity_depth)]

struct FeaturePerFrame { double point[3]; double uv[2]; double velocity[2]; double cur_td; double depth; };   // feature_manager.h:19-45
struct FeaturePerId {                                                         // feature_manager.h:47-72
  int feature_id = 0, start_frame = 0;
  std::vector<FeaturePerFrame> feature_per_frame;
  double estimated_depth = -1.0;
  bool lidar_depth_flag = false;
  int solve_flag = 0;                                                         // 0 not solved, 1 ok, 2 failed
  int endFrame() const { return start_frame + (int)feature_per_frame.size() - 1; }
};

// lidar_backend.h:5-26 / lidar_frontend.h (LidarFrame, VIOData) with Eigen types flattened
struct LidarICPConstraint {
  int constraint_mode = 0;                                   // 1 bad fitness, 2 VIO agrees, 3 VIO drift (constraint used), 4 zero velocity, 5 pure rotation
  double lidar_ta = 0, lidar_tb = 0, lidar_tc = 0, lidar_td = 0, lidar_ti = 0, lidar_tj = 0;
  double lidar_trans[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};   // row-major 4 x 4
  double lidar_sqrt_info00 = 0;                              // lidar_sqrt_info(0, 0): the only entry the functor reads (lidar_backend.h:156-158)
};
struct LidarLPSConstraint { double LPSq[4] = {0, 0, 0, 1}; double LPSt[3] = {0, 0, 0}; double lidar_t = 0; };
struct VIOData { double ti = 0, tj = 0, dt = 0; double Qwbi[4] = {0, 0, 0, 1}, Pwbi[3] = {0, 0, 0}, Vbi[3] = {0, 0, 0}, Qwbj[4] = {0, 0, 0, 1}, Pwbj[3] = {0, 0, 0}, Vbj[3] = {0, 0, 0}; };
struct LidarFrame {
  int frameID = 0; bool keylidar = false; double time = 0;   // lidarData.point_cloud.time
  std::vector<float> cloud;                                  // voxel-filtered, x y z intensity packed
  VIOData vioData; double last_image_t = 0, next_image_t = 0;
  double lidar_R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, lidar_T[3] = {0, 0, 0}; int mode = 0;
};

class Estimator {
 public:
  enum SolverFlag { INITIAL, NON_LINEAR };
  enum MarginalizationFlag { MARGIN_OLD = 0, MARGIN_SECOND_NEW = 1 };

  // window_size = the reference's compile-time WINDOW_SIZE (parameters.h:12), a runtime value here; frames = window_size + 1
  Estimator(const vils_config& cfg, int window_size, int num_iterations = 8);
  ~Estimator();
  Estimator(const Estimator&) = delete;
  Estimator& operator=(const Estimator&) = delete;

  void setParameter(const double ric_rowmajor[9], const double tic[3], double td_);          // estimator.cpp:21-33
  void clearState();                                                                          // estimator.cpp:35-84
  // Bootstrap for the steady state (the reference gets here through initialStructure): sets frame k and marks NON_LINEAR.
  void setFrameState(int k, const double P[3], const double Q_xyzw[4], const double V[3], const double Ba[3], const double Bg[3]);

  void processIMU(double dt, const double linear_acceleration[3], const double angular_velocity[3]);   // estimator.cpp:86-120
  void processImage(const ImageFeatures& image, const Header& header);                                 // estimator.cpp:506-616
  void optimization();                                                                                 // estimator.cpp:1124-1687
  void slideWindow();                                                                                  // estimator.cpp:1689-1814
  void triangulate();                                          // f_manager.triangulate(Ps, tic, ric) in solveOdometry (estimator.cpp:903-914)
  bool TRIANGULATE = true;
  // estimator.cpp:122-504.  cloud: n PCL PointXYZI points (stride floats apart, 8 for PCL), deskewed IN PLACE with the NaN-ed points removed
  // (count in current_lidar_points); cloud_time = CloudData::time; time_ = the camera-LiDAR time offset handed over by the node.
  void processLidar(float* xyzi, int n_points, int stride_floats, double cloud_time, double time_);
  bool failureDetection();                                     // estimator.cpp:1076-1122
  // LPS_call / LPS_q / LPS_t / LPS_time of estimator_node.cpp:552-575: the LiDAR mapper's absolute pose, consumed by the next optimization()
  void setLPS(const double q_xyzw[4], const double t[3], double time);
  // Scan-to-map factors (vils_lidar_associate outputs) attached to window keyframes for the next optimization(): plane rows p(3) n(3) d,
  // edge rows p(3) a(3) b(3), each with its keyframe index.  Cleared after the solve.
  void setLidarPointFactors(int n_plane, const double* plane_pnd7, const int32_t* plane_kf, int n_edge, const double* edge_pab9, const int32_t* edge_kf);
  bool initialStructure();                                     // estimator.cpp:618-871 (host SfM + visual-inertial alignment, csrc/host/vils_initial.cpp)

  std::deque<LidarICPConstraint> LidarICPConstraints;          // estimator.h:150
  std::deque<LidarLPSConstraint> LidarLPSConstraints;
  LidarFrame current_lidar;
  std::map<double, LidarFrame> all_lidar_frame;
  bool lidar_init_flag = true;                                 // lidarCalibration.lidar_init_flag: true until the LiDAR extrinsic has been set
  double RLB[9], TLB[3], TBL[3];                               // p_l = RLB p_b + TLB (estimator.cpp:449-451)
  double LidarTimeStep = 0.1, MinDistance = 0.5, MaxDistance = 70.0, LeafSize = 0.3;   // yaml:125-130
  bool ADD_LIDAR_ICP = true, ADD_LPS = true;                   // yaml add_lidar2lidar / add_lps
  double SOLVER_TIME = 0.05;                                   // yaml max_solver_time
  int lidar_count = 0, lidar_count_ = 0, current_lidar_points = 0, failure_occur = 0;
  int last_n_icp = 0, last_n_lps = 0, last_n_fixed = 0, last_n_plane = 0, last_n_edge = 0;
  double last_P[3] = {0, 0, 0}, last_P0[3] = {0, 0, 0}, last_Q[4] = {0, 0, 0, 1}, last_Q0[4] = {0, 0, 0, 1};
  double initial_timestamp = 0;
  double G_DIRECTION[3] = {0, 0, 0};                           // parameters.cpp:33: start value of the gravity direction in Estimate_vel_g_s_tic
  int device() const { return cfg_.device; }
  const vils_config& config() const { return cfg_; }

  // ---- public state, reference names (estimator.h:67-121) ----
  int WINDOW_SIZE;
  SolverFlag solver_flag = INITIAL;
  MarginalizationFlag marginalization_flag = MARGIN_OLD;
  std::vector<std::array<double, 3>> Ps, Vs, Bas, Bgs;
  std::vector<std::array<double, 4>> Qs;                 // Rs as unit quaternions x y z w
  std::vector<Header> Headers;
  double ric[4] = {0, 0, 0, 1}, tic[3] = {0, 0, 0};      // qic x y z w
  double td = 0.0;
  double g[3] = {0, 0, 9.795};
  int frame_count = 0;
  std::deque<FeaturePerId> feature;                       // f_manager.feature
  double MIN_PARALLAX = 10.0 / 460.0;                     // keyframe_parallax / FOCAL_LENGTH (parameters.cpp:118-119)
  double INIT_DEPTH = 5.0;                                // parameters.cpp:189
  vils_summary last_summary{};
  int last_status = VILS_OK;                              // status of the last optimization() (VILS_OK / VILS_ERR_*)
  vils_solve_opts solve_opts{};

  // what optimization() handed to the library last time (for tests / inspection)
  int last_n_proj = 0, last_n_feat = 0, last_prior_n = 0;

 private:
  bool addFeatureCheckParallax(int frame_count_, const ImageFeatures& image, double td_);   // feature_manager.cpp:45-106
  double compensatedParallax2(const FeaturePerId& it, int frame_count_) const;              // feature_manager.cpp:386-421
  void removeBackShiftDepth();                                                               // feature_manager.cpp:286-344
  void removeFront(int frame_count_);                                                        // feature_manager.cpp:364-384
  void removeFailures();                                                                     // feature_manager.cpp:171-184
  void removeBack();                                                                         // feature_manager.cpp:346-362
  bool FindNearest2ID(double tl, int& id_a, int& id_b) const;                                // lidar_backend.cpp:3-36
  bool FindWindowsID(double ta, double tb, double tc, double td, int& id_a, int& id_b, int& id_c, int& id_d) const;   // :38-97

  vils_config cfg_;
  vils_ba* ba_ = nullptr;
  vils_frontend* init_fe_ = nullptr;                         // RANSAC fundamental matrix of relativePose (created on first use)
  bool first_imu_ = false;
  double acc_0_[3] = {0, 0, 0}, gyr_0_[3] = {0, 0, 0};
  // raw IMU samples per interval (dt_buf / linear_acceleration_buf / angular_velocity_buf, estimator.h:99-101); the
  // pre-integration itself (IntegrationBase) is done on the GPU for all intervals at once in optimization()
  struct ImuBuf { std::vector<double> dt, acc, gyr; double acc0[3], gyr0[3]; double ba[3], bg[3]; bool started = false; };
  std::vector<ImuBuf> imu_;
  // marginalization prior carried across frames (last_marginalization_info, estimator.h:120-121)
  std::vector<double> prior_J_, prior_r_, prior_x0_;
  std::vector<int32_t> prior_blk_;
  int prior_n_ = 0;
  // all_image_frame (estimator.h:118): the stamps of every image still inside / behind the window, with what initialStructure needs
 public:
  struct ImageFrame { double t = 0; ImageFeatures points; double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, T[3] = {0, 0, 0}; bool is_key_frame = false; ImuBuf pre; };
  std::map<double, ImageFrame> all_image_frame;
 private:
  double ric_cfg_[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, tic_cfg_[3] = {0, 0, 0}, td_cfg_ = 0;   // RIC / TIC / TD as configured (reboot)
  ImuBuf tmp_pre_;                                             // tmp_pre_integration (estimator.cpp:104-105,524-526)
  bool LPS_call_ = false; double LPS_q_[4] = {0, 0, 0, 1}, LPS_t_[3] = {0, 0, 0}, LPS_time_ = 0;
  bool first_zv_ = true; double tem_zv_r_[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, tem_zv_t_[3] = {0, 0, 0};
  std::vector<double> lp_plane_, lp_edge_; std::vector<int32_t> lp_plane_kf_, lp_edge_kf_;
};

// feature_tracker_/src/feature_tracker.{h,cpp}.  readImage = CLAHE (EQUALIZE) -> calcOpticalFlowPyrLK -> border check / reduceVector ->
// [PUB_THIS_FRAME: setMask -> goodFeaturesToTrack -> addPoints] -> undistortedPoints (liftProjective + velocities), every image-sized or
// per-point step on the device through the C-ABI (vils_clahe / vils_klt_* / vils_set_mask / vils_good_features / vils_lift_projective).
// rejectWithF (:169-202) runs on the device as well (vils_reject_with_f: deterministic RANSAC, OpenCV's epipolar distance).
class FeatureTracker {
 public:
  FeatureTracker(int rows, int cols, int max_cnt = 150, int device = 0);
  ~FeatureTracker();
  void readImage(const uint8_t* img, int stride, double cur_time);
  // Caller-supplied corners (ids assigned immediately); the detection path of readImage uses ids = -1 + updateID like the reference.
  void addPoints(const float* xy, int n);
  bool updateID(unsigned int i);                            // feature_tracker.cpp:204-214
  void undistortedPoints();                                 // :258-306
  void rejectWithF(std::vector<std::array<float, 2>>& forw_pts);   // :169-202
  double FOCAL_LENGTH = 460.0, F_THRESHOLD = 1.0;            // parameters.h:11, yaml F_threshold
  bool REJECT_WITH_F = true;
  // parameters.h / config yaml of the reference
  bool EQUALIZE = false, PUB_THIS_FRAME = false;
  int MIN_DIST = 30;
  double cam[8] = {356.37000498, 354.92225534, 326.87903275, 250.93806883, -0.29326213, 0.07505211, 0.0002761, -0.00026777};
  std::vector<std::array<float, 2>> cur_pts, prev_pts, cur_un_pts, pts_velocity;
  std::vector<int> ids, track_cnt;
  double cur_time = 0, prev_time = 0;
  int last_status = VILS_OK;
 private:
  bool inBorder(float x, float y) const;                  // feature_tracker.cpp:13-19
  int rows_, cols_, max_cnt_, n_id_ = 0;
  vils_klt* klt_ = nullptr;
  vils_frontend* fe_ = nullptr;
  std::vector<uint8_t> cur_img_;
  std::map<int, std::array<float, 2>> prev_un_pts_map_;
  bool has_img_ = false;
};

// fast_gicp::FastVGICP as estimator.cpp:269-297 drives it: setResolution / setNumThreads / setInputSource / setInputTarget /
// align(guess) / getFitnessScore / getFinalTransformation / hasConverged.  Clouds are PCL PointXYZI buffers (stride floats per
// point, 8 for PCL); the alignment runs on the device through vils_vgicp_align.  Matrices are row-major 4 x 4 floats (Eigen::Matrix4f
// values; Eigen itself is column-major, the ROS-side adapter transposes).
class FastVGICP {
 public:
  FastVGICP(int device = 0);
  void setResolution(double r) { opts_.resolution = r; }
  void setNumThreads(int) {}                                 // OpenMP knob of the CPU implementation; no meaning on the device
  void setNeighborSearchMethod(int n_offsets) { opts_.neighbor_search = n_offsets; }   // 1 | 7 | 27 = DIRECT1 / 7 / 27
  void setMaximumIterations(int n) { opts_.max_iterations = n; }
  void setTransformationEpsilon(double e) { opts_.transformation_epsilon = e; }
  void setRotationEpsilon(double e) { opts_.rotation_epsilon = e; }
  void setInputSource(const float* xyzi, int n, int stride_floats = 8);
  void setInputTarget(const float* xyzi, int n, int stride_floats = 8);
  // aligned (may be null, capacity n_source x 4): the source moved by the float final transformation, like pcl::transformPointCloud
  int align(float* aligned_xyzi = nullptr, const float* guess4x4 = nullptr);
  double getFitnessScore() const { return res_.fitness; }
  const float* getFinalTransformation() const { return final_; }
  const double* getFinalHessian() const { return res_.H; }
  bool hasConverged() const { return res_.converged != 0; }
  vils_vgicp_result last_result() const { return res_; }
  int last_status = VILS_OK;
 private:
  int device_;
  vils_vgicp_opts opts_;
  vils_vgicp_result res_;
  std::vector<float> src_, tgt_;
  float final_[16];
};

// pcl::ApproximateVoxelGrid<PointXYZI>::filter as processLidar uses it (estimator.cpp:241-247): out = x y z intensity packed
void approximate_voxel_grid_public(const float* xyzi, int n, int stride_floats, float leaf, std::vector<float>& out);

// lidar_frontend.h:287 — in place on a PCL PointXYZI buffer (8 floats per point)
int TransformToEnd(float* xyzi, int n_points, const float q_xyzw[4], const float t[3], float time_factor, double min_r, double max_r, int device = 0);

}  // namespace vils
