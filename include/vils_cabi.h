/*
 * vils_cabi.h — C-ABI of libvils_b200.so, the B200-native (sm_100a) hot path of a sliding-window
 * visual-inertial-LiDAR estimator.
 *
 * The reference system (Stan994265/mVIL-Fusion) has no FFI boundary; the seams this library replaces
 * are C++ call sites.  Every entry point cites the reference interface it stands in for
 * (paths relative to the reference tree).  Plain pointers and sizes only; all pointers are HOST
 * pointers unless the name says `_dev`.  The caller owns every host array; the library copies
 * in (pinned staging -> HBM) and never retains a host pointer after the call returns.
 *
 * Error model (reference: Evaluate() always returns true, divergence is caught after the solve by
 * Estimator::failureDetection(), vils_estimator/src/estimator.cpp:1076-1122): every function returns
 * an int status, never aborts, and leaves the stored state unchanged on failure.
 *
 * Threading (reference: optimization() and processLidar() are serialised by m_estimator,
 * vils_estimator/src/estimator_node.cpp:352-367,388-524): calls on one handle must be serialised by
 * the caller; different handles may be used from different host threads.
 */
#ifndef VILS_CABI_H_
#define VILS_CABI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VILS_ABI_VERSION 2

/* ---- status codes ------------------------------------------------------------------------- */
#define VILS_OK 0
#define VILS_ERR_BAD_ARG 1     /* null pointer, negative size, index out of range               */
#define VILS_ERR_NO_DEVICE 2   /* no CUDA device / not sm_100 — there is NO CPU fallback         */
#define VILS_ERR_CUDA 3        /* a CUDA runtime call failed; see vils_last_error()             */
#define VILS_ERR_NOT_FINITE 4  /* a non-finite value appeared in the state or the cost           */
#define VILS_ERR_CHOLESKY 5    /* reduced camera system not positive definite after damping     */
#define VILS_ERR_CAPACITY 6    /* window/batch larger than the handle was created for           */

/* ---- parameter-block ids (replace the pointer identity keys of
 *      vils_estimator/src/factor/marginalization_factor.cpp:100-106) ------------------------ */
#define VILS_BLK_POSE 0        /* para_Pose[k]       global 7, local 6  (estimator.h:110)        */
#define VILS_BLK_SPEEDBIAS 1   /* para_SpeedBias[k]  global 9, local 9  (estimator.h:111)        */
#define VILS_BLK_EXPOSE 2      /* para_Ex_Pose[0]    global 7, local 6  (estimator.h:113)        */
#define VILS_BLK_TD 3          /* para_Td[0]         global 1, local 1  (estimator.h:115)        */
#define VILS_BLK_ID(type, idx) (((int32_t)(type) << 16) | (int32_t)(idx))
#define VILS_BLK_TYPE(id) ((id) >> 16)
#define VILS_BLK_INDEX(id) ((id) & 0xffff)

/* ---- solver modes -------------------------------------------------------------------------- */
#define VILS_MODE_GN 0 /* exactly max_iters Gauss-Newton steps, Jacobi-diagonal regularised by mu   */
#define VILS_MODE_LM 1 /* Levenberg-Marquardt trust region to convergence (<= max_iters)            */
#define VILS_MODE_DOGLEG 2 /* ceres TRADITIONAL_DOGLEG trust region — what the reference configures
                            * (options.trust_region_strategy_type = ceres::DOGLEG, estimator.cpp:1406):
                            * Gauss-Newton step regularised by mu (min_mu 1e-8, x10 on failure), Cauchy point,
                            * dogleg interpolation inside the elliptical region, Jacobi scaling fixed at x0       */

#define VILS_MARGIN_OLD 0        /* estimator.h MarginalizationFlag MARGIN_OLD                      */
#define VILS_MARGIN_SECOND_NEW 1 /* MARGIN_SECOND_NEW                                               */

/* Replaces the mutable globals of vils_estimator/src/parameters.{h,cpp} that the hot path reads. */
typedef struct vils_config {
  double focal_length;       /* FOCAL_LENGTH (parameters.h:11); projection sqrt_info = focal/2 (estimator.cpp:18) */
  double gravity[3];         /* G (parameters.cpp:32,103), subtracted as a positive vector                 */
  double tr;                 /* TR rolling-shutter read-out time (parameters.cpp:203-208)                  */
  double row;                /* ROW image height (parameters.cpp:105)                                      */
  double cauchy_visual;      /* CauchyLoss(1.0) on projection / ICP / LPS (estimator.cpp:1129)             */
  double huber_lidar;        /* HuberLoss(0.1) on LiDAR edge/plane (lidar_mapping/src/localMapping.cpp:597) */
  double rlb[9];             /* RLB row-major: p_l = RLB p_b + TLB (estimator.cpp:449-451)                 */
  double tlb[3];             /* TLB                                                                       */
  int32_t estimate_extrinsic; /* ESTIMATE_EXTRINSIC: 0 => para_Ex_Pose constant (estimator.cpp:1154-1158)  */
  int32_t estimate_td;        /* ESTIMATE_TD: 1 => ProjectionTdFactor + para_Td (estimator.cpp:1162-1232)  */
  int32_t max_kf;            /* capacity: keyframes per window (reference: WINDOW_SIZE+1 = 7)              */
  int32_t max_feat;          /* capacity: landmarks per window (reference: NUM_OF_F = 1000)                */
  int32_t max_proj;          /* capacity: projection factors per window                                   */
  int32_t max_lidar;         /* capacity: plane + edge factors per window                                 */
  int32_t device;            /* CUDA device ordinal                                                       */
  int32_t reserved;
  double imu_noise[4];       /* ACC_N GYR_N ACC_W GYR_W (parameters.cpp:96-99; yaml acc_n gyr_n acc_w gyr_w): the noise of
                              * IntegrationBase (integration_base.h:21-27), used by the host mirror's re-propagation  */
} vils_config;

/* Result of IntegrationBase (vils_estimator/src/factor/integration_base.h:203-220). 467 doubles. */
typedef struct vils_preint {
  double delta_p[3];
  double delta_q[4];        /* x y z w */
  double delta_v[3];
  double lin_ba[3];         /* linearized_ba */
  double lin_bg[3];         /* linearized_bg */
  double sum_dt;
  double jacobian[225];     /* 15x15 column-major (Eigen default), order O_P O_R O_V O_BA O_BG */
  double covariance[225];   /* 15x15 column-major */
} vils_preint;

/* LidarICPConstraint after FindWindowsID (vils_estimator/src/lidar_backend.h:5-17,97-184). */
typedef struct vils_icp {
  double ta, tb, tc, td, ti, tj;
  double trans_t[3];        /* lidar_trans(0:3,3) — the only part the functor reads (:152) */
  double sqrt_info;         /* lidar_sqrt_info(0,0) (:156-158)                             */
  int32_t kf[4];            /* id_a id_b id_c id_d                                         */
} vils_icp;

/* LidarLPSConstraint after FindNearest2ID (lidar_backend.h:19-26,35-95). */
typedef struct vils_lps {
  double tl, tr, tk;
  double q[4];              /* LPSq, x y z w */
  int32_t kf[2];            /* id_l id_r     */
} vils_lps;

/*
 * One sliding window = what Estimator::optimization() hands to ceres::Problem between vector2double()
 * and double2vector() (vils_estimator/src/estimator.cpp:1124-1419).  Factor data is SoA.
 */
typedef struct vils_window {
  int32_t n_kf, n_feat, n_imu, n_proj, n_plane, n_edge, n_icp, n_lps;
  /* state — para_Pose / para_SpeedBias / para_Ex_Pose / para_Feature / para_Td (estimator.h:110-116) */
  const double* pose;          /* n_kf x 7  [px py pz qx qy qz qw]  (estimator.cpp:920-927) */
  const double* speedbias;     /* n_kf x 9  [v ba bg]               (estimator.cpp:929-939) */
  const double* ex_pose;       /* 7         tic, qic                 (estimator.cpp:943-950) */
  const double* inv_depth;     /* n_feat    inverse depth            (feature_manager.cpp:195-212) */
  const uint8_t* depth_fixed;  /* n_feat    lidar_depth_flag => constant block (estimator.cpp:1217-1221); may be NULL */
  const uint8_t* kf_fixed;     /* n_kf      pose+speed-bias constant (zero-velocity mode, estimator.cpp:1368-1370); may be NULL */
  double td;
  /* IMUFactor (factor/imu_factor.h): factor k links keyframes imu_kf[k] and imu_kf[k]+1 */
  const vils_preint* imu;      /* n_imu */
  const int32_t* imu_kf;       /* n_imu */
  /* ProjectionTdFactor / ProjectionFactor (factor/projection_td_factor.cpp:6-19,34-141) */
  const double* pts_i;         /* n_proj x 3 */
  const double* pts_j;         /* n_proj x 3 */
  const double* vel_i;         /* n_proj x 2 */
  const double* vel_j;         /* n_proj x 2 */
  const double* td_i;          /* n_proj */
  const double* td_j;          /* n_proj */
  const double* row_i;         /* n_proj  uv.y of the observation; the library subtracts ROW/2 (:18-19) */
  const double* row_j;         /* n_proj */
  const int32_t* kf_i;         /* n_proj  anchor keyframe (start_frame) */
  const int32_t* kf_j;         /* n_proj  observing keyframe            */
  const int32_t* feat;         /* n_proj  feature_index                 */
  /* LidarPlaneNormFactor / LidarEdgeFactor (lidar_mapping/src/lidarFactor.hpp:12-55,106-138), attached to
   * keyframe kf through the fixed LiDAR<->body extrinsic (vils_config.rlb/tlb) */
  const double* plane_p;       /* n_plane x 3  point in the LiDAR frame of keyframe plane_kf */
  const double* plane_n;       /* n_plane x 3  unit normal (world)                            */
  const double* plane_d;       /* n_plane      negative_OA_dot_norm                           */
  const int32_t* plane_kf;     /* n_plane */
  const double* edge_p;        /* n_edge x 3 */
  const double* edge_a;        /* n_edge x 3   last_point_a (world) */
  const double* edge_b;        /* n_edge x 3   last_point_b (world) */
  const int32_t* edge_kf;      /* n_edge */
  const vils_icp* icp;         /* n_icp (<=5,  estimator.cpp:1345) */
  const vils_lps* lps;         /* n_lps (<=7,  estimator.cpp:1283) */
  /* MarginalizationFactor (factor/marginalization_factor.cpp:340-400): r = r_lin + J_lin * (x [-] x0) */
  int32_t prior_n;             /* residual count n (0 = no prior)   */
  int32_t prior_nblk;          /* number of kept parameter blocks   */
  const double* prior_J;       /* n x n column-major linearized_jacobians */
  const double* prior_r;       /* n linearized_residuals                  */
  const int32_t* prior_blk;    /* prior_nblk block ids (VILS_BLK_ID), in column order of J_lin */
  const double* prior_x0;      /* concatenated global-size x0 snapshots (7/9/7/1 per block)   */
} vils_window;

typedef struct vils_solve_opts {
  int32_t mode;              /* VILS_MODE_GN | VILS_MODE_LM                                      */
  int32_t max_iters;         /* NUM_ITERATIONS (config yaml max_num_iterations, estimator.cpp:1404) */
  double mu;                 /* GN: (H + mu*diag(H)) dx = -g ; Ceres dogleg's min_mu is 1e-8     */
  double lm_initial_radius;  /* LM: Ceres default initial_trust_region_radius = 1e4              */
  double function_tolerance; /* LM: Ceres default 1e-6                                           */
  double parameter_tolerance;/* LM: Ceres default 1e-8                                           */
  double min_relative_decrease; /* LM: Ceres default 1e-3                                        */
  double max_solver_time;    /* seconds, 0 = no cap: options.max_solver_time_in_seconds = SOLVER_TIME (estimator.cpp:1407-1411;
                              * yaml max_solver_time 0.05, x4/5 when MARGIN_OLD).  Checked like ceres does, at the top of every
                              * iteration, against the time THIS window's solve has been running on the device (%globaltimer);
                              * a capped solve keeps its last accepted state and returns VILS_OK with summary.reserved = 1        */
} vils_solve_opts;

typedef struct vils_summary {
  int32_t status;            /* VILS_OK | VILS_ERR_NOT_FINITE | VILS_ERR_CHOLESKY                */
  int32_t iterations;        /* linearisations performed                                         */
  int32_t accepted;          /* steps accepted                                                   */
  int32_t reserved;          /* 1: stopped by max_solver_time                                     */
  double cost_initial;       /* 1/2 sum rho(|r|^2) before                                        */
  double cost_final;         /* ... after                                                        */
} vils_summary;

/* Output of MarginalizationInfo::marginalize + getParameterBlocks (marginalization_factor.cpp:176-338). */
typedef struct vils_prior_out {
  int32_t n;                 /* out: residual count                                              */
  int32_t nblk;              /* out: kept blocks                                                 */
  int32_t m;                 /* out: marginalised dimension                                      */
  int32_t capacity_n;        /* in : J has room for capacity_n^2, r for capacity_n, ...          */
  double* J;                 /* n x n column-major                                               */
  double* r;                 /* n                                                                */
  int32_t* blk;              /* nblk ids ALREADY re-addressed to the slid window (addr_shift, estimator.cpp:1599-1611) */
  double* x0;                /* concatenated global-size snapshots                               */
} vils_prior_out;

typedef struct vils_ba vils_ba;    /* opaque: device buffers, stream, staging */
typedef struct vils_klt vils_klt;  /* opaque: pyramids + point buffers        */

/* ---- library ------------------------------------------------------------------------------- */
int vils_abi_version(void);
const char* vils_last_error(void);
/* Fills a config with the constants of config/mynteye_leishen_indoor.yaml + parameters.h. */
void vils_default_config(vils_config* cfg);
void vils_default_solve_opts(vils_solve_opts* opts);

/* ---- sliding-window BA: Estimator::optimization() (estimator.cpp:1124-1687) ------------------ */
/* max_windows independent windows live on the device at once (batched mode). */
int vils_ba_create(const vils_config* cfg, int32_t max_windows, vils_ba** out);
void vils_ba_destroy(vils_ba* ba);
/* Stage one window (host copy into pinned memory + index lists).  Replaces vector2double() +
 * problem.AddResidualBlock(...) (estimator.cpp:1169-1398). */
int vils_ba_set_window(vils_ba* ba, int32_t slot, const vils_window* w);
/* The same for n windows into slots [slot0, slot0 + n), packed on all host threads (VILS_PACK_THREADS overrides the count). */
int vils_ba_set_windows(vils_ba* ba, int32_t slot0, int32_t n, const vils_window* ws);
/* Host->device copy of every staged window (async on the handle's stream, then synchronised). */
int vils_ba_upload(vils_ba* ba, int32_t n_windows);
/* Device-only solve of slots [0, n_windows): replaces ceres::Solve (estimator.cpp:1400-1414).
 * Reads the uploaded state, writes the solved state to a separate device buffer (idempotent). */
int vils_ba_solve_device(vils_ba* ba, int32_t n_windows, const vils_solve_opts* opts);
/* Device->host copy of solved states + summaries. */
int vils_ba_download(vils_ba* ba, int32_t n_windows);
/* upload + solve_device + download: the call a host makes per optimization(). */
int vils_ba_solve(vils_ba* ba, int32_t n_windows, const vils_solve_opts* opts);
/* vector2double() + AddResidualBlock(...) + ceres::Solve for n independent windows handed over as the caller's own arrays
 * (estimator.cpp:1169-1414): window k is packed into slot k by the host thread pool WHILE earlier chunks are being copied and
 * solved, so the call runs from vils_window arrays to solved states with one synchronisation. */
int vils_ba_solve_windows(vils_ba* ba, int32_t n, const vils_window* ws, const vils_solve_opts* opts);
/* Solved state of one slot (raw solver output, i.e. before double2vector()'s gauge re-anchoring). */
int vils_ba_get_state(vils_ba* ba, int32_t slot, double* pose, double* speedbias, double* ex_pose,
                      double* inv_depth, double* td, vils_summary* summary);
/* Overwrite the solved state of one slot on the device.  The reference runs double2vector() (estimator.cpp:1419) and then
 * vector2double() (:1487) BEFORE it builds MarginalizationInfo, so its prior is linearised — and its x0 snapshots are taken — at
 * the yaw / position re-anchored state: solve -> vils_ba_get_state -> vils_double2vector -> vils_ba_put_state -> vils_ba_marginalize. */
int vils_ba_put_state(vils_ba* ba, int32_t slot, const double* pose, const double* speedbias, const double* ex_pose,
                      const double* inv_depth, double td);
/* Estimator::double2vector() gauge re-anchoring (estimator.cpp:962-1011) applied on the host to a
 * solved state, given the pre-solve pose of frame 0. */
int vils_double2vector(int32_t n_kf, const double* pose0_before, double* pose, double* speedbias);

/* Materialised CostFunction::Evaluate of every factor of one slot at its uploaded state, AFTER the
 * robust corrector of ResidualBlockInfo::Evaluate (marginalization_factor.cpp:37-68):
 *   residuals: concatenated [imu 15 each | proj 2 each | plane 1 | edge 3 | icp 3 | lps 3 | prior n]
 *   jacobians: tangent-space blocks, row-major per factor:
 *     imu 15x30 (pose_i sb_i pose_j sb_j) | proj 2x20 (pose_i pose_j ex lambda td) | plane 1x6 |
 *     edge 3x6 | icp 3x24 | lps 3x12.  Either pointer may be NULL. apply_loss=0 gives raw Evaluate. */
int vils_ba_evaluate(vils_ba* ba, int32_t slot, int32_t apply_loss, double* residuals, double* jacobians);
/* Evaluate-only kernel over all slots, results kept on the device (HBM roofline measurement). */
int vils_ba_evaluate_device(vils_ba* ba, int32_t n_windows, int32_t apply_loss);
/* Reduced camera system of one slot at its uploaded state: S (D x D row-major, D = 15 n_kf + 7),
 * g_r (D) with landmarks eliminated, no damping. Also the cost. */
int vils_ba_linearize(vils_ba* ba, int32_t slot, double* S, double* g, double* cost);
/* Marginalization (estimator.cpp:1483-1684 + marginalization_factor.cpp:110-338) of one slot at its
 * SOLVED state (or the state written by vils_ba_put_state). */
int vils_ba_marginalize(vils_ba* ba, int32_t slot, int32_t flag, vils_prior_out* out);
/* Latency mode.  When few windows are solved at once (the drop-in case: ONE window per optimization()), solves run as one
 * thread-block CLUSTER per window — 16, 8, 4 or 2 SMs split the factor pass, the landmark Schur complement and the assembly, one CTA runs the
 * Cholesky chain — instead of one CTA per window.  cluster_size: 0 = automatic (the largest of 16 / 8 / 4 / 2 with n_windows x size <= SM
 * count; 16 exceeds the portable cluster size and is used only where the device reports that it can co-schedule such a cluster),
 * 1 = always one CTA per window, 2 / 4 / 8 / 16 = forced when it fits.  The environment variable VILS_CLUSTER sets the default.
 * Gauss-Newton solves use a kernel in which the update is distributed too; dogleg, Levenberg-Marquardt and time-capped solves run their
 * trust-region loop on the first CTA of the cluster while all CTAs serve its linearisations. */
int vils_ba_set_cluster(vils_ba* ba, int32_t cluster_size);
int vils_ba_last_cluster(vils_ba* ba, int32_t* cluster_size);
/* Timing of the last vils_ba_solve_device in ms (CUDA events on the handle's stream). */
int vils_ba_last_device_ms(vils_ba* ba, float* ms);
/* Number of kernel launches issued by the last solve/evaluate call. */
int vils_ba_last_launches(vils_ba* ba, int32_t* n);
/* Bytes moved by the last vils_ba_upload (host->device) and vils_ba_download (device->host). */
int vils_ba_last_transfer_bytes(vils_ba* ba, size_t* h2d, size_t* d2h);
/* Factor-sharded mode: ONE window over several GPUs (large windows only; at 10 keyframes the all-reduce latency is of
 * the order of a whole single-GPU iteration).  Every rank stages in slot 0 a window with the FULL state but only ITS share
 * of the factors (landmarks with all their projection factors by feature % G, LiDAR factors round-robin, IMU / prior /
 * ICP / LPS on rank 0) and uploads it.  Per Gauss-Newton iteration:
 *   vils_ba_sharded_linearize : local factors -> partial system [H (D x D row-major, no damping, constant blocks untouched)
 *                               | g (D) | diag for damping (D) | cost] in the device buffer of vils_ba_sharded_buffer;
 *   caller                    : all-reduce (sum) of that buffer over ranks (ncclAllReduce on the device pointer);
 *   vils_ba_sharded_update    : damping + Cholesky + back-substitution (identical on every rank), landmark
 *                               back-substitution and Plus for the local landmarks.
 * Iteration 0 starts from the uploaded state; afterwards vils_ba_download / vils_ba_get_state as usual (a rank owns the
 * inverse depths of its own landmarks; poses, speed-biases, extrinsic and td are identical on all ranks). */
int vils_ba_sharded_buffer(vils_ba* ba, void** dev_ptr, size_t* n_doubles);
int vils_ba_sharded_linearize(vils_ba* ba, int32_t iteration, const vils_solve_opts* opts);
int vils_ba_sharded_update(vils_ba* ba, const vils_solve_opts* opts);
/* Native form ("one ncclAllReduce per GN iteration ... fused directly behind the assembly kernel, same stream, no host
 * sync"): the library owns an NCCL communicator with one rank per GPU.  vils_nccl_unique_id on rank 0, distribute the 128 bytes by any
 * means, vils_ba_sharded_init on every rank; vils_ba_sharded_solve then enqueues, for every Gauss-Newton iteration, linearise ->
 * ncclAllReduce(sum, double, D^2 + 2D + 1) -> update on the handle's stream, exchanges the inverse depths with one last all-reduce (every
 * rank ends with the FULL solved state, readable with vils_ba_get_state) and synchronises once.  libnccl.so.2 is bound at first use. */
int vils_nccl_unique_id(uint8_t id[128]);
int vils_ba_sharded_init(vils_ba* ba, int32_t rank, int32_t nranks, const uint8_t id[128]);
int vils_ba_sharded_solve(vils_ba* ba, const vils_solve_opts* opts, vils_summary* summary);
/* Host-mediated access to the same buffer (tests; transports other than NCCL). */
int vils_ba_sharded_read(vils_ba* ba, double* host);
int vils_ba_sharded_write(vils_ba* ba, const double* host);

/* ---- IMU pre-integration: IntegrationBase::push_back/repropagate (integration_base.h:30-158) - */
/* n_intervals independent intervals; interval k integrates samples [off[k], off[k+1]).  acc/gyr are
 * n_samples x 3; acc0/gyr0/ba/bg are n_intervals x 3; noise = {ACC_N, GYR_N, ACC_W, GYR_W}. */
int vils_preintegrate(int32_t n_intervals, const int32_t* off, const double* dt, const double* acc,
                      const double* gyr, const double* acc0, const double* gyr0, const double* ba,
                      const double* bg, const double noise[4], vils_preint* out, int32_t device);

/* ---- KLT: cv::calcOpticalFlowPyrLK(prev, next, in, out, status, err, Size(21,21), 3)
 *      (feature_tracker_/src/feature_tracker.cpp:113) ----------------------------------------- */
int vils_klt_create(int32_t rows, int32_t cols, int32_t max_pts, int32_t win, int32_t max_level,
                    int32_t device, vils_klt** out);
void vils_klt_destroy(vils_klt* k);
/* prev/next: rows x cols u8 with row stride `stride` bytes. */
int vils_klt_track(vils_klt* k, const uint8_t* prev, const uint8_t* next, int32_t stride,
                   const float* prev_xy, int32_t n, float* next_xy, uint8_t* status, float* err);
/* Device-resident variant for throughput measurement: images already uploaded by vils_klt_upload. */
int vils_klt_upload(vils_klt* k, const uint8_t* prev, const uint8_t* next, int32_t stride,
                    const float* prev_xy, int32_t n);
int vils_klt_track_device(vils_klt* k);
int vils_klt_download(vils_klt* k, float* next_xy, uint8_t* status, float* err);
/* Device-resident chain (FeatureTracker keeps forw_img as the next call's cur_img, feature_tracker.cpp:160-164): next_dev is the new image
 * ALREADY ON THE DEVICE (vils_frontend_current), pitch_bytes apart.  Its pyramid, built by this call, is the next call's `prev`: one pyramid
 * per frame, no image over PCIe.  The first call only loads the image (status is zeroed). */
int vils_klt_advance(vils_klt* k, const uint8_t* next_dev, int32_t pitch_bytes, void* ready_event /* cudaEvent_t of the producer or NULL */,
                     const float* prev_xy, int32_t n, float* next_xy, uint8_t* status, float* err);
int vils_klt_last_device_ms(vils_klt* k, float* ms);

/* ---- the rest of FeatureTracker::readImage around the LK call (feature_tracker_/src/feature_tracker.cpp:81-167) ---------------
 * One handle per camera stream: device images, mask, score map, candidate list. */
typedef struct vils_frontend vils_frontend;
int vils_frontend_create(int32_t rows, int32_t cols, int32_t max_pts, int32_t device, vils_frontend** out);
void vils_frontend_destroy(vils_frontend* f);
/* cv::createCLAHE(clip_limit, Size(tiles_x, tiles_y))->apply(src, dst)  (feature_tracker.cpp:87-93; reference: 3.0, 8 x 8). */
int vils_clahe(vils_frontend* f, const uint8_t* src, int32_t stride, double clip_limit, int32_t tiles_x, int32_t tiles_y,
               uint8_t* dst, int32_t dst_stride);
/* Device-resident first step of readImage (:87-93): ONE upload of the raw frame, CLAHE on the device when equalize != 0; the image the rest of
 * the frame works on stays in HBM (vils_frontend_current -> vils_klt_advance, vils_good_features_resident) and never returns to the host.
 * vils_frontend_load does NOT synchronise: it records an event (returned by vils_frontend_current) that vils_klt_advance waits on in-stream, so
 * one frame costs one host synchronisation. */
int vils_frontend_load(vils_frontend* f, const uint8_t* src, int32_t stride, int32_t equalize, double clip_limit, int32_t tiles_x, int32_t tiles_y);
int vils_frontend_current(vils_frontend* f, const uint8_t** image_dev, int32_t* pitch_bytes, void** ready_event /* may be NULL */);
int vils_good_features_resident(vils_frontend* f, int32_t max_corners, double quality, double min_distance, int32_t use_mask, float* xy_out,
                                int32_t* n_out);
/* FeatureTracker::setMask (:36-69): visit points in descending track_cnt, keep a point if its pixel is still unmasked, then blank a
 * filled disc of `radius` (MIN_DIST) around it exactly as cv::circle(mask, p, radius, 0, -1) rasterises it.  keep_idx[0..*n_keep) are the
 * surviving input indices in pick order.  The mask stays on the device for vils_good_features. */
int vils_set_mask(vils_frontend* f, const float* xy, const int32_t* track_cnt, int32_t n, int32_t radius, int32_t* keep_idx, int32_t* n_keep);
int vils_get_mask(vils_frontend* f, uint8_t* mask, int32_t stride);
/* cv::goodFeaturesToTrack(img, corners, max_corners, quality, min_distance, mask) as called at :149 (min-eigenvalue score, blockSize 3,
 * Sobel 3).  use_mask: 0 none, 1 the mask of the last vils_set_mask.  xy_out needs room for max_corners points (x, y). */
int vils_good_features(vils_frontend* f, const uint8_t* img, int32_t stride, int32_t max_corners, double quality, double min_distance,
                       int32_t use_mask, float* xy_out, int32_t* n_out);
int vils_frontend_get_eig(vils_frontend* f, float* eig /* rows x cols */);
/* PinholeCamera::liftProjective (camera_model/src/camera_models/PinholeCamera.cc:450-510, recursive distortion model, 8 iterations) for
 * n pixel positions; cam = {fx, fy, cx, cy, k1, k2, p1, p2}; rays = n x 3 (mx_u, my_u, 1) — what undistortedPoints() (:258-306) and
 * rejectWithF() (:169-202) consume. */
int vils_lift_projective(vils_frontend* f, const double cam[8], const float* uv, int32_t n, double* rays);
/* cv::findFundamentalMat(pts1, pts2, cv::FM_RANSAC, threshold, 0.99, status) as called by rejectWithF (:169-202) on the virtual-pinhole
 * points (FOCAL_LENGTH * ray + (COL/2, ROW/2)).  Deterministic 1024-hypothesis RANSAC, OpenCV's symmetric epipolar distance.
 * F (may be NULL): the winning fundamental matrix, row-major. */
int vils_reject_with_f(vils_frontend* f, const float* pts1, const float* pts2, int32_t n, double threshold, uint8_t* status, double* F);
int vils_frontend_last_device_ms(vils_frontend* f, float* ms);

/* ---- FeatureManager::triangulate (vils_estimator/src/feature_manager.cpp:214-268): batched DLT + SVD, one feature per thread.
 * Feature f is observed in keyframes start[f], start[f]+1, ... with bearing points pts[off[f]..off[f+1]) (x y z); Ps n_kf x 3, Rs n_kf x 9
 * row-major, tic 3, ric 9 row-major.  depth[f] = svd_V[2]/svd_V[3] (INIT_DEPTH when negative, :262-266). */
int vils_triangulate(int32_t n_feat, const int32_t* start, const int32_t* off, const double* pts, int32_t n_kf, const double* Ps,
                     const double* Rs, const double tic[3], const double ric[9], double init_depth, double* depth, int32_t device);

/* ---- LiDAR: PointProcessor::PointToRing stamp (lidar_compensator/src/PointProcessor.cc:127-341)
 *      + TransformToEnd deskew (vils_estimator/src/lidar_frontend.cpp:1001-1041) --------------- */
/* xyzi: n points, stride_floats floats apart (8 for pcl::PointXYZI: x y z pad intensity pad pad pad).
 * In place, like the reference. */
int vils_deskew(float* xyzi, int32_t n, int32_t stride_floats, const float q[4] /* x y z w */,
                const float t[3], float time_factor, float min_r, float max_r, int32_t device);
/* Ring id + relative time stamp. ring_out[i] = ring id or -1 (dropped); intensity <- int(I) + rel_time.
 * start_ori is taken from the first kept point as in the reference. */
int vils_stamp_rings(float* xyzi, int32_t n, int32_t stride_floats, float lower_deg, float upper_deg,
                     int32_t n_rings, float scan_period, int32_t* ring_out, int32_t device);
/* PointProcessor::PointToRing complete (PointProcessor.cc:106-341): stamp + the ring-major re-ordering of :114-117.  out (capacity n points):
 * the kept points, ring 0 first, scan order inside a ring; ring_start[n_rings + 1].  The input is not modified. */
int vils_point_to_ring(const float* xyzi, int32_t n, int32_t stride_floats, float lower_deg, float upper_deg, int32_t n_rings,
                       float scan_period, float* out, int32_t* ring_start, int32_t device);
/* Fused stamp + deskew on a device-resident cloud (throughput measurement). */
int vils_lidar_dev_alloc(int32_t n, int32_t stride_floats, int32_t device, void** handle);
int vils_lidar_dev_upload(void* handle, const float* xyzi);
int vils_lidar_dev_deskew(void* handle, const float q[4], const float t[3], float time_factor,
                          float min_r, float max_r, float* ms);
int vils_lidar_dev_download(void* handle, float* xyzi);
void vils_lidar_dev_free(void* handle);

/* ---- scan-to-map association: the producer of the LiDAR edge / plane factor arrays (lidar_mapping/src/localMapping.cpp:611-766) ----
 * map / scan: x y z intensity, 4 packed floats per point.  q_w_curr (x y z w), t_w_curr: current scan-to-map pose (pointAssociateToMap,
 * :170-179).  mode 0: corner points, 5-NN, line test l2 > 3 l1 -> out[i] = p(3) a(3) b(3) -  (LidarEdgeFactor::Create inputs, :626-667);
 * mode 1: surface points, 10-NN -> 5 closest intensities, plane fit + 0.2 m check -> out[i] = p(3) n(3) d - - -  (LidarPlaneNormFactor
 * inputs, :675-747).  valid[i] = 1 where the reference adds a residual block; nn_idx (may be NULL): n_scan x 5 map indices used.  The search is
 * exact wherever the reference uses its result (5th neighbour within 1 m); for the other points (valid = 0) nn_idx holds the nearest points found
 * in the cells around the query, 0x7fffffff where there were fewer than five. */
int vils_lidar_associate(const float* map_xyzi, int32_t n_map, const float* scan_xyzi, int32_t n_scan, const double q_w_curr[4],
                         const double t_w_curr[3], int32_t mode, double* out /* n_scan x 10 */, uint8_t* valid, int32_t* nn_idx,
                         float* device_ms, int32_t device);

/* ---- DepthRegister::get_depth (feature_tracker_/src/feature_tracker.h:98-343): LiDAR depth of the tracked features (the 8th feature
 * channel; depth > 0 makes the landmark's inverse depth a constant block, estimator.cpp:1217-1221).  cloud: n x (x y z intensity) in the
 * world frame; T1, T2: the two 3x4 row-major float transforms applied in sequence (:129-139); num_bins = 360; feat: m x (x, y, 1)
 * undistorted features; depth[i] = depth along the camera z axis or -1. */
int vils_depth_register(const float* cloud_xyzi, int32_t n, const float T1[12], const float T2[12], int32_t num_bins,
                        const float* feat_xyz, int32_t m, float* depth, float* device_ms, int32_t device);

/* ---- voxelised GICP scan matching: the producer of the LidarICPConstraint measurement (estimator.cpp:263-303 -> fast_gicp::FastVGICP;
 * algorithm: fast_gicp/gicp/impl/{fast_gicp_impl.hpp:240-298, fast_vgicp_impl.hpp:76-203, lsq_registration_impl.hpp:52-166},
 * fast_gicp/gicp/fast_vgicp_voxel.hpp:109-170, fast_gicp/so3/so3.hpp:56-76).  Clouds: x y z intensity, 4 packed floats per point, HOST
 * pointers.  Transforms: row-major 4 x 4 doubles mapping source coordinates into the target frame.  Tangent order of H, b: rotation(3),
 * translation(3), left perturbation, as in LsqRegistration. */
typedef struct {
  double resolution;              /* voxel edge; FastVGICP default 1.0, the estimator uses 0.5 (estimator.cpp:270) */
  double rotation_epsilon;        /* 2e-3  */
  double transformation_epsilon;  /* 5e-4  */
  double lm_init_lambda_factor;   /* 1e-9  */
  int32_t k_correspondences;      /* 20 (the only supported value: the reference never changes FastGICP's default) */
  int32_t neighbor_search;        /* 1 | 7 | 27 = NeighborSearchMethod::DIRECT1 / 7 / 27 (default DIRECT1) */
  int32_t max_iterations;         /* 64 */
  int32_t lm_max_iterations;      /* 10 */
  int32_t compute_fitness;        /* 1: also pcl::Registration::getFitnessScore() of the result */
  int32_t reserved;
} vils_vgicp_opts;
typedef struct {
  double T[16];                   /* final transformation (double; the reference stores its float cast) */
  double H[36];                   /* getFinalHessian(): H of the last accepted step, identity if none */
  double error;                   /* error sum at the final pose (correspondences of the last linearize) */
  double fitness;                 /* mean squared nearest-neighbour distance after alignment, -1 if not computed */
  int32_t iterations;             /* nr_iterations_ */
  int32_t converged;              /* hasConverged() */
  int32_t n_corr;                 /* correspondences of the last linearize */
  int32_t n_voxels;               /* voxels of the target map */
  int32_t n_linearize;            /* linearize launches */
  float elapsed_ms;               /* after the upload of the clouds to the last kernel, host LM steps included */
} vils_vgicp_result;
void vils_vgicp_default_opts(vils_vgicp_opts* opts);
/* FastGICP::calculate_covariances (PLANE regularisation): cov6 = n x (xx xy xz yy yz zz); nn_idx (may be NULL) = n x 20, nearest first. */
int vils_vgicp_covariances(const float* xyzi, int32_t n, double* cov6, int32_t* nn_idx, int32_t device);
/* One FastVGICP::linearize at T: H (6 x 6 row-major), b (6), *error = sum of w e^T M e.  voxels (may be NULL, capacity n_tgt x 10): the
 * target voxel map, mean(3) cov6(6) num_points at the index of each voxel's first point, zero rows elsewhere. */
int vils_vgicp_linearize(const float* src_xyzi, int32_t n_src, const float* tgt_xyzi, int32_t n_tgt, const double T[16],
                         const vils_vgicp_opts* opts, double* H, double* b, double* error, int32_t* n_corr, int32_t* n_voxels,
                         double* voxels, int32_t device);
/* FastVGICP::align(guess) + getFitnessScore().  guess may be NULL (identity). */
int vils_vgicp_align(const float* src_xyzi, int32_t n_src, const float* tgt_xyzi, int32_t n_tgt, const double guess[16],
                     const vils_vgicp_opts* opts, vils_vgicp_result* result, int32_t device);

#ifdef __cplusplus
}
#endif
#endif /* VILS_CABI_H_ */
