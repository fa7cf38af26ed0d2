// ba_layout.h — the "window blob": one sliding window packed by the host into a single contiguous, 16-byte aligned
// byte range (SoA factor arrays, index lists, prior), uploaded with one copy and read by one CTA.
// Shared by the host packer and the kernels of vils_ba.cu.
#pragma once
#include <cstdint>

namespace vb {

enum {  // byte offsets stored in WinHdr::off[]
  OFF_X = 0,        // state: pose 7N | speedbias 9N | ex 7 | td 1 | inv_depth M            (doubles)
  OFF_FIXED,        // uint8: depth_fixed[M] then kf_fixed[N]
  OFF_IMU,          // n_imu x 467 doubles (vils_preint)
  OFF_IMU_KF,       // int32[n_imu]
  OFF_PROJ,         // 14 arrays of n_proj doubles (SoA, landmark-sorted): pts_i xyz, pts_j xyz, vel_i xy, vel_j xy, td_i, td_j, row_i, row_j
  OFF_PROJ_IDX,     // 4 arrays of n_proj int32: kf_i, kf_j, landmark rank, original index
  OFF_LM_START,     // int32[n_lm + 1]  CSR of factors per landmark rank
  OFF_LM_FEAT,      // int32[n_lm]      landmark rank -> feature index
  OFF_PAIR,         // int32[n_pair x 4]: start, count, kf_i, kf_j  (factors grouped by keyframe pair)
  OFF_PAIR_PERM,    // int32[n_proj]    pair-grouped order -> landmark-sorted factor index
  OFF_PAIR_ID,      // int32[N x N]     (i, j) -> pair index or -1
  OFF_PLANE,        // 7 arrays of n_plane doubles (kf-sorted): pb xyz (body frame), n xyz, d
  OFF_PLANE_IDX,    // 2 arrays int32: kf, original index
  OFF_PLANE_START,  // int32[N + 1]
  OFF_EDGE,         // 9 arrays of n_edge doubles (kf-sorted): pb xyz, a xyz, b xyz
  OFF_EDGE_IDX,     // 2 arrays int32: kf, original index
  OFF_EDGE_START,   // int32[N + 1]
  OFF_ICP,          // n_icp x 14 doubles: ta tb tc td ti tj trans_t(3) sqrt_info kf(4, as doubles)
  OFF_LPS,          // n_lps x 9 doubles: tl tr tk q(4 xyzw) kf(2, as doubles)
  OFF_PRIOR_J,      // n x n column-major
  OFF_PRIOR_R,      // n
  OFF_PRIOR_X0,     // concatenated global-size snapshots
  OFF_PRIOR_BLK,    // int32[nblk x 4]: type, index, x0 offset (doubles), first column
  OFF_PRIOR_COL,    // int32[n]: tangent offset of each prior column in the camera system
  OFF_COUNT
};

struct WinHdr {
  int32_t n_kf, n_feat, n_imu, n_proj, n_plane, n_edge, n_icp, n_lps;
  int32_t prior_n, prior_nblk, n_lm, n_pair;
  int32_t bytes;          // used bytes of this blob
  int32_t reserved[3];
  int32_t off[OFF_COUNT + (8 - OFF_COUNT % 8) % 8];
  double td0;             // unused padding to keep 8-byte alignment explicit
};

// state vector offsets (doubles)
__host__ __device__ inline int XP(int k) { return 7 * k; }
__host__ __device__ inline int XS(int N, int k) { return 7 * N + 9 * k; }
__host__ __device__ inline int XE(int N) { return 16 * N; }
__host__ __device__ inline int XT(int N) { return 16 * N + 7; }
__host__ __device__ inline int XL(int N) { return 16 * N + 8; }

// per-slot device scratch (doubles)
struct ScratchLayout {
  int64_t w_imu, E, part, pairpart, pairctx, priorA, priorb0, lmsave, imuprod, Hg, Hvg, linvg, total;
  int32_t Dv_pad;
};

}  // namespace vb
