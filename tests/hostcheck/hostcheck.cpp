// hostcheck.cpp — TEST-ONLY host build of the product's __host__ __device__ factor arithmetic
// (mvil_fusion_b200/csrc/factors.cuh), so that the exact code the kernels run can be compared with the oracle on a
// machine without a GPU.  Never linked into libvils_b200.so; the product has no CPU path.
#include "../../mvil_fusion_b200/csrc/factors.cuh"
using namespace vf;
extern "C" {
void hc_proj_eval(double s_info, double tr_over_row, double half_row, int use_td, const double* c, const double* pi,
                  const double* pj, const double* ex, double lam, double td, double* r, double* J) {
  BaCfg cfg{}; cfg.s_info = s_info; cfg.tr_over_row = tr_over_row; cfg.half_row = half_row; cfg.use_td = use_td;
  proj_eval(cfg, c, pi, pj, ex, lam, td, r, J);
}
void hc_proj_eval_ctx(double s_info, double tr_over_row, double half_row, int use_td, const double* c, const double* pi,
                      const double* pj, const double* ex, double lam, double td, double* r, double* J) {
  BaCfg cfg{}; cfg.s_info = s_info; cfg.tr_over_row = tr_over_row; cfg.half_row = half_row; cfg.use_td = use_td;
  double ctx[PCTX_LD]; proj_pair_ctx(pi, pj, ex, ctx);
  proj_eval_ctx(cfg, ctx, c, lam, td, r, J);
}
void hc_proj_eval_rows(double s_info, double tr_over_row, double half_row, int use_td, const double* c, const double* pi,
                       const double* pj, const double* ex, double lam, double td, double loss_a, double* r, double* J) {
  BaCfg cfg{}; cfg.s_info = s_info; cfg.tr_over_row = tr_over_row; cfg.half_row = half_row; cfg.use_td = use_td;
  proj_eval_rows(cfg, c, q2R(ldq(pi + 3)), q2R(ldq(pj + 3)), q2R(ldq(ex + 3)), ld3(pi), ld3(pj), ld3(ex), lam, td, r, J, loss_a);
}
void hc_proj_eval_loss(double s_info, double tr_over_row, double half_row, int use_td, const double* c, const double* pi,
                       const double* pj, const double* ex, double lam, double td, double loss_a, double* r, double* J) {
  BaCfg cfg{}; cfg.s_info = s_info; cfg.tr_over_row = tr_over_row; cfg.half_row = half_row; cfg.use_td = use_td;
  proj_eval(cfg, c, pi, pj, ex, lam, td, r, J, loss_a);
}
int hc_imu_sqrt_info(const double* cov, double* W) { return imu_sqrt_info(cov, W) ? 0 : 1; }
void hc_imu_eval_raw(const double* pre, const double* G, const double* pi, const double* sbi, const double* pj, const double* sbj, double* r, double* J) {
  for (int i = 0; i < 450; i++) J[i] = 0;
  imu_eval_raw(pre, G, pi, sbi, pj, sbj, r, J);
}
void hc_imu_eval_tbl(const double* pre, const double* G, const double* pi, const double* sbi, const double* pj, const double* sbj, double* r, double* J) {
  double core[IMU_CORE_LD]; uint16_t tbl[450];
  imu_core(pre, G, pi, sbi, pj, sbj, core);
  for (int e = 0; e < 36; e++) core[IC_JAC + e] = imu_core_jac(pre, e);
  imu_build_table(tbl);
  for (int e = 0; e < 450; e++) J[e] = imu_tbl_value(tbl[e], core);
  for (int e = 0; e < 15; e++) r[e] = core[IC_R + e];
}
void hc_imu_eval_parts(const double* pre, const double* G, const double* pi, const double* sbi, const double* pj, const double* sbj, double* r, double* J) {
  for (int i = 0; i < 450; i++) J[i] = 0;
  for (int part = 0; part < IMU_PARTS; part++) imu_eval_part(part, pre, G, pi, sbi, pj, sbj, r, J);
}
double hc_plane_eval(const double* pose, const double* pb, const double* n, double d, double* J) { return plane_eval(pose, ld3(pb), ld3(n), d, J); }
void hc_edge_eval(const double* pose, const double* pb, const double* a, const double* b, double* r, double* J) { edge_eval(pose, ld3(pb), ld3(a), ld3(b), r, J); }
void hc_lps_eval(const double* c, const double* pa, const double* pb, double* r, double* J) { lps_eval(c, pa, pb, r, J); }
void hc_icp_eval(const double* c, const double* pa, const double* pb, const double* pc, const double* pd, double* r, double* J) { icp_eval(c, pa, pb, pc, pd, r, J); }
void hc_prior_dx_pose(const double* x, const double* x0, double* dx) { prior_dx_pose(x, x0, dx); }
void hc_pose_plus(double* x, const double* d) { vm::pose_plus(x, d); }
}
