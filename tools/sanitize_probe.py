"""Small solves for compute-sanitizer (memcheck / racecheck): one config-2 window on the one-CTA kernel, on a cluster, in dogleg mode, and one
20-keyframe window whose reduced system lives in global memory (staged Cholesky)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvil_fusion_b200 import cabi, synth, lib
cfg = cabi.default_config()
w = synth.make_window(2, 3)
gn = cabi.default_solve_opts(cabi.VILS_MODE_GN, 2, 1e-8)
ba = lib.BA(cfg, 2)
for cl in (1, 8):
    ba.set_cluster(cl); ba.solve_windows([w], gn); print("cluster", ba.last_cluster, "status", ba.get_state(0)["status"])
ba.set_cluster(1); ba.solve_windows([w, w], cabi.default_solve_opts(cabi.VILS_MODE_DOGLEG, 3, 0.0)); print("dogleg status", ba.get_state(1)["status"])
ba.set_cluster(8); ba.solve_windows([w], cabi.default_solve_opts(cabi.VILS_MODE_DOGLEG, 3, 0.0)); print("dogleg on a cluster of", ba.last_cluster, "status", ba.get_state(0)["status"])
ba.close()
cfg4 = cabi.default_config(max_kf=21, max_feat=320, max_proj=4000, max_lidar=5000)
w4 = synth.make_window(config_id=4, window_idx=1, N=20, M=120, n_lidar=600, n_icp=2, n_lps=2)
b4 = lib.BA(cfg4, 1)
for cl in (1, 4):
    b4.set_cluster(cl); b4.solve_windows([w4], gn); print("20 KF, cluster", b4.last_cluster, "status", b4.get_state(0)["status"])
b4.set_cluster(4); b4.solve_windows([w4], cabi.default_solve_opts(cabi.VILS_MODE_LM, 2, 0.0)); print("20 KF, LM on a cluster of", b4.last_cluster, "status", b4.get_state(0)["status"])
b4.close()
