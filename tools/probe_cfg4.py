"""Timing probe for configs[3]: 20-KF window, 300 features (3333 projection factors), 5000 LiDAR factors, 3 ICP + 3 LPS constraints.
D = 307: H and Hv live in the per-window L2 scratch (solve_kernel<false>)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvil_fusion_b200 import cabi, synth, lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
cfg = cabi.default_config(max_kf=20, max_feat=300, max_proj=3600, max_lidar=5000)
ws = [synth.make_window(4, k, N=20, M=300, n_lidar=5000, n_icp=3, n_lps=3) for k in range(4)]
ba = lib.BA(cfg, B)
for k in range(B):
    ba.set_window(k, ws[k % 4])
ba.upload(B)
opts = cabi.default_solve_opts()
for it in range(3):
    ba.solve_device(B, opts); print("config-4 solve_device ms", ba.last_ms, "solves/s", B / ba.last_ms * 1e3)
ba.solve_device(1, opts); print("config-4 single window ms", ba.last_ms)
for it in range(3):
    t = time.perf_counter(); ba.solve(B, opts); dt = time.perf_counter() - t
print("config-4 e2e ms", dt * 1e3, "solves/s", B / dt)
s = ba.get_state(0); print("status", s["status"], s["cost_initial"], s["cost_final"])
