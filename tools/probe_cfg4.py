"""Timing probe for configs[3] exactly as SURVEY.md 8d states it: 20-KF window, 300 features (3333 projection factors), 5000 LiDAR factors,
3 ICP + 3 LPS constraints and the REAL marginalization prior (n = 136) obtained by solving / marginalising / sliding a 21-frame window
through the product path (bench.build_config4_windows).  D = 307: H and Hv live in the per-window L2 scratch (solve_kernel<false>)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvil_fusion_b200 import cabi, lib
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
cfg = cabi.default_config(max_kf=21, max_feat=320, max_proj=4000, max_lidar=5000)
opts = cabi.default_solve_opts()
ws = bench.build_config4_windows(lib, cfg, 4, opts)
print("prior n", ws[0]["prior_n"], "proj", len(ws[0]["kf_i"]), "plane", len(ws[0]["plane_kf"]), "edge", len(ws[0]["edge_kf"]))
ba = lib.BA(cfg, B)
batch = [ws[k % 4] for k in range(B)]
ba.set_windows(0, batch)
ba.upload(B)
for it in range(3):
    ba.solve_device(B, opts); print("config-4 solve_device ms", ba.last_ms, "solves/s", B / ba.last_ms * 1e3)
ba.solve_device(1, opts); print("config-4 single window ms", ba.last_ms)
arr, keep = lib.BA.window_array(batch)
for it in range(3):
    t = time.perf_counter(); ba.solve_windows(batch, opts, arr); dt = time.perf_counter() - t
print("config-4 e2e (caller arrays) ms", dt * 1e3, "solves/s", B / dt)
s = ba.get_state(0); print("status", s["status"], s["cost_initial"], s["cost_final"])
