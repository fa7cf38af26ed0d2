import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
from mvil_fusion_b200 import cabi, synth, lib
import oracle_lib as ol
w = synth.make_window(config_id=9, window_idx=11, N=6, M=25, n_lidar=120, n_icp=3, n_lps=3)
cfg = cabi.default_config()
ba = lib.BA(cfg, 1); ba.set_window(0, w); ba.upload(1)
r, J = ba.evaluate(0, False)
ro, Jo, _ = ol.evaluate_window(cfg, w, False)
d = np.abs(r - ro); print("bad r idx", np.nonzero(d > 1e-6 * (1 + np.abs(ro)))[0][:20])
dj = np.abs(J - Jo); print("bad J idx", np.nonzero(dj > 1e-6 * (1 + np.abs(Jo)))[0][:20])
S, g, c = ba.linearize(0); So, go, co = ol.linearize_window(cfg, w)
print("cost", c, co, "S err", np.abs(S - So).max() / np.abs(So).max(), "g err", np.abs(g - go).max() / np.abs(go).max())
bad = np.argwhere(np.abs(S - So) > 1e-8 * np.abs(So).max()); print("bad S entries", len(bad), bad[:10].tolist())
ba.solve(1, cabi.default_solve_opts()); print(ba.get_state(0)["status"])
