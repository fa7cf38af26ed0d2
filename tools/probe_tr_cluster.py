"""Trust-region (dogleg / LM) and time-capped solves on the cluster-assisted kernel against the one-CTA kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from mvil_fusion_b200 import cabi, synth, lib
import helpers
cfg = cabi.default_config()
w = synth.make_window(2, 5)
cases = [("dogleg8", cabi.default_solve_opts(cabi.VILS_MODE_DOGLEG, 8, 0.0)), ("dogleg30", cabi.default_solve_opts(cabi.VILS_MODE_DOGLEG, 30, 0.0)),
         ("lm10", cabi.default_solve_opts(cabi.VILS_MODE_LM, 10, 0.0))]
o = cabi.default_solve_opts(cabi.VILS_MODE_GN, 5, 1e-8); o.max_solver_time = 10.0; cases.append(("gn5+cap", o))
for name, opts in cases:
    ref = lib.BA(cfg, 1); ref.set_cluster(1); ref.set_window(0, w); ref.solve(1, opts); r = ref.get_state(0)
    ms = []
    for _ in range(4): ref.solve_device(1, opts); ms.append(ref.last_ms)
    line = f"{name}: one CTA status {r['status']} it {r['iterations']} acc {r['accepted']} cost {r['cost_final']:.9e} ms {np.mean(ms[1:]):.3f}"
    for G in (8, 16):
        h = lib.BA(cfg, 1); h.set_cluster(G); h.set_window(0, w); h.solve(1, opts); s = h.get_state(0)
        ms = []
        for _ in range(4): h.solve_device(1, opts); ms.append(h.last_ms)
        line += f" | G={h.last_cluster} status {s['status']} it {s['iterations']} acc {s['accepted']} cost {s['cost_final']:.9e} delta {helpers.rel_state_delta(s, r):.2e} ms {np.mean(ms[1:]):.3f}"
        h.close()
    print(line, flush=True)
    ref.close()
