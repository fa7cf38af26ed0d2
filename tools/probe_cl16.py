import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
from mvil_fusion_b200 import cabi, synth, lib
w = synth.make_window(2, 0)
ba = lib.BA(cabi.default_config(), 1)
opts = cabi.default_solve_opts()
ba.set_window(0, w); ba.upload(1)
ref = None
for g in (1, 8, 16):
    ba.set_cluster(g)
    try:
        ms = []
        for _ in range(8):
            ba.solve_device(1, opts); ms.append(ba.last_ms)
        st = ba.get_state(0)
        if ref is None: ref = st
        d = max(np.abs(st[k] - ref[k]).max() for k in ("pose", "speedbias", "inv_depth"))
        print("cluster", g, "ran as", ba.last_cluster, "ms", np.mean(ms[3:]), "status", st["status"], "max delta vs one CTA", d)
    except Exception as e:
        print("cluster", g, "failed", e)
