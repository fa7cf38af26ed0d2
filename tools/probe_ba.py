"""Quick timing probe (not the bench): batched GN-5 solves of config-2 windows + the materialised evaluate kernel."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mvil_fusion_b200 import cabi, synth, lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
uniq = min(B, 16)
ws = [synth.make_window(2, k) for k in range(uniq)]
ba = lib.BA(cabi.default_config(), B)
t = time.time()
for k in range(B):
    ba.set_window(k, ws[k % uniq])
print("pack ms/window", (time.time() - t) * 1e3 / B)
opts = cabi.default_solve_opts()
ba.upload(B)
for it in range(3):
    ba.solve_device(B, opts)
    print("solve_device ms", ba.last_ms, "solves/s", B / ba.last_ms * 1e3)
ba.solve_device(1, opts); print("single window ms", ba.last_ms)
t = time.time(); ba.solve(B, opts); dt = time.time() - t
print("e2e solve (upload+solve+download) ms", dt * 1e3, "solves/s", B / dt)
s = ba.get_state(0); print("status", s["status"], s["cost_initial"], s["cost_final"])
for it in range(2):
    ba.evaluate_device(B, True); print("evaluate_device ms", ba.last_ms)
