"""Reference-like solver schedule: Levenberg-Marquardt to convergence (Ceres defaults, <= 30 iterations) on config-2
windows — GPU batch vs the CPU restatement on all host threads."""
import ctypes, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
from mvil_fusion_b200 import cabi, synth, lib
import oracle_lib as ol
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
ws = [synth.make_window(2, k) for k in range(16)]
opts = cabi.default_solve_opts(cabi.VILS_MODE_LM, 30, 1e-8)
ba = lib.BA(cabi.default_config(), B)
for k in range(B):
    ba.set_window(k, ws[k % 16])
ba.upload(B)
for it in range(3):
    ba.solve_device(B, opts)
print("GPU LM<=30: ms per", B, "windows", ba.last_ms, "solves/s", B / ba.last_ms * 1e3)
ba.download(B)
its = [ba.get_state(k)["iterations"] for k in range(16)]
print("iterations per window (first 16):", its)
cfg = cabi.default_config()
cores = os.cpu_count()
t0 = time.perf_counter()
th = [threading.Thread(target=lambda i=i: ol.solve_window(cfg, ws[i % 16], opts)) for i in range(cores)]
[t.start() for t in th]; [t.join() for t in th]
dt = time.perf_counter() - t0
print("CPU restatement LM<=30:", cores / dt, "solves/s on", cores, "threads")
