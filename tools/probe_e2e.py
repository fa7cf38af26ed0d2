"""Timing probe: end-to-end vils_ba_solve (host buffers -> H2D -> solve -> D2H) for several pipeline chunk sizes (VILS_CHUNK)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvil_fusion_b200 import cabi, synth, lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
ws = [synth.make_window(2, k) for k in range(16)]
ba = lib.BA(cabi.default_config(), B)
for k in range(B):
    ba.set_window(k, ws[k % 16])
opts = cabi.default_solve_opts()
for it in range(6):
    t = time.perf_counter(); ba.solve(B, opts); dt = time.perf_counter() - t
    print(f"VILS_CHUNK={os.environ.get('VILS_CHUNK','default')} e2e ms {dt*1e3:.3f} solves/s {B/dt:.0f}")
s = ba.get_state(B - 1); print("status", s["status"], s["cost_initial"], s["cost_final"])
t = time.perf_counter(); ba.solve(1, opts); dt = time.perf_counter() - t
for it in range(3):
    t = time.perf_counter(); ba.solve(1, opts); dt = time.perf_counter() - t
    print(f"single window e2e ms {dt*1e3:.3f}")
