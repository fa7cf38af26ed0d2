"""Timing probe: end to end from caller arrays (vils_ba_solve_windows: host pack | H2D | solve | D2H) and from pre-packed staging
(vils_ba_solve), plus the pack alone.  Environment knobs: VILS_CHUNK, VILS_PACK_THREADS, VILS_PACK_NT."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvil_fusion_b200 import cabi, synth, lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
U = int(sys.argv[2]) if len(sys.argv) > 2 else 74
ws = [synth.make_window(2, k) for k in range(U)]
batch = [ws[k % U] for k in range(B)]
ba = lib.BA(cabi.default_config(), B)
arr, keep = lib.BA.window_array(batch)
opts = cabi.default_solve_opts()
tag = f"chunk={os.environ.get('VILS_CHUNK','d')} threads={os.environ.get('VILS_PACK_THREADS','d')} nt={os.environ.get('VILS_PACK_NT','1')}"
best = 1e9
for it in range(5):
    t = time.perf_counter(); ba.set_windows(0, batch, arr); best = min(best, time.perf_counter() - t)
print(f"{tag} pack-only ms {best*1e3:.3f} ({best/B*1e6:.2f} us/window)")
best = 1e9
for it in range(8):
    t = time.perf_counter(); ba.solve_windows(batch, opts, arr); best = min(best, time.perf_counter() - t)
print(f"{tag} e2e caller-arrays ms {best*1e3:.3f} solves/s {B/best:.0f}")
best = 1e9
for it in range(6):
    t = time.perf_counter(); ba.solve(B, opts); best = min(best, time.perf_counter() - t)
print(f"{tag} e2e pre-packed ms {best*1e3:.3f} solves/s {B/best:.0f}")
s = ba.get_state(B - 1); print("status", s["status"], s["cost_final"])
for it in range(3):
    t = time.perf_counter(); ba.solve_windows(batch[:1], opts, arr); dt = time.perf_counter() - t
print(f"single window e2e from caller arrays ms {dt*1e3:.3f}")
