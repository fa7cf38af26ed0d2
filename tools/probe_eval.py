import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvil_fusion_b200 import cabi, synth, lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
ws = [synth.make_window(2, k) for k in range(8)]
ba = lib.BA(cabi.default_config(), B)
for k in range(B): ba.set_window(k, ws[k % 8])
ba.upload(B)
best = 1e9
for it in range(12):
    ba.evaluate_device(B, True); best = min(best, ba.last_ms)
print("VILS_EV_MINB", os.environ.get("VILS_EV_MINB"), "evaluate_device best ms", best, "GB/s", 769448 * B / best / 1e6, "frac", 769448 * B / best / 1e6 / 6544.3)
