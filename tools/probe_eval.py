import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvil_fusion_b200 import cabi, synth, lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
ws = [synth.make_window(2, k) for k in range(8)]
ba = lib.BA(cabi.default_config(), B)
for k in range(B): ba.set_window(k, ws[k % 8])
ba.upload(B)
for it in range(5):
    ba.evaluate_device(B, True); print("evaluate_device ms", ba.last_ms, "GB/s", 769448 * B / ba.last_ms / 1e6)
