import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
from mvil_fusion_b200 import lib
import test_assoc_gpu as ta
rng = np.random.default_rng(3)
edge, surf = ta.room_map(rng, 12000, 60000)
q = np.array([0.01, -0.02, 0.03, 1.0]); q /= np.linalg.norm(q); t = np.array([0.3, -0.2, 0.1])
sc = ta.make_scan(rng, edge, q, t, 1250); ss = ta.make_scan(rng, surf, q, t, 3750)
for name, mp, scan, mode in (("corner 1250 x 12k", edge, sc, 0), ("surf 3750 x 60k", surf, ss, 1)):
    best = 1e9
    for _ in range(6):
        best = min(best, lib.lidar_associate(mp, scan, q, t, mode)[3])
    print(name, "ms", best)
