"""TransformToEnd on a device-resident PCL cloud: the one-point-per-thread kernel vs the TMA (cp.async.bulk + mbarrier) tile pipeline.
Run once per variant (VILS_DESKEW_TMA=0 / 1): prints device time, GB/s of the 64 B/point algorithmic traffic and a checksum of the result."""
import ctypes as C, hashlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mvil_fusion_b200 import cabi, lib
L = lib.load()
out = {"tma": os.environ.get("VILS_DESKEW_TMA", "default")}
for n in (28800, 1 << 20, 1 << 22):
    rng = np.random.default_rng(n)
    pts8 = np.zeros((n, 8), np.float32); pts8[:, :3] = rng.uniform(-40, 40, (n, 3)); pts8[:, 4] = np.floor(rng.uniform(0, 100, n)) + rng.uniform(0, 0.105, n)
    h = C.c_void_p(); assert L.vils_lidar_dev_alloc(n, 8, 0, C.byref(h)) == 0
    q = np.array([0.01, -0.02, 0.03, 1], np.float32); q /= np.linalg.norm(q); t = np.array([0.1, -0.05, 0.02], np.float32); msv = C.c_float(); tms = []
    for it in range(25):
        assert L.vils_lidar_dev_upload(h, pts8.ctypes.data_as(cabi.c_float_p)) == 0
        assert L.vils_lidar_dev_deskew(h, q.ctypes.data_as(cabi.c_float_p), t.ctypes.data_as(cabi.c_float_p), 10.0, 0.5, 70.0, C.byref(msv)) == 0; tms.append(msv.value)
    res = np.zeros_like(pts8); assert L.vils_lidar_dev_download(h, res.ctypes.data_as(cabi.c_float_p)) == 0
    L.vils_lidar_dev_free(h)
    m = float(np.min(tms[3:]))
    out[str(n)] = {"device_ms_best": m, "device_ms_median": float(np.median(tms[3:])), "gbs": 64 * n / m / 1e6, "sha": hashlib.sha1(res.tobytes()).hexdigest()[:12], "nan_points": int(np.isnan(res[:, 0]).sum())}
print(json.dumps(out))
