"""The C++ host mirror (csrc/host: vils::Estimator / vils::FeatureTracker / vils::TransformToEnd) driven through its C
wrapper: raw IMU samples + per-frame feature observations go through processIMU / processImage exactly as the reference's
estimator_node would feed them; the window it hands to the C-ABI must be the same problem as the one packed directly."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers
from mvil_fusion_b200 import cabi, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host():
    from mvil_fusion_b200 import lib
    lib.load()
    H = C.CDLL(os.path.join(ROOT, "mvil_fusion_b200", "libvils_host.so"))
    H.vh_estimator_create.restype = C.c_void_p
    H.vh_estimator_create.argtypes = [C.POINTER(cabi.VilsConfig), C.c_int, C.c_int, C.c_int, C.c_double]
    for f in ("vh_estimator_destroy", "vh_set_parameter", "vh_set_frame_state", "vh_process_imu", "vh_process_image", "vh_get_frame", "vh_get_info"):
        getattr(H, f).restype = None
    dp = cabi.c_double_p
    H.vh_estimator_destroy.argtypes = [C.c_void_p]
    H.vh_set_parameter.argtypes = [C.c_void_p, dp, dp, C.c_double]
    H.vh_set_frame_state.argtypes = [C.c_void_p, C.c_int, dp, dp, dp, dp, dp]
    H.vh_process_imu.argtypes = [C.c_void_p, C.c_double, dp, dp]
    H.vh_process_image.argtypes = [C.c_void_p, C.c_int, cabi.c_int32_p, dp, C.c_double]
    H.vh_frame_count.argtypes = [C.c_void_p]; H.vh_last_status.argtypes = [C.c_void_p]
    H.vh_get_frame.argtypes = [C.c_void_p, C.c_int, dp, dp, dp, dp, dp]
    H.vh_get_info.argtypes = [C.c_void_p, dp]
    H.vh_tracker_create.restype = C.c_void_p
    H.vh_tracker_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    H.vh_tracker_destroy.argtypes = [C.c_void_p]; H.vh_tracker_destroy.restype = None
    H.vh_tracker_add.argtypes = [C.c_void_p, cabi.c_float_p, C.c_int]; H.vh_tracker_add.restype = None
    H.vh_tracker_read.argtypes = [C.c_void_p, cabi.c_uint8_p, C.c_int, C.c_double]
    H.vh_tracker_get.argtypes = [C.c_void_p, cabi.c_float_p, cabi.c_int32_p, cabi.c_int32_p, C.c_int]
    H.vh_transform_to_end.argtypes = [cabi.c_float_p, C.c_int, cabi.c_float_p, cabi.c_float_p, C.c_float, C.c_double, C.c_double]
    return H


def d(a):
    a = np.ascontiguousarray(a, np.float64)
    return a.ctypes.data_as(cabi.c_double_p)


def frame_features(w, j):
    raw = w["raw"]; M = len(raw["start"])
    ids, feats = [], []
    for f in range(M):
        if raw["start"][f] <= j and (f, j) in raw["obs"]:
            x, y = raw["obs"][(f, j)]; vx, vy = raw["vel"][(f, j)]
            depth = raw["depth"][f] if (w["depth_fixed"][f] and j == raw["start"][f]) else 0.0
            ids.append(f); feats.append([x, y, 1.0, cabi.FX * x + cabi.CX, cabi.FY * y + cabi.CY, vx, vy, depth])
    return np.array(ids, np.int32), np.array(feats, np.float64)


def test_estimator_cpp_builds_the_same_window(host):
    from mvil_fusion_b200 import lib
    N = 10
    w = synth.make_window(1, 3, N=N, M=150, n_lidar=0, ex_prior=False)
    raw = w["raw"]
    cfg = cabi.default_config(max_kf=N, max_feat=400, max_proj=4000, max_lidar=16)
    mu = 1e-3
    est = host.vh_estimator_create(C.byref(cfg), N - 1, 5, cabi.VILS_MODE_GN, mu)
    host.vh_set_parameter(est, d(raw["ric"].reshape(-1)), d(raw["tic"]), w["td"])
    pose, sb = w["pose"], w["speedbias"]

    def set_state(k):
        host.vh_set_frame_state(est, k, d(pose[k, :3]), d(pose[k, 3:]), d(sb[k, 0:3]), d(sb[k, 3:6]), d(sb[k, 6:9]))
    for k in range(N):
        set_state(k)
    kf = raw["kf"]
    host.vh_process_imu(est, synth.IMU_DT, d(raw["acc"][0]), d(raw["gyr"][0]))          # first sample only latches acc_0 / gyr_0
    for j in range(N):
        if j > 0:
            for s in range(kf[j - 1] + 1, kf[j] + 1):
                host.vh_process_imu(est, synth.IMU_DT, d(raw["acc"][s]), d(raw["gyr"][s]))
            set_state(j)                                                                   # undo the IMU prediction: same start as the direct solve
        ids, feats = frame_features(w, j)
        host.vh_process_image(est, len(ids), ids.ctypes.data_as(cabi.c_int32_p), d(feats), float(raw["ts"][kf[j]]))
    assert host.vh_last_status(est) == 0
    info = np.zeros(9); host.vh_get_info(est, d(info))
    assert info[3] == len(w["kf_i"]) and info[4] == 150 and info[2] == 5      # same factor / landmark counts, 5 GN iterations
    assert info[1] < 1e-3 * info[0]
    # the same problem packed directly: biases of frame j linearise interval j-1 -> j, cur_td = td, unknown depths start at INIT_DEPTH
    w2 = dict(w)
    idx = kf[:-1, None] + 1 + np.arange(synth.SAMPLES)[None, :]
    w2["imu"] = lib.preintegrate(np.arange(N) * synth.SAMPLES, np.full((N - 1) * synth.SAMPLES, synth.IMU_DT), raw["acc"][idx].reshape(-1, 3),
                                 raw["gyr"][idx].reshape(-1, 3), raw["acc"][kf[:-1]], raw["gyr"][kf[:-1]], sb[1:, 3:6], sb[1:, 6:9],
                                 np.array([cabi.ACC_N, cabi.GYR_N, cabi.ACC_W, cabi.GYR_W]))
    w2["td_i"] = np.full(len(w["kf_i"]), w["td"]); w2["td_j"] = np.full(len(w["kf_i"]), w["td"])
    w2["inv_depth"] = np.where(w["depth_fixed"] == 1, 1.0 / raw["depth"], 1.0 / 5.0)
    ba = lib.BA(cfg, 1); ba.set_window(0, w2)
    ba.solve(1, cabi.default_solve_opts(cabi.VILS_MODE_GN, 5, mu))
    ref = ba.get_state(0)
    assert ref["status"] == 0
    rp, rs = lib.double2vector(w["pose"][0], ref["pose"], ref["speedbias"])
    # after optimization() the window slid: MARGIN_OLD moves frame k -> k-1, MARGIN_SECOND_NEW keeps 0..N-3 and moves N-1 -> N-2
    P = np.zeros(3); Q = np.zeros(4); V = np.zeros(3); Ba = np.zeros(3); Bg = np.zeros(3)
    old = info[7] == 0
    for k in range(N - 1):
        host.vh_get_frame(est, k, d(P), d(Q), d(V), d(Ba), d(Bg))
        src = k + 1 if old else (k if k < N - 2 else N - 1)
        assert np.abs(P - rp[src, :3]).max() <= 1e-7 and np.abs(V - rs[src, :3]).max() <= 1e-7
        assert min(np.abs(Q - rp[src, 3:]).max(), np.abs(Q + rp[src, 3:]).max()) <= 1e-8
    # one more frame through the steady state: the new solve carries the prior built on the device
    last = raw["ts"][-1]
    for s in range(5):
        host.vh_process_imu(est, synth.IMU_DT, d(raw["acc"][-1]), d(raw["gyr"][-1]))
    ids, feats = frame_features(w, N - 1)
    host.vh_process_image(est, len(ids), ids.ctypes.data_as(cabi.c_int32_p), d(feats), float(last + 0.025))
    info2 = np.zeros(9); host.vh_get_info(est, d(info2))
    assert host.vh_last_status(est) == 0 and info2[5] > 0 and np.isfinite(info2[1])
    host.vh_estimator_destroy(est)


def test_feature_tracker_cpp_and_transform_to_end(host):
    cv2 = pytest.importorskip("cv2")
    from test_klt_gpu import make_pair
    import oracle_lib as ol
    img, nxt, pts = make_pair(7, (5.5, -3.25))
    ft = host.vh_tracker_create(480, 640, 150, 0)
    big = np.zeros((480, 704), np.uint8); big[:, :640] = img          # cv::Mat with step > cols
    assert host.vh_tracker_read(ft, big.ctypes.data_as(cabi.c_uint8_p), 704, 0.0) == 0
    host.vh_tracker_add(ft, pts.ctypes.data_as(cabi.c_float_p), len(pts))
    assert host.vh_tracker_read(ft, np.ascontiguousarray(nxt).ctypes.data_as(cabi.c_uint8_p), 640, 0.033) == 0
    xy = np.zeros((200, 2), np.float32); ids = np.zeros(200, np.int32); cnt = np.zeros(200, np.int32)
    n = host.vh_tracker_get(ft, xy.ctypes.data_as(cabi.c_float_p), ids.ctypes.data_as(cabi.c_int32_p), cnt.ctypes.data_as(cabi.c_int32_p), 200)
    ref, st, _ = cv2.calcOpticalFlowPyrLK(img, nxt, pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3)
    ref = ref.reshape(-1, 2); st = st.reshape(-1).astype(bool)
    inb = (np.rint(ref[:, 0]) >= 1) & (np.rint(ref[:, 0]) < 639) & (np.rint(ref[:, 1]) >= 1) & (np.rint(ref[:, 1]) < 479)
    keep = st & inb
    assert n == keep.sum() and np.array_equal(ids[:n], np.nonzero(keep)[0]) and np.all(cnt[:n] == 2)
    assert np.abs(xy[:n] - ref[keep]).max() <= 0.02
    host.vh_tracker_destroy(ft)
    cloud = np.zeros((500, 8), np.float32); cloud[:, 0] = np.linspace(1, 40, 500); cloud[:, 1] = 2.0; cloud[:, 4] = 7 + np.linspace(0, 0.0999, 500)
    q = synth.small_quat(np.array([0.0, 0.01, 0.03])).astype(np.float32); t = np.array([0.2, 0.0, -0.01], np.float32)
    exp = ol.deskew(cloud, 8, q, t, 10.0, 0.5, 70.0)
    got = cloud.copy()
    assert host.vh_transform_to_end(got.ctypes.data_as(cabi.c_float_p), 500, q.ctypes.data_as(cabi.c_float_p), t.ctypes.data_as(cabi.c_float_p), 10.0, 0.5, 70.0) == 0
    assert np.allclose(got, exp, rtol=2e-6, atol=1e-6, equal_nan=True)
