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
    H.vh_tracker_config.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, cabi.c_double_p]; H.vh_tracker_config.restype = None
    H.vh_tracker_update_ids.argtypes = [C.c_void_p]; H.vh_tracker_update_ids.restype = None
    H.vh_tracker_get_un.argtypes = [C.c_void_p, cabi.c_float_p, cabi.c_float_p, C.c_int]
    H.vh_transform_to_end.argtypes = [cabi.c_float_p, C.c_int, cabi.c_float_p, cabi.c_float_p, C.c_float, C.c_double, C.c_double]
    H.vh_vgicp_align.argtypes = [cabi.c_float_p, C.c_int, cabi.c_float_p, C.c_int, C.c_int, C.c_double, cabi.c_float_p, cabi.c_float_p, cabi.c_double_p]
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
    # the same problem packed directly: biases of frame j linearise interval j-1 -> j, cur_td = td
    w2 = dict(w)
    idx = kf[:-1, None] + 1 + np.arange(synth.SAMPLES)[None, :]
    w2["imu"] = lib.preintegrate(np.arange(N) * synth.SAMPLES, np.full((N - 1) * synth.SAMPLES, synth.IMU_DT), raw["acc"][idx].reshape(-1, 3),
                                 raw["gyr"][idx].reshape(-1, 3), raw["acc"][kf[:-1]], raw["gyr"][kf[:-1]], sb[1:, 3:6], sb[1:, 6:9],
                                 np.array([cabi.ACC_N, cabi.GYR_N, cabi.ACC_W, cabi.GYR_W]))
    w2["td_i"] = np.full(len(w["kf_i"]), w["td"]); w2["td_j"] = np.full(len(w["kf_i"]), w["td"])
    # unknown depths: FeatureManager::triangulate at the initial state (feature_manager.cpp:214-268), as processImage -> solveOdometry does
    def q2R(q):
        x, y, z, wq = q
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * wq), 2 * (x * z + y * wq)], [2 * (x * y + z * wq), 1 - 2 * (x * x + z * z), 2 * (y * z - x * wq)],
                         [2 * (x * z - y * wq), 2 * (y * z + x * wq), 1 - 2 * (x * x + y * y)]])
    t_start, t_off, t_pts = [], [0], []
    for f in range(150):
        s0 = int(raw["start"][f])
        for j in range(s0, N):
            t_pts.append([raw["obs"][(f, j)][0], raw["obs"][(f, j)][1], 1.0])
        t_start.append(s0); t_off.append(len(t_pts))
    tri = lib.triangulate(t_start, t_off, np.array(t_pts), pose[:, :3], np.stack([q2R(q) for q in pose[:, 3:]]).reshape(N, 9), raw["tic"], raw["ric"].reshape(9), 5.0)
    w2["inv_depth"] = np.where(w["depth_fixed"] == 1, 1.0 / raw["depth"], 1.0 / tri)
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


def test_feature_tracker_full_read_image_sequence(host):
    """vils::FeatureTracker::readImage with EQUALIZE + PUB_THIS_FRAME over a 4-frame sequence against an emulation of
    feature_tracker.cpp:81-167 built from the OpenCV calls the reference makes (CLAHE, calcOpticalFlowPyrLK, circle mask,
    goodFeaturesToTrack) and the liftProjective restatement; rejectWithF is left out on both sides."""
    cv2 = pytest.importorskip("cv2")
    from test_frontend_gpu import texture, ref_set_mask, lift_oracle, CAM
    rows, cols, MAX_CNT, MIN_DIST = 480, 640, 150, 30
    base = texture(21)
    frames = []
    for k in range(4):
        M = cv2.getRotationMatrix2D((cols / 2, rows / 2), 0.4 * k, 1.0); M[:, 2] += (3.7 * k, -2.1 * k)
        frames.append(cv2.warpAffine(base, M, (cols, rows), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101))
    # ---- emulation of the reference
    clahe = cv2.createCLAHE(3.0, (8, 8))
    st = dict(cur_img=None, cur_pts=np.zeros((0, 2), np.float32), ids=[], cnt=[], prev_map={}, n_id=0, t=0.0)
    ref_out = []
    for k, raw in enumerate(frames):
        t = 0.1 * k
        img = clahe.apply(raw)
        forw = np.zeros((0, 2), np.float32)
        if len(st["cur_pts"]):
            f2, status, _ = cv2.calcOpticalFlowPyrLK(st["cur_img"], img, st["cur_pts"].reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3)
            f2 = f2.reshape(-1, 2); status = status.reshape(-1).astype(bool)
            inb = (np.rint(f2[:, 0]) >= 1) & (np.rint(f2[:, 0]) < cols - 1) & (np.rint(f2[:, 1]) >= 1) & (np.rint(f2[:, 1]) < rows - 1)
            keep = status & inb
            forw = f2[keep]; st["ids"] = [i for i, k2 in zip(st["ids"], keep) if k2]; st["cnt"] = [c for c, k2 in zip(st["cnt"], keep) if k2]
        st["cnt"] = [c + 1 for c in st["cnt"]]
        kept, mask = ref_set_mask(forw, st["cnt"], rows, cols, MIN_DIST)
        forw = forw[kept] if len(kept) else np.zeros((0, 2), np.float32)
        st["ids"] = [st["ids"][i] for i in kept]; st["cnt"] = [st["cnt"][i] for i in kept]
        n_max = MAX_CNT - len(forw)
        if n_max > 0:
            npts = cv2.goodFeaturesToTrack(img, n_max, 0.01, MIN_DIST, mask=mask)
            npts = np.zeros((0, 2), np.float32) if npts is None else npts.reshape(-1, 2)
            forw = np.concatenate([forw, npts]).astype(np.float32); st["ids"] += [-1] * len(npts); st["cnt"] += [1] * len(npts)
        un = lift_oracle(CAM, forw)[:, :2].astype(np.float32) if len(forw) else np.zeros((0, 2), np.float32)
        vel = np.zeros_like(un)
        if st["prev_map"]:
            for i, fid in enumerate(st["ids"]):
                if fid != -1 and fid in st["prev_map"]:
                    vel[i] = ((un[i].astype(np.float64) - st["prev_map"][fid]) / (t - st["t"])).astype(np.float32)
        cur_map = {}
        for i, fid in enumerate(st["ids"]):
            cur_map.setdefault(fid, un[i].astype(np.float64))
        st.update(cur_img=img, cur_pts=forw, prev_map=cur_map, t=t)
        for i in range(len(st["ids"])):                      # updateID loop of feature_tracker_node.cpp:120-128
            if st["ids"][i] == -1:
                st["ids"][i] = st["n_id"]; st["n_id"] += 1
        ref_out.append((forw.copy(), list(st["ids"]), list(st["cnt"]), un.copy(), vel.copy()))
    # ---- the C++ host mirror
    ft = host.vh_tracker_create(rows, cols, MAX_CNT, 0)
    host.vh_tracker_config(ft, 1, 1, MIN_DIST, d(CAM))
    for k, raw in enumerate(frames):
        assert host.vh_tracker_read(ft, np.ascontiguousarray(raw).ctypes.data_as(cabi.c_uint8_p), cols, 0.1 * k) == 0
        xy = np.zeros((400, 2), np.float32); ids = np.zeros(400, np.int32); cnt = np.zeros(400, np.int32); un = np.zeros((400, 2), np.float32); vel = np.zeros((400, 2), np.float32)
        host.vh_tracker_update_ids(ft)
        n = host.vh_tracker_get(ft, xy.ctypes.data_as(cabi.c_float_p), ids.ctypes.data_as(cabi.c_int32_p), cnt.ctypes.data_as(cabi.c_int32_p), 400)
        assert host.vh_tracker_get_un(ft, un.ctypes.data_as(cabi.c_float_p), vel.ctypes.data_as(cabi.c_float_p), 400) == n
        rxy, rids, rcnt, run, rvel = ref_out[k]
        assert abs(n - len(rxy)) <= 2, (k, n, len(rxy))
        got = {int(i): j for j, i in enumerate(ids[:n])}
        common = [(j, got[i]) for j, i in enumerate(rids) if i in got]
        assert len(common) >= 0.97 * len(rids), (k, len(common), len(rids))
        jr = np.array([a for a, _ in common]); jg = np.array([b for _, b in common])
        assert np.abs(xy[jg] - rxy[jr]).max() <= 0.05, k
        assert np.array_equal(cnt[jg], np.array(rcnt)[jr]), k
        assert np.abs(un[jg] - run[jr]).max() <= 2e-4, k
        if k > 0:
            assert np.abs(vel[jg] - rvel[jr]).max() <= 5e-3, k       # 0.05 px / 356 px focal / 0.1 s
            assert (cnt[:n] > 1).sum() > 100                        # the sequence really is being tracked
    host.vh_tracker_destroy(ft)
    # with rejectWithF switched on (bit 1) the clean sequence keeps (almost) all of its tracks
    ft = host.vh_tracker_create(rows, cols, MAX_CNT, 0)
    host.vh_tracker_config(ft, 1, 3, MIN_DIST, d(CAM))
    for k, raw in enumerate(frames):
        assert host.vh_tracker_read(ft, np.ascontiguousarray(raw).ctypes.data_as(cabi.c_uint8_p), cols, 0.1 * k) == 0
        host.vh_tracker_update_ids(ft)
    xy = np.zeros((400, 2), np.float32); ids = np.zeros(400, np.int32); cnt = np.zeros(400, np.int32)
    n = host.vh_tracker_get(ft, xy.ctypes.data_as(cabi.c_float_p), ids.ctypes.data_as(cabi.c_int32_p), cnt.ctypes.data_as(cabi.c_int32_p), 400)
    assert (cnt[:n] == 4).sum() >= 0.85 * (np.array(ref_out[-1][2]) == 4).sum()
    host.vh_tracker_destroy(ft)


def test_fast_vgicp_host_class_matches_cabi(host):
    """vils::FastVGICP driven like estimator.cpp:269-297 (PCL PointXYZI stride 8, float guess) gives the C-ABI result cast to float."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vgicp_oracle as vo
    from mvil_fusion_b200 import lib
    rng = np.random.default_rng(5)
    tgt = vo.room_scan(rng, 3000)
    T = np.eye(4); T[:3, :3] = vo.so3_exp(np.array([0.003, 0.002, -0.015])); T[:3, 3] = [-0.1, 0.06, 0.02]
    src = vo.room_scan(rng, 3000, pose=T)
    pcl = lambda c: np.ascontiguousarray(np.c_[c[:, :3], np.ones(len(c)), c[:, 3], np.zeros((len(c), 3))], np.float32)   # x y z pad intensity pad x 3
    ps, pt = pcl(src), pcl(tgt)
    guess = np.eye(4, dtype=np.float32); guess[:3, 3] = [-0.08, 0.05, 0.0]
    Tf = np.zeros(16, np.float32); info = np.zeros(3)
    fp = cabi.c_float_p
    assert host.vh_vgicp_align(ps.ctypes.data_as(fp), len(ps), pt.ctypes.data_as(fp), len(pt), 8, 0.5, guess.ctypes.data_as(fp), Tf.ctypes.data_as(fp), d(info)) == 0
    r = lib.vgicp_align(src, tgt, guess.astype(np.float64), lib.vgicp_opts(0.5))
    assert (Tf.reshape(4, 4) == r["T"].astype(np.float32)).all()
    assert info[0] == r["fitness"] and bool(info[1]) == r["converged"] and int(info[2]) == r["iterations"]
    assert np.abs(Tf.reshape(4, 4)[:3, 3] - T[:3, 3]).max() < 0.03


def _lidar_host(host):
    dp = cabi.c_double_p
    host.vh_process_lidar.argtypes = [C.c_void_p, cabi.c_float_p, C.c_int, C.c_int, C.c_double, C.c_double]
    host.vh_set_lidar_init_flag.argtypes = [C.c_void_p, C.c_int]; host.vh_set_lidar_init_flag.restype = None
    host.vh_set_lps.argtypes = [C.c_void_p, dp, dp, C.c_double]; host.vh_set_lps.restype = None
    host.vh_set_lidar_point_factors.argtypes = [C.c_void_p, C.c_int, dp, cabi.c_int32_p, C.c_int, dp, cabi.c_int32_p]; host.vh_set_lidar_point_factors.restype = None
    host.vh_get_lidar_info.argtypes = [C.c_void_p, dp]; host.vh_get_lidar_info.restype = None
    host.vh_get_icp.argtypes = [C.c_void_p, C.c_int, dp]
    host.vh_failure_detection.argtypes = [C.c_void_p]
    host.vh_set_solver.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]; host.vh_set_solver.restype = None
    host.vh_get_header_stamps.argtypes = [C.c_void_p, dp]; host.vh_get_header_stamps.restype = None
    host.vh_voxel_filter.argtypes = [cabi.c_float_p, C.c_int, C.c_int, C.c_float, cabi.c_float_p, cabi.c_int32_p]; host.vh_voxel_filter.restype = None
    return host


def _sweep(rng, t_mid, n, RLB, TLB, drift=np.zeros(3)):
    """One LiDAR sweep of the synthetic room centred at t_mid (PCL PointXYZI, stride 8): every point is captured from the LiDAR pose at ITS OWN
    time inside [t_mid - 0.05, t_mid + 0.05] (intensity = int + rel_time, PointProcessor.cc:318-331), so the cloud carries real motion distortion.
    drift: a world-frame offset of the sensor (a scan that disagrees with the VIO states)."""
    per = n // 6
    pts = []
    for axis, val in ((0, -8.0), (0, 8.0), (1, -6.0), (1, 6.0), (2, -1.5), (2, 3.0)):
        p = np.stack([rng.uniform(-8, 8, per), rng.uniform(-6, 6, per), rng.uniform(-1.5, 3, per)], 1); p[:, axis] = val; pts.append(p)
    pw = np.concatenate(pts) + rng.normal(0, 0.01, (per * 6, 3))
    rel = rng.uniform(0.0, 0.0999, len(pw))
    P, _, _, R, _ = synth.traj(t_mid - 0.05 + rel)
    pb = np.einsum("nji,nj->ni", R, pw - (P + drift))
    pl = pb @ RLB.T + TLB
    cloud = np.zeros((len(pw), 8), np.float32)
    cloud[:, :3] = pl; cloud[:, 4] = np.floor(rng.uniform(1, 100, len(pw))) + rel
    return cloud


def test_estimator_cpp_process_lidar_and_lidar_factors(host):
    """Estimator::processLidar (estimator.cpp:122-504) through the C++ class only: deskew -> voxel filter -> VGICP against the previous key scan ->
    constraint classification -> ICP queue, then optimization() with the queued LidarICPConstraint / zero-velocity freeze / LPS / point factors."""
    from mvil_fusion_b200 import lib
    H = _lidar_host(host)
    N = 10
    w = synth.make_window(1, 5, N=N, M=150, n_lidar=0, ex_prior=False)
    raw = w["raw"]; kf = raw["kf"]; tk = raw["ts"][kf]
    cfg = cabi.default_config(max_kf=N, max_feat=400, max_proj=4000, max_lidar=64)
    RLB = np.array(list(cfg.rlb)).reshape(3, 3); TLB = np.array(list(cfg.tlb))
    est = H.vh_estimator_create(C.byref(cfg), N - 1, 5, cabi.VILS_MODE_GN, 1e-3)
    H.vh_set_parameter(est, d(raw["ric"].reshape(-1)), d(raw["tic"]), w["td"])
    H.vh_set_lidar_init_flag(est, 0)                      # LiDAR extrinsic known (the reference adopts the configured one after 15 key scans)
    truth = w["truth"]

    def set_true(k, src):
        H.vh_set_frame_state(est, k, d(truth["pose"][src, :3]), d(truth["pose"][src, 3:]), d(truth["speedbias"][src, 0:3]), d(truth["speedbias"][src, 3:6]), d(truth["speedbias"][src, 6:9]))
    for k in range(N):
        set_true(k, k)
    H.vh_process_imu(est, synth.IMU_DT, d(raw["acc"][0]), d(raw["gyr"][0]))
    for j in range(N):                                      # the whole window: the last frame triggers the first optimization() + slideWindow()
        if j > 0:
            for s in range(kf[j - 1] + 1, kf[j] + 1):
                H.vh_process_imu(est, synth.IMU_DT, d(raw["acc"][s]), d(raw["gyr"][s]))
            set_true(j, j)
        ids, feats = frame_features(w, j)
        H.vh_process_image(est, len(ids), ids.ctypes.data_as(cabi.c_int32_p), d(feats), float(tk[j]))
    assert H.vh_last_status(est) == 0
    # processLidar works on a FULL window (solver_flag != INITIAL): look the slid header table up and put the true states back on its frames
    stamps = np.zeros(N); H.vh_get_header_stamps(est, d(stamps))
    src = [int(np.argmin(np.abs(tk - t))) for t in stamps]
    for k in range(N):
        set_true(k, src[k])
    tk = stamps.copy()                                     # from here on frame indices are those of the slid window
    info = np.zeros(13)
    rng = np.random.default_rng(5)
    # scan A between frames 2 and 3, scan B between 4 and 5 (consistent with the VIO states), scan C between 6 and 7 from a sensor that sits 0.3 m off
    def feed(t_mid, drift=np.zeros(3)):
        c = _sweep(rng, t_mid, 12000, RLB, TLB, drift)
        st = H.vh_process_lidar(est, c.ctypes.data_as(cabi.c_float_p), len(c), 8, float(t_mid), 0.0)
        assert st == 0
        H.vh_get_lidar_info(est, d(info))
        return c, info.copy()
    assert tk[2] < tk[3] < tk[4] < tk[5] < tk[6] < tk[7]
    cA, iA = feed(tk[2] + 0.04)
    assert iA[0] == 1 and iA[3] == 0                        # key scan, nothing to match against yet
    kept = int(iA[1]); assert 0.9 * len(cA) <= kept <= len(cA)
    assert np.isfinite(cA[:kept, :3]).all() and np.all(cA[:kept, 4] == np.floor(cA[:kept, 4]))   # deskewed in place, NaN points removed, intensity <- int
    cB, iB = feed(tk[4] + 0.04)
    assert iB[3] == 1 and iB[2] == 2                        # VIO and LiDAR agree: mode 2, queued but not used by the solve
    icp = np.zeros(24); assert H.vh_get_icp(est, 0, d(icp)) == 0
    assert icp[1] == tk[2] and icp[2] == tk[3] and icp[3] == tk[4] and icp[4] == tk[5]
    cC, iC = feed(tk[6] + 0.04, drift=np.array([0.3, 0.0, 0.0]))
    assert iC[3] == 2 and iC[2] == 3                        # 0.3 m disagreement: mode 3 (VIO drift), the constraint carries the LiDAR measurement
    assert H.vh_get_icp(est, 1, d(icp)) == 0
    assert icp[0] == 3 and icp[7] > 100.0                   # sqrt_info = 100 / fitness with fitness < 1
    # measured translation (body frame of scan B's sweep end -> scan C's): truth + the injected 0.3 m world offset seen from body B
    PB, _, _, RB, _ = synth.traj(np.array([tk[4] + 0.09])); PC, _, _, RC, _ = synth.traj(np.array([tk[6] + 0.09]))
    expect = RB[0].T @ (PC[0] + np.array([0.3, 0, 0]) - PB[0])
    assert np.abs(icp[8:24].reshape(4, 4)[:3, 3] - expect).max() < 0.05
    # the last frame arrives: optimization() takes the mode-3 constraint (FindWindowsID), an LPS rotation and a handful of point factors
    Pm, _, _, Rm, _ = synth.traj(np.array([tk[5] + 0.03]))
    q_wl = synth.R_to_quat(Rm[0] @ RLB.T); t_wl = Pm[0] + Rm[0] @ (-RLB.T @ TLB)
    H.vh_set_lps(est, d(q_wl), d(t_wl), float(tk[5] + 0.03))
    Pk, _, _, Rk, _ = synth.traj(tk[:N - 1])
    lf = synth._lidar_factors(np.random.default_rng(9), 40, N - 1, Rk, Pk, RLB, TLB)        # scan-to-map factors on the window frames 0..N-2
    pl7 = np.ascontiguousarray(np.concatenate([lf["plane_p"], lf["plane_n"], lf["plane_d"][:, None]], 1)); ed9 = np.ascontiguousarray(np.concatenate([lf["edge_p"], lf["edge_a"], lf["edge_b"]], 1))
    H.vh_set_lidar_point_factors(est, len(pl7), d(pl7), lf["plane_kf"].ctypes.data_as(cabi.c_int32_p), len(ed9), d(ed9), lf["edge_kf"].ctypes.data_as(cabi.c_int32_p))
    for s in range(5):
        H.vh_process_imu(est, synth.IMU_DT, d(raw["acc"][-1]), d(raw["gyr"][-1]))
    ids, feats = frame_features(w, N - 1)
    H.vh_process_image(est, len(ids), ids.ctypes.data_as(cabi.c_int32_p), d(feats), float(raw["ts"][-1] + 0.025))
    assert H.vh_last_status(est) == 0
    H.vh_get_lidar_info(est, d(info))
    assert info[6] == 1 and info[7] == 1 and info[8] == 0   # one ICP + one LPS constraint entered the solve, nothing frozen
    assert info[9] == 30 and info[10] == 10                 # and the 30 plane + 10 edge factors
    inf = np.zeros(9); H.vh_get_info(est, d(inf))
    assert np.isfinite(inf[1]) and inf[1] < inf[0]
    assert H.vh_failure_detection(est) == 0
    H.vh_estimator_destroy(est)


def test_estimator_cpp_zero_velocity_and_reboot(host):
    """Two identical scans -> constraint_mode 4 (zero velocity): the next optimization() freezes frame WINDOW_SIZE-1 (estimator.cpp:1368-1370).
    failureDetection() (:1076-1122) trips on an absurd bias and processImage reboots the estimator (:588-597)."""
    H = _lidar_host(host)
    N = 8
    w = synth.make_window(1, 8, N=N, M=60, n_lidar=0, ex_prior=False)
    raw = w["raw"]; kf = raw["kf"]; tk = raw["ts"][kf]
    cfg = cabi.default_config(max_kf=N, max_feat=200, max_proj=2000, max_lidar=16)
    RLB = np.array(list(cfg.rlb)).reshape(3, 3); TLB = np.array(list(cfg.tlb))
    est = H.vh_estimator_create(C.byref(cfg), N - 1, 4, cabi.VILS_MODE_GN, 1e-3)
    H.vh_set_parameter(est, d(raw["ric"].reshape(-1)), d(raw["tic"]), w["td"])
    H.vh_set_lidar_init_flag(est, 0)
    pose, sb = w["pose"], w["speedbias"]
    for k in range(N):
        H.vh_set_frame_state(est, k, d(pose[k, :3]), d(pose[k, 3:]), d(sb[k, 0:3]), d(sb[k, 3:6]), d(sb[k, 6:9]))
    H.vh_process_imu(est, synth.IMU_DT, d(raw["acc"][0]), d(raw["gyr"][0]))
    for j in range(N):
        if j > 0:
            for s in range(kf[j - 1] + 1, kf[j] + 1):
                H.vh_process_imu(est, synth.IMU_DT, d(raw["acc"][s]), d(raw["gyr"][s]))
            H.vh_set_frame_state(est, j, d(pose[j, :3]), d(pose[j, 3:]), d(sb[j, 0:3]), d(sb[j, 3:6]), d(sb[j, 6:9]))
        ids, feats = frame_features(w, j)
        H.vh_process_image(est, len(ids), ids.ctypes.data_as(cabi.c_int32_p), d(feats), float(tk[j]))
    assert H.vh_last_status(est) == 0
    stamps = np.zeros(N); H.vh_get_header_stamps(est, d(stamps))
    # a stationary platform as far as the states go: identical poses and zero velocity on frames 2..6 of the (full, slid) window, identical scans
    for k in (2, 3, 4, 5, 6):
        H.vh_set_frame_state(est, k, d(pose[2, :3]), d(pose[2, 3:]), d(np.zeros(3)), d(sb[k, 3:6]), d(sb[k, 6:9]))
    rng = np.random.default_rng(3)
    scan = synth.room_scan(rng, 9000)
    c8 = np.zeros((len(scan), 8), np.float32); c8[:, :3] = scan[:, :3]; c8[:, 4] = np.floor(scan[:, 3]) + 0.05
    info = np.zeros(13)
    for t_mid in (stamps[2] + 0.04, stamps[4] + 0.04):
        c = c8.copy()
        assert H.vh_process_lidar(est, c.ctypes.data_as(cabi.c_float_p), len(c), 8, float(t_mid), 0.0) == 0
    H.vh_get_lidar_info(est, d(info))
    assert info[2] == 4 and info[3] == 1                    # zero velocity
    for s in range(5):
        H.vh_process_imu(est, synth.IMU_DT, d(raw["acc"][-1]), d(raw["gyr"][-1]))
    ids, feats = frame_features(w, N - 1)
    H.vh_process_image(est, len(ids), ids.ctypes.data_as(cabi.c_int32_p), d(feats), float(raw["ts"][-1] + 0.025))
    assert H.vh_last_status(est) == 0
    H.vh_get_lidar_info(est, d(info))
    assert info[8] == 1                                     # frame WINDOW_SIZE-1 held constant in that solve
    # reboot: an absurd accelerometer bias on the newest frame
    big = np.array([3.0, 0.0, 0.0])
    H.vh_set_frame_state(est, N - 1, d(pose[N - 1, :3]), d(pose[N - 1, 3:]), d(sb[N - 1, 0:3]), d(big), d(sb[N - 1, 6:9]))
    assert H.vh_failure_detection(est) == 1
    H.vh_estimator_destroy(est)


def test_voxel_filter_matches_its_definition(host):
    """ApproximateVoxelGrid restatement: every output point is the float centroid of consecutive input points falling in one 0.3 m voxel; the
    output never has more points than the input, each input point contributes to exactly one centroid (mass conservation)."""
    H = _lidar_host(host)
    rng = np.random.default_rng(0)
    pts = synth.room_scan(rng, 6000)
    out = np.zeros_like(pts); n_out = C.c_int32()
    H.vh_voxel_filter(pts.ctypes.data_as(cabi.c_float_p), len(pts), 4, 0.3, out.ctypes.data_as(cabi.c_float_p), C.cast(C.byref(n_out), cabi.c_int32_p))
    m = n_out.value
    assert 0 < m < len(pts)
    # replay in numpy
    inv = np.float32(1.0) / np.float32(0.3)
    hist = {}
    ref = []
    for p in pts:
        ix, iy, iz = (int(np.floor(p[0] * inv)), int(np.floor(p[1] * inv)), int(np.floor(p[2] * inv)))
        h = (ix * 7171 + iy * 3079 + iz * 4231) & 511
        e = hist.get(h)
        if e is not None and e[0] != (ix, iy, iz):
            ref.append(e[1] / np.float32(e[2])); e = None
        if e is None:
            e = [(ix, iy, iz), np.zeros(4, np.float32), 0]
        e[1] = e[1] + p; e[2] += 1; hist[h] = e
    for h in sorted(hist):
        ref.append(hist[h][1] / np.float32(hist[h][2]))
    ref = np.array(ref, np.float32)
    assert len(ref) == m
    np.testing.assert_allclose(out[:m], ref, rtol=1e-6, atol=1e-6)


def test_estimator_cpp_initial_structure_from_scratch(host):
    """processImage in the INITIAL state (estimator.cpp:553-580): fill the window, then initialStructure = relativePose (RANSAC F on the device +
    recoverPose) -> global SfM (PnP, triangulation, bundle adjustment) -> visual-inertial alignment (gyro bias / extrinsic rotation / time offsets,
    then velocities / gravity / per-frame scale / accelerometer bias) -> gravity-aligned, metric window -> first optimization().
    Checked against the synthetic truth: metric scale, gravity, relative rotation and velocity."""
    H = _lidar_host(host)
    N = 11
    # keyframes 0.3 s apart (the reference's window only keeps frames with enough parallax): 3 s of motion gives the accelerometer something to
    # measure — with 0.1 s spacing the scale is barely observable and any regularised solver, ceres included, drifts to a near-static solution
    old = (synth.KF_DT, synth.SAMPLES)
    synth.KF_DT, synth.SAMPLES = 0.3, 60
    try:
        w = synth.make_window(1, 21, N=N, M=150, n_lidar=0, ex_prior=False, start=np.zeros(150, np.int32))   # every landmark seen from the first frame on
    finally:
        synth.KF_DT, synth.SAMPLES = old
    raw = w["raw"]; kf = raw["kf"]; tk = raw["ts"][kf]; truth = w["truth"]
    w["depth_fixed"] = np.zeros(150, np.uint8)               # a camera-only bootstrap: no LiDAR depths
    cfg = cabi.default_config(max_kf=N, max_feat=400, max_proj=4000, max_lidar=16)
    est = H.vh_estimator_create(C.byref(cfg), N - 1, 8, cabi.VILS_MODE_DOGLEG, 0.0)
    H.vh_set_parameter(est, d(raw["ric"].reshape(-1)), d(raw["tic"]), 0.0)
    H.vh_process_imu(est, synth.IMU_DT, d(raw["acc"][0]), d(raw["gyr"][0]))
    info = np.zeros(13)
    for j in range(N):
        if j > 0:
            for s in range(kf[j - 1] + 1, kf[j] + 1):
                H.vh_process_imu(est, synth.IMU_DT, d(raw["acc"][s]), d(raw["gyr"][s]))
        ids, feats = frame_features(w, j)
        H.vh_process_image(est, len(ids), ids.ctypes.data_as(cabi.c_int32_p), d(feats), float(tk[j]))
    H.vh_get_lidar_info(est, d(info))
    assert H.vh_last_status(est) == 0
    assert info[12] == 1                                      # solver_flag == NON_LINEAR: the bootstrap succeeded
    # the window slid once after the first optimization: frame k of the estimator is synthetic frame k + 1 (MARGIN_OLD) — identify by stamps
    stamps = np.zeros(N); H.vh_get_header_stamps(est, d(stamps))
    src = [int(np.argmin(np.abs(tk - t))) for t in stamps]
    P = np.zeros((N, 3)); Q = np.zeros((N, 4)); V = np.zeros((N, 3)); Ba = np.zeros(3); Bg = np.zeros(3)
    for k in range(N):
        H.vh_get_frame(est, k, d(P[k]), d(Q[k]), d(V[k]), d(Ba), d(Bg))
        P[k], Q[k], V[k] = P[k].copy(), Q[k].copy(), V[k].copy()
    P = np.array([np.frombuffer(p.tobytes(), np.float64) for p in P])
    a, b = 0, N - 3
    # metric scale: distance travelled between two frames (gauge free)
    d_est = np.linalg.norm(P[b] - P[a]); d_true = np.linalg.norm(truth["pose"][src[b], :3] - truth["pose"][src[a], :3])
    assert abs(d_est / d_true - 1.0) < 0.1, (d_est, d_true)
    # relative rotation between the two frames
    def rel(qa, qb):
        return synth.quat_mul(synth.quat_conj(qa), qb)
    qe = rel(Q[a], Q[b]); qt = rel(truth["pose"][src[a], 3:], truth["pose"][src[b], 3:])
    ang = 2 * np.arccos(min(1.0, abs(float(np.dot(qe / np.linalg.norm(qe), qt / np.linalg.norm(qt))))))
    assert ang < 0.02, ang
    # gravity alignment: the world z axis of the estimate is the true vertical (roll / pitch of a frame agree with the truth up to yaw)
    ze = synth.quat_rot(synth.quat_conj(Q[a]), np.array([0, 0, 1.0])); zt = synth.quat_rot(synth.quat_conj(truth["pose"][src[a], 3:]), np.array([0, 0, 1.0]))
    assert np.arccos(min(1.0, float(np.dot(ze, zt)))) < 0.03
    # speed
    assert abs(np.linalg.norm(V[a]) - np.linalg.norm(truth["speedbias"][src[a], :3])) < 0.15
    H.vh_estimator_destroy(est)
