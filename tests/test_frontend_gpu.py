"""The rest of FeatureTracker::readImage around the LK call (feature_tracker_/src/feature_tracker.cpp:81-167) on the GPU, against the very
functions the reference calls: cv2.createCLAHE(3.0, (8, 8)), cv2.circle-based setMask, cv2.goodFeaturesToTrack(img, N, 0.01, 30, mask)
(OpenCV 4.13.0), and a numpy restatement of PinholeCamera::liftProjective (camera_model/src/camera_models/PinholeCamera.cc:450-510,646-662).
Bars: CLAHE and the mask bit-exact; corners identical to OpenCV's list up to score ties (>= 97 % identical positions, same order on the
common prefix); min-eigenvalue map within 1e-6 of its maximum; lifted rays within 1e-14."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
pytestmark = pytest.mark.gpu

CAM = np.array([356.37000498, 354.92225534, 326.87903275, 250.93806883, -0.29326213, 0.07505211, 0.0002761, -0.00026777])   # config/mynteye_leishen_indoor.yaml:13-22


def texture(seed, rows=480, cols=640):
    rng = np.random.default_rng(seed)
    img = cv2.GaussianBlur(rng.uniform(0, 255, (rows, cols)).astype(np.float32), (0, 0), 2.0)
    return cv2.normalize(img, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)


def lift_oracle(cam, uv):
    """PinholeCamera::liftProjective, recursive distortion model (n = 8), FP64."""
    fx, fy, cx, cy, k1, k2, p1, p2 = cam
    out = np.zeros((len(uv), 3))
    for i, (u, v) in enumerate(np.asarray(uv, np.float64)):
        mx_d = (1.0 / fx) * u + (-cx / fx); my_d = (1.0 / fy) * v + (-cy / fy)
        ux, uy = mx_d, my_d
        for _ in range(8):
            mx2, my2, mxy = ux * ux, uy * uy, ux * uy
            rho2 = mx2 + my2; rad = k1 * rho2 + k2 * rho2 * rho2
            dux = ux * rad + 2.0 * p1 * mxy + p2 * (rho2 + 2.0 * mx2); duy = uy * rad + 2.0 * p2 * mxy + p1 * (rho2 + 2.0 * my2)
            ux, uy = mx_d - dux, my_d - duy
        out[i] = (ux, uy, 1.0)
    return out


@pytest.mark.parametrize("seed,shape", [(1, (480, 640)), (2, (480, 640)), (3, (240, 320)), (4, (100, 150))])
def test_clahe_bit_exact(seed, shape):
    from mvil_fusion_b200 import lib
    img = texture(seed, *shape)
    if seed == 2:
        img = (img // 3 + 40).astype(np.uint8)     # low-contrast input: heavy clipping / redistribution
    ref = cv2.createCLAHE(3.0, (8, 8)).apply(img)
    f = lib.Frontend(*shape)
    out = f.clahe(img, 3.0, (8, 8))
    assert np.array_equal(out, ref), np.abs(out.astype(int) - ref.astype(int)).max()
    f.close()


def test_lift_projective_matches_reference_formula():
    from mvil_fusion_b200 import lib
    rng = np.random.default_rng(7)
    uv = np.stack([rng.uniform(0, 640, 300), rng.uniform(0, 480, 300)], 1).astype(np.float32)
    f = lib.Frontend(480, 640, 512)
    rays = f.lift_projective(CAM, uv)
    ref = lift_oracle(CAM, uv)
    assert np.abs(rays - ref).max() <= 1e-14
    # independent check: OpenCV's iterative undistortPoints agrees near the centre (it iterates differently, so only ~1e-6)
    K = np.array([[CAM[0], 0, CAM[2]], [0, CAM[1], CAM[3]], [0, 0, 1]])
    und = cv2.undistortPoints(uv.reshape(-1, 1, 2).astype(np.float64), K, CAM[4:8]).reshape(-1, 2)
    central = np.hypot(uv[:, 0] - CAM[2], uv[:, 1] - CAM[3]) < 150
    assert np.abs(rays[central, :2] - und[central]).max() < 1e-4
    nod = f.lift_projective(np.r_[CAM[:4], 0, 0, 0, 0], uv)
    assert np.allclose(nod[:, 0], (uv[:, 0].astype(np.float64) - CAM[2]) / CAM[0], atol=1e-14)
    assert f.lift_projective(CAM, np.zeros((0, 2), np.float32)).shape == (0, 3)
    f.close()


def ref_set_mask(pts, cnt, rows, cols, radius):
    """FeatureTracker::setMask (:36-69) with cv2.circle; ties in track_cnt keep input order (std::sort leaves them unspecified)."""
    mask = np.full((rows, cols), 255, np.uint8)
    order = sorted(range(len(pts)), key=lambda i: -cnt[i])
    keep = []
    for i in order:
        px, py = int(np.rint(pts[i][0])), int(np.rint(pts[i][1]))
        if mask[py, px] == 255:
            keep.append(i)
            cv2.circle(mask, (px, py), radius, 0, -1)
    return keep, mask


@pytest.mark.parametrize("radius", [30, 7, 1])
def test_set_mask_matches_cv_circle(radius):
    from mvil_fusion_b200 import lib
    rng = np.random.default_rng(11)
    pts = np.stack([rng.uniform(1, 638, 200), rng.uniform(1, 478, 200)], 1).astype(np.float32)
    pts[:6] = [[1.2, 1.4], [638.4, 477.6], [320.5, 240.5], [321.5, 241.5], [5.0, 470.0], [630.0, 3.0]]
    cnt = rng.integers(1, 30, 200).astype(np.int32)
    f = lib.Frontend(480, 640, 512)
    keep = f.set_mask(pts, cnt, radius)
    kref, mref = ref_set_mask(pts, cnt, 480, 640, radius)
    assert list(keep) == kref
    assert np.array_equal(f.get_mask(), mref)
    f.close()


@pytest.mark.parametrize("seed,with_mask", [(1, False), (2, True), (3, False), (5, True)])
def test_good_features_matches_opencv(seed, with_mask):
    from mvil_fusion_b200 import lib
    img = cv2.createCLAHE(3.0, (8, 8)).apply(texture(seed))
    f = lib.Frontend(480, 640, 512)
    mask = None
    if with_mask:
        rng = np.random.default_rng(seed)
        old = np.stack([rng.uniform(5, 634, 60), rng.uniform(5, 474, 60)], 1).astype(np.float32)
        keep = f.set_mask(old, np.ones(60, np.int32), 30)
        _, mask = ref_set_mask(old, np.ones(60, np.int32), 480, 640, 30)
        assert len(keep) > 20
    ref = cv2.goodFeaturesToTrack(img, 150, 0.01, 30, mask=mask)
    ref = np.zeros((0, 2), np.float32) if ref is None else ref.reshape(-1, 2)
    out = f.good_features(img, 150, 0.01, 30.0, use_mask=with_mask)
    eref = cv2.cornerMinEigenVal(img, 3, ksize=3)
    e = f.eig()
    assert np.abs(e - eref).max() <= 1e-6 * eref.max()      # same formula; OpenCV's SIMD summation order of the 3x3 box differs in the last bits (cancellation in a+c-sqrt)
    sref = {tuple(p) for p in ref.astype(int)}; sout = {tuple(p) for p in out.astype(int)}
    assert len(out) >= 0.97 * len(ref) and len(sref & sout) >= 0.97 * len(sref), (len(ref), len(out), len(sref & sout))
    assert (out == np.floor(out)).all()
    # pairwise minimum distance and mask are honoured exactly
    d = np.linalg.norm(out[:, None] - out[None], axis=2) + 1e9 * np.eye(len(out))
    assert d.min() >= 30.0
    if with_mask:
        assert all(mask[int(y), int(x)] == 255 for x, y in out)
    # the strongest corners come out in OpenCV's order
    k = min(20, len(ref), len(out))
    assert np.array_equal(out[:k], ref[:k])
    f.close()


def test_good_features_small_and_flat():
    from mvil_fusion_b200 import lib
    f = lib.Frontend(64, 80, 64)
    flat = np.full((64, 80), 77, np.uint8)
    assert len(f.good_features(flat, 10, 0.01, 5.0)) == 0
    img = texture(9, 64, 80)
    ref = cv2.goodFeaturesToTrack(img, 10, 0.01, 5).reshape(-1, 2)
    out = f.good_features(img, 10, 0.01, 5.0)
    assert len({tuple(p) for p in ref.astype(int)} & {tuple(p) for p in out.astype(int)}) >= len(ref) - 1
    f.close()


def two_view(seed, n=150, outlier_frac=0.2, noise=0.1):
    """Static scene seen from two poses of a virtual pinhole (f = 460, centre of a 640 x 480 image): what rejectWithF feeds to RANSAC."""
    rng = np.random.default_rng(seed)
    X = np.stack([rng.uniform(-6, 6, n), rng.uniform(-4, 4, n), rng.uniform(4, 20, n)], 1)
    ang = rng.normal(0, 0.03, 3); t = rng.normal(0, 0.25, 3)
    R, _ = cv2.Rodrigues(ang)
    K = np.array([[460.0, 0, 320], [0, 460.0, 240], [0, 0, 1]])
    p1 = (K @ X.T).T; p1 = p1[:, :2] / p1[:, 2:]
    X2 = (R @ X.T).T + t
    p2 = (K @ X2.T).T; p2 = p2[:, :2] / p2[:, 2:]
    p1 += rng.normal(0, noise, p1.shape); p2 += rng.normal(0, noise, p2.shape)
    out = rng.random(n) < outlier_frac
    p2[out] += rng.uniform(-60, 60, (out.sum(), 2)) + np.sign(rng.normal(size=(out.sum(), 2))) * 8
    return p1.astype(np.float32), p2.astype(np.float32), out


@pytest.mark.parametrize("seed,noise,agree", [(1, 0.1, 0.9), (2, 0.1, 0.9), (3, 0.1, 0.9), (4, 0.3, 0.8)])
def test_reject_with_f_matches_opencv_ransac(seed, noise, agree):
    """cv::findFundamentalMat(..., FM_RANSAC, 1.0, 0.99, status) (feature_tracker.cpp:191).  OpenCV's sampling is random (7-point sets from
    its own RNG), so the bar is agreement of the masks, not identity: with 0.1 px noise (labels unambiguous at the 1 px gate) >= 90 % of the
    points get the same label (OpenCV stops at 0.99 confidence and may miss true inliers), >= 97 % of the TRUE inliers are kept; with 0.3 px noise borderline points may flip (>= 80 %).  Gross outliers are rejected, OpenCV's inliers kept."""
    from mvil_fusion_b200 import lib
    p1, p2, is_out = two_view(seed, noise=noise)
    Fcv, mcv = cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, 1.0, 0.99)
    mcv = mcv.reshape(-1).astype(bool)
    f = lib.Frontend(480, 640, 512)
    st, F = f.reject_with_f(p1, p2, 1.0)
    assert (st == mcv).mean() >= agree, ((st == mcv).mean(), st.sum(), mcv.sum())
    assert st[is_out].mean() <= 0.25                     # gross outliers go (one that lands within 1 px of its epipolar line is a legitimate inlier)
    if noise <= 0.1:
        assert st[~is_out].mean() >= 0.97                # true inliers stay
    assert st[mcv].mean() >= (0.9 if noise <= 0.1 else 0.8)   # OpenCV's inliers stay
    # the returned F really is the epipolar geometry of the kept points (symmetric epipolar distance <= 1 px, as computeError defines it)
    h1 = np.c_[p1, np.ones(len(p1))].astype(np.float64); h2 = np.c_[p2, np.ones(len(p2))].astype(np.float64)
    l2 = h1 @ F.T; l1 = h2 @ F
    d2 = (np.sum(h2 * l2, 1) ** 2) / (l2[:, 0] ** 2 + l2[:, 1] ** 2); d1 = (np.sum(h1 * l1, 1) ** 2) / (l1[:, 0] ** 2 + l1[:, 1] ** 2)
    assert np.array_equal(np.maximum(d1, d2) <= 1.0, st)
    st2, _ = f.reject_with_f(p1, p2, 1.0)
    assert np.array_equal(st, st2)                        # deterministic
    few, _ = f.reject_with_f(p1[:5], p2[:5], 1.0)
    assert few.all()
    f.close()


def test_triangulate_matches_svd_restatement():
    """vils_triangulate against a numpy restatement of FeatureManager::triangulate (feature_manager.cpp:214-268): np.linalg.svd plays
    Eigen::JacobiSVD.  Well-conditioned tracks agree to 1e-9 relative; a zero-parallax track and a behind-the-camera solution fall back to
    INIT_DEPTH exactly like the reference (:262-266)."""
    from mvil_fusion_b200 import lib, synth
    rng = np.random.default_rng(5)
    N = 10
    Ps = np.cumsum(rng.normal(0, 0.15, (N, 3)), 0)
    def rot(v):
        R, _ = cv2.Rodrigues(np.asarray(v, np.float64)); return R
    Rs = np.stack([rot(rng.normal(0, 0.05, 3)) for _ in range(N)])
    ric = rot([0.01, -0.02, 0.015]); tic = np.array([0.05, -0.02, 0.01])
    start, off, pts, truth = [], [0], [], []
    for f in range(120):
        s = int(rng.integers(0, N - 3)); L = int(rng.integers(2, N - s + 1))
        depth = rng.uniform(2, 15); b = np.array([rng.uniform(-0.5, 0.5), rng.uniform(-0.4, 0.4), 1.0])
        Xc0 = b * depth                                   # in the anchor camera
        Xw = Rs[s] @ (ric @ Xc0 + tic) + Ps[s]
        for j in range(s, s + L):
            Xc = ric.T @ (Rs[j].T @ (Xw - Ps[j]) - tic)
            pts.append(Xc / Xc[2] + np.r_[rng.normal(0, 1e-4, 2), 0])
        start.append(s); off.append(len(pts)); truth.append(depth)
    # zero parallax (identical poses) and a track whose bearings point backwards
    start.append(0); pts += [np.array([0.1, 0.1, 1.0])] * 2; off.append(len(pts))
    pts_arr = np.array(pts)

    def oracle(k):
        i0 = start[k]; rows = []
        t0 = Ps[i0] + Rs[i0] @ tic; R0 = Rs[i0] @ ric
        for j in range(off[k + 1] - off[k]):
            t1 = Ps[i0 + j] + Rs[i0 + j] @ tic; R1 = Rs[i0 + j] @ ric
            t = R0.T @ (t1 - t0); R = R0.T @ R1
            P = np.c_[R.T, -R.T @ t]
            fv = pts_arr[off[k] + j] / np.linalg.norm(pts_arr[off[k] + j])
            rows += [fv[0] * P[2] - fv[2] * P[0], fv[1] * P[2] - fv[2] * P[1]]
        V = np.linalg.svd(np.array(rows))[2][-1]
        d = V[2] / V[3]
        return d if d >= 0 else 5.0

    PsZ = Ps.copy(); RsZ = Rs.copy()
    out = lib.triangulate(start, off, pts_arr, Ps, Rs.reshape(N, 9), tic, ric.reshape(9), 5.0)
    ref = np.array([oracle(k) for k in range(len(start))])
    good = np.arange(120)
    assert np.abs(out[good] - ref[good]).max() / np.abs(ref[good]).max() <= 1e-9
    assert np.median(np.abs(out[good] - np.array(truth)) / np.array(truth)) < 0.05      # and it is the right depth
    assert out[120] == 5.0 or abs(out[120] - ref[120]) <= 1e-6 * abs(ref[120])          # degenerate: INIT_DEPTH or the same null vector
    assert lib.triangulate([], [0], np.zeros((0, 3)), Ps, Rs.reshape(N, 9), tic, ric.reshape(9)).shape == (0,)


def test_good_features_many_corners_takes_the_full_sort_path():
    """More corners than the top-K pre-selection holds (the greedy pick runs out of sorted candidates): the library falls back to sorting
    every candidate; the result must still be OpenCV's."""
    from mvil_fusion_b200 import lib
    img = cv2.createCLAHE(3.0, (8, 8)).apply(texture(12))
    f = lib.Frontend(480, 640, 512)
    ref = cv2.goodFeaturesToTrack(img, 6000, 0.01, 3).reshape(-1, 2)
    out = f.good_features(img, 6000, 0.01, 3.0)
    assert len(ref) > 4500                                              # really beyond the pre-selection
    sref = {tuple(p) for p in ref.astype(int)}; sout = {tuple(p) for p in out.astype(int)}
    assert abs(len(out) - len(ref)) <= 0.02 * len(ref) and len(sref & sout) >= 0.97 * len(sref)
    assert np.array_equal(out[:50], ref[:50])
    f.close()
