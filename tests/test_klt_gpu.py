"""GPU pyramidal LK (vils_klt_*) against the very function the reference calls: cv2.calcOpticalFlowPyrLK (OpenCV 4.13.0),
window 21x21, maxLevel 3 (feature_tracker_/src/feature_tracker.cpp:113).  Tolerance: positions within 0.02 px (the integer
patch pipeline is bit-exact; only the FP32 accumulation order of A and b differs), status identical, err within 1e-3."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
pytestmark = pytest.mark.gpu


def make_pair(seed, shift=(-4.3, 3.2), angle=0.0, rows=480, cols=640):
    rng = np.random.default_rng(seed)
    img = rng.uniform(0, 255, (rows, cols)).astype(np.float32)
    img = cv2.GaussianBlur(img, (0, 0), 2.0)
    img = cv2.normalize(img, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)
    img = cv2.createCLAHE(3.0, (8, 8)).apply(img)                      # readImage applies CLAHE before LK (:87-93)
    M = cv2.getRotationMatrix2D((cols / 2, rows / 2), angle, 1.0); M[:, 2] += shift
    nxt = cv2.warpAffine(img, M, (cols, rows), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
    pts = cv2.goodFeaturesToTrack(img, 150, 0.01, 30).reshape(-1, 2).astype(np.float32)
    return img, nxt, pts


@pytest.mark.parametrize("seed,shift,angle", [(1, (-4.3, 3.2), 0.0), (2, (11.5, -7.25), 1.5), (3, (0.4, 0.1), 0.0), (4, (25.0, 18.0), 0.0)])
def test_klt_matches_opencv(seed, shift, angle):
    from mvil_fusion_b200 import lib
    img, nxt, pts = make_pair(seed, shift, angle)
    ref, st_ref, err_ref = cv2.calcOpticalFlowPyrLK(img, nxt, pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3)
    ref = ref.reshape(-1, 2); st_ref = st_ref.reshape(-1); err_ref = err_ref.reshape(-1)
    k = lib.KLT(480, 640, 512, 21, 3)
    out, st, err = k.track(img, nxt, pts)
    assert np.array_equal(st, st_ref)
    ok = st_ref == 1
    assert ok.sum() > 100
    d = np.abs(out[ok] - ref[ok]).max(axis=1)
    assert d.max() <= 0.02, (d.max(), np.sort(d)[-5:])
    assert np.median(d) <= 1e-3
    assert np.abs(err[ok] - err_ref[ok]).max() <= 1e-3 * max(1.0, err_ref[ok].max())
    # and it actually tracks the motion
    truth = pts + np.array(shift, np.float32)
    if angle == 0.0:
        assert np.median(np.abs(out[ok] - truth[ok])) < 0.1
    k.close()


def test_klt_border_points_and_empty():
    from mvil_fusion_b200 import lib
    img, nxt, _ = make_pair(5, (6.0, -5.0))
    pts = np.array([[2.0, 3.0], [637.5, 477.0], [320.0, 1.0], [1.0, 240.0], [100.2, 100.7], [638.9, 5.5]], np.float32)
    ref, st_ref, _ = cv2.calcOpticalFlowPyrLK(img, nxt, pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3)
    k = lib.KLT(480, 640, 64, 21, 3)
    out, st, err = k.track(img, nxt, pts)
    assert np.array_equal(st, st_ref.reshape(-1))
    ok = st == 1
    assert np.abs(out[ok] - ref.reshape(-1, 2)[ok]).max() <= 0.05
    out0, st0, _ = k.track(img, nxt, np.zeros((0, 2), np.float32))
    assert out0.shape == (0, 2)
    # strided input (cv::Mat::step > cols)
    big = np.zeros((480, 704), np.uint8); big[:, :640] = img
    big2 = np.zeros((480, 704), np.uint8); big2[:, :640] = nxt
    v1, v2 = big[:, :640], big2[:, :640]
    assert v1.strides[0] == 704 and not v1.flags["C_CONTIGUOUS"]          # the library really sees a 704-byte row pitch
    out2, st2, _ = k.track(v1, v2, pts)
    assert np.array_equal(out2, out) and np.array_equal(st2, st)
    k.close()


def test_device_resident_chain_equals_pairwise_tracking():
    """vils_frontend_load -> vils_klt_advance (the frame never leaves the device, the pyramid of frame k is reused as `prev` for frame k + 1) gives
    bit-identical tracks to the stateless pairwise call on the same equalised images, over a 4-frame sequence, and the resident corner detector
    returns the same corners as the host-buffer one."""
    from mvil_fusion_b200 import lib
    rng = np.random.default_rng(11)
    base = cv2.GaussianBlur(rng.uniform(0, 255, (480, 640)).astype(np.float32), (0, 0), 2.0)
    base = cv2.normalize(base, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)
    frames = []
    for k in range(4):
        M = cv2.getRotationMatrix2D((320, 240), 0.4 * k, 1.0); M[:, 2] += (3.1 * k, -2.2 * k)
        frames.append(cv2.warpAffine(base, M, (640, 480), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101))
    f = lib.Frontend(480, 640, 512); k1 = lib.KLT(480, 640, 512, 21, 3); k2 = lib.KLT(480, 640, 512, 21, 3)
    eq = [f.clahe(im) for im in frames]                                         # what the pairwise path sees (downloaded CLAHE output)
    dev, pitch = f.load(frames[0], True)
    k1.advance(dev, pitch, np.zeros((0, 2), np.float32), f.ready_event)                       # first frame: load only
    pts = f.good_features_resident(150, 0.01, 30.0)
    assert np.array_equal(pts, f.good_features(eq[0], 150, 0.01, 30.0))
    assert len(pts) > 100
    for k in range(1, 4):
        dev, pitch = f.load(frames[k], True)
        out, st, err = k1.advance(dev, pitch, pts, f.ready_event)
        ref, st_ref, err_ref = k2.track(eq[k - 1], eq[k], pts)
        assert np.array_equal(st, st_ref) and np.array_equal(out, ref) and np.array_equal(err, err_ref)
        pts = out[st == 1]
        assert len(pts) > 80
    f.close(); k1.close(); k2.close()
