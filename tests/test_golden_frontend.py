"""Committed golden vectors (tests/golden/frontend_golden.npz, made by tests/golden/make_frontend_golden.py with cv2 4.13.0): results of
the OpenCV calls of FeatureTracker::readImage (feature_tracker_/src/feature_tracker.cpp:87-93,113,149) on seeded images that are
regenerated here WITHOUT OpenCV.  The GPU test needs no cv2 at run time; the CPU test re-derives the fixture when cv2 is present."""
import importlib.util
import os
import zlib

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_frontend_golden", os.path.join(HERE, "golden", "make_frontend_golden.py"))


def _gen():
    try:
        import cv2  # noqa: F401  (the generator imports it at module level)
    except Exception:
        return None
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m


def _texture(seed, rows=480, cols=640):
    rng = np.random.default_rng(seed)
    img = rng.uniform(0, 255, (rows + 8, cols + 8))
    c = np.pad(np.cumsum(np.cumsum(img, 0), 1), ((1, 0), (1, 0)))
    k = 8
    box = ((c[k:, k:] - c[:-k, k:] - c[k:, :-k] + c[:-k, :-k]) / (k * k))[:rows, :cols]
    return np.clip((box - box.min()) / (box.max() - box.min()) * 255.0, 0, 255).astype(np.uint8)


GOLD = np.load(os.path.join(HERE, "golden", "frontend_golden.npz"))


def test_golden_fixture_is_what_cv2_produces():
    """CPU: the committed fixture equals a fresh run of the generator (guards against a stale or hand-edited file)."""
    m = _gen()
    if m is None:
        pytest.skip("cv2 not available")
    import cv2
    assert np.array_equal(m.texture(101), _texture(101))
    clahe = cv2.createCLAHE(3.0, (8, 8))
    for seed in (101, 102):
        eq = clahe.apply(m.texture(seed))
        assert zlib.crc32(eq.tobytes()) == int(GOLD[f"clahe_crc_{seed}"])
        assert np.array_equal(cv2.goodFeaturesToTrack(eq, 150, 0.01, 30).reshape(-1, 2), GOLD[f"gftt_{seed}"])


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [101, 102])
def test_frontend_against_golden(seed):
    from mvil_fusion_b200 import lib
    raw = _texture(seed)
    f = lib.Frontend(480, 640, 512)
    eq = f.clahe(raw, 3.0, (8, 8))
    assert zlib.crc32(eq.tobytes()) == int(GOLD[f"clahe_crc_{seed}"])                # CLAHE: bit-exact
    assert np.array_equal(eq[::97, ::53], GOLD[f"clahe_rows_{seed}"])
    gold_pts = GOLD[f"gftt_{seed}"]
    pts = f.good_features(eq, 150, 0.01, 30.0)
    common = {tuple(p) for p in pts.astype(int)} & {tuple(p) for p in gold_pts.astype(int)}
    assert len(common) >= 0.97 * len(gold_pts) and np.array_equal(pts[:20], gold_pts[:20])
    nxt = f.clahe(np.roll(np.roll(raw, -3, 0), 5, 1), 3.0, (8, 8))
    k = lib.KLT(480, 640, 512, 21, 3)
    out, st, err = k.track(eq, nxt, gold_pts)
    assert np.array_equal(st, GOLD[f"klt_status_{seed}"])
    ok = st == 1
    assert np.abs(out[ok] - GOLD[f"klt_pts_{seed}"][ok]).max() <= 0.02
    assert np.abs(err[ok] - GOLD[f"klt_err_{seed}"][ok]).max() <= 1e-3 * max(1.0, float(GOLD[f"klt_err_{seed}"][ok].max()))
    k.close(); f.close()
