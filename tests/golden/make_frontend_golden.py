"""Generates tests/golden/frontend_golden.npz: outputs of the OpenCV calls the reference's FeatureTracker::readImage makes
(feature_tracker_/src/feature_tracker.cpp:87-93,113,149: CLAHE(3.0, 8x8), calcOpticalFlowPyrLK(21x21, maxLevel 3),
goodFeaturesToTrack(150, 0.01, 30)) on seeded synthetic images, produced HERE by cv2 4.13.0 — the third-party library whose
results the GPU front end has to reproduce.  Small by construction: images are regenerated from the seed, only results are stored
(corner lists, tracked points, a CRC per CLAHE image).

    python tests/golden/make_frontend_golden.py
"""
import os
import zlib

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def texture(seed, rows=480, cols=640):
    """Seeded texture WITHOUT any OpenCV call, so that the fixture inputs do not depend on the cv2 build: box-blurred uniform noise."""
    rng = np.random.default_rng(seed)
    img = rng.uniform(0, 255, (rows + 8, cols + 8))
    c = np.cumsum(np.cumsum(img, 0), 1)
    c = np.pad(c, ((1, 0), (1, 0)))
    k = 8
    box = (c[k:, k:] - c[:-k, k:] - c[k:, :-k] + c[:-k, :-k]) / (k * k)
    box = box[:rows, :cols]
    return np.clip((box - box.min()) / (box.max() - box.min()) * 255.0, 0, 255).astype(np.uint8)


def shifted(img, dx, dy):
    """Integer shift with edge replication (no interpolation: independent of cv2)."""
    out = np.roll(np.roll(img, dy, 0), dx, 1)
    return out


def main():
    out = {"cv2_version": np.array(cv2.__version__)}
    clahe = cv2.createCLAHE(3.0, (8, 8))
    for seed in (101, 102):
        raw = texture(seed)
        eq = clahe.apply(raw)
        out[f"clahe_crc_{seed}"] = np.array(zlib.crc32(eq.tobytes()), np.uint64)
        out[f"clahe_rows_{seed}"] = eq[::97, ::53].copy()                      # a sparse sample of the equalised image
        pts = cv2.goodFeaturesToTrack(eq, 150, 0.01, 30).reshape(-1, 2)
        out[f"gftt_{seed}"] = pts.astype(np.float32)
        nxt = clahe.apply(shifted(raw, 5, -3))
        p1, st, err = cv2.calcOpticalFlowPyrLK(eq, nxt, pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3)
        out[f"klt_pts_{seed}"] = p1.reshape(-1, 2).astype(np.float32)
        out[f"klt_status_{seed}"] = st.reshape(-1).astype(np.uint8)
        out[f"klt_err_{seed}"] = err.reshape(-1).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "frontend_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "frontend_golden.npz"), {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
