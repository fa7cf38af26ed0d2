"""vils_lidar_associate against a numpy / scipy restatement of the scan-to-map association of lidar_mapping/src/localMapping.cpp:611-766
(cKDTree plays pcl::KdTreeFLANN, numpy.linalg.eigh the SelfAdjointEigenSolver, lstsq the colPivHouseholderQr solve).
Bars: identical neighbour sets and valid flags (up to exact distance ties), line end points / plane parameters within 1e-9."""
import numpy as np
import pytest

scipy_spatial = pytest.importorskip("scipy.spatial")
pytestmark = pytest.mark.gpu


def room_map(rng, n_edge=6000, n_surf=20000):
    """Points on the planes and plane-plane intersection lines of a 16 x 12 x 3 m room, 2 cm noise, with an intensity channel."""
    surf = []
    for axis, val in ((0, -8.0), (0, 8.0), (1, -6.0), (1, 6.0), (2, 0.0), (2, 3.0)):
        p = np.stack([rng.uniform(-8, 8, n_surf // 6), rng.uniform(-6, 6, n_surf // 6), rng.uniform(0, 3, n_surf // 6)], 1)
        p[:, axis] = val
        surf.append(p)
    surf = np.concatenate(surf) + rng.normal(0, 0.02, (n_surf // 6 * 6, 3))
    edge = []
    for x in (-8.0, 8.0):
        for y in (-6.0, 6.0):
            edge.append(np.stack([np.full(n_edge // 4, x), np.full(n_edge // 4, y), rng.uniform(0, 3, n_edge // 4)], 1))
    edge = np.concatenate(edge) + rng.normal(0, 0.02, (n_edge // 4 * 4, 3))
    f = lambda p: np.c_[p, rng.uniform(0, 100, len(p))].astype(np.float32)
    return f(edge), f(surf)


def qrot(q, v):
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return v @ R.T


def make_scan(rng, mp, q, t, n):
    """Scan points = map-like points moved into the sensor frame (inverse pose) + noise; a few far outliers."""
    pick = mp[rng.integers(0, len(mp), n)].astype(np.float64)
    w = pick[:, :3] + rng.normal(0, 0.03, (n, 3))
    w[: n // 20] += rng.uniform(3, 6, (n // 20, 3))                       # nothing within 1 m
    qi = np.array([-q[0], -q[1], -q[2], q[3]])
    s = qrot(qi, w - t)
    return np.c_[s, pick[:, 3] + rng.normal(0, 5, n)].astype(np.float32)


def oracle(mp, scan, q, t, mode):
    tree = scipy_spatial.cKDTree(mp[:, :3].astype(np.float64))
    sel = (qrot(q, scan[:, :3].astype(np.float64)) + t).astype(np.float32).astype(np.float64)
    K = 5 if mode == 0 else 10
    dist, idx = tree.query(sel, k=K)
    out = np.zeros((len(scan), 10)); valid = np.zeros(len(scan), bool); nn = np.zeros((len(scan), 5), np.int64); cand = np.zeros(len(scan), bool)
    for i in range(len(scan)):
        out[i, :3] = scan[i, :3]
        d5 = np.float32(np.sum((mp[idx[i, 4], :3] - sel[i].astype(np.float32)) ** 2, dtype=np.float32))
        if mode == 0:
            use = idx[i, :5]
        else:
            diff = np.abs(mp[idx[i], 3] - scan[i, 3]).astype(np.float32)
            order = sorted(range(K), key=lambda k: (diff[k], idx[i, k]))
            use = idx[i, order[:5]]
        nn[i] = use
        if not d5 < 1.0:
            continue
        cand[i] = True
        P = mp[use, :3].astype(np.float64)
        if mode == 0:
            c = P.sum(0) / 5.0
            C = (P - c).T @ (P - c)
            w, V = np.linalg.eigh(C)
            if w[2] > 3 * w[1]:
                valid[i] = True
                out[i, 3:6] = 0.1 * V[:, 2] + c; out[i, 6:9] = -0.1 * V[:, 2] + c
        else:
            n = np.linalg.lstsq(P, -np.ones(5), rcond=None)[0]
            d = 1.0 / np.linalg.norm(n); n = n / np.linalg.norm(n)
            if np.all(np.abs(P @ n + d) <= 0.2):
                valid[i] = True
                out[i, 3:6] = n; out[i, 6] = d
    return out, valid, nn, cand


@pytest.mark.parametrize("mode", [0, 1])
def test_associate_matches_restatement(mode):
    from mvil_fusion_b200 import lib
    rng = np.random.default_rng(31 + mode)
    edge, surf = room_map(rng)
    mp = edge if mode == 0 else surf
    q = np.array([0.01, -0.02, 0.15, 1.0]); q /= np.linalg.norm(q); t = np.array([0.7, -0.4, 0.1])
    scan = make_scan(rng, mp, q, t, 1500)
    out, valid, nn, ms = lib.lidar_associate(mp, scan, q, t, mode)
    ro, rv, rn, cand = oracle(mp, scan, q, t, mode)
    same_nn = np.array([set(a) == set(b) for a, b in zip(nn, rn)])
    assert not valid[~cand].any()                                          # nothing within 1 m: no residual block (neighbour order there is a float tie-break)
    assert same_nn[cand].mean() >= 0.995                                   # exact search; only float distance ties may differ
    agree = same_nn & (valid == rv)
    assert agree[cand].mean() >= 0.99, (agree.mean(), valid.sum(), rv.sum())
    assert 0.3 * len(scan) < valid.sum() < 0.97 * len(scan)               # the far points and the bad fits are rejected
    ok = agree & valid
    assert np.abs(out[ok, :3] - ro[ok, :3]).max() == 0.0
    if mode == 0:
        a, b, ra, rb = out[ok, 3:6], out[ok, 6:9], ro[ok, 3:6], ro[ok, 6:9]
        d = np.minimum(np.abs(a - ra).max(1) + np.abs(b - rb).max(1), np.abs(a - rb).max(1) + np.abs(b - ra).max(1))   # eigenvector sign is free
        assert d.max() <= 1e-9
    else:
        assert np.abs(out[ok, 3:7] - ro[ok, 3:7]).max() <= 1e-9
    assert ms > 0


def test_associate_sparse_rim_is_still_exact():
    """Surface points with only 6 map points within 1 m: the 10 nearest include points 1-2 m away (5 x 5 x 5 shell of the grid) or further than
    2 m (exhaustive fallback).  The neighbour sets must still be the exact 10-NN -> 5 by intensity of the k-d tree."""
    from mvil_fusion_b200 import lib
    rng = np.random.default_rng(77)
    filler = np.c_[rng.uniform(200, 260, (6000, 3)), rng.uniform(0, 100, 6000)]            # far away: makes the map large enough for the grid
    sites, pts = [], []
    for k in range(48):
        c = np.array([12.0 * (k % 8), 12.0 * (k // 8), 0.37 * k])
        sites.append(c)
        near = c + rng.normal(0, 0.12, (6, 3))
        r = rng.uniform(1.15, 1.8, 7) if k % 2 == 0 else rng.uniform(2.4, 3.4, 7)          # even sites: shell, odd sites: exhaustive search
        u = rng.normal(0, 1, (7, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
        pts.append(np.c_[np.vstack([near, c + r[:, None] * u]), rng.uniform(0, 100, 13)])
    mp = np.vstack([filler] + pts).astype(np.float32)
    scan = np.c_[np.array(sites), rng.uniform(0, 100, len(sites))].astype(np.float32)
    q = np.array([0, 0, 0, 1.0]); t = np.zeros(3)
    out, valid, nn, _ = lib.lidar_associate(mp, scan, q, t, 1)
    ro, rv, rn, cand = oracle(mp, scan, q, t, 1)
    assert cand.all()
    assert all(set(a) == set(b) for a, b in zip(nn, rn))
    assert np.array_equal(valid.astype(bool), rv)
    ok = rv
    if ok.any():
        assert np.abs(out[ok, 3:7] - ro[ok, 3:7]).max() <= 1e-9


def test_associate_edge_cases():
    from mvil_fusion_b200 import lib
    q = np.array([0, 0, 0, 1.0]); t = np.zeros(3)
    mp = np.zeros((3, 4), np.float32)                                      # fewer map points than K: nothing is valid
    scan = np.zeros((4, 4), np.float32)
    out, valid, nn, _ = lib.lidar_associate(mp, scan, q, t, 0)
    assert not valid.any()
    out, valid, nn, _ = lib.lidar_associate(mp, np.zeros((0, 4), np.float32), q, t, 1)
    assert out.shape == (0, 10)


def depth_oracle(cloud, T1, T2, feat, nb=360):
    """numpy restatement of DepthRegister::get_depth (feature_tracker_/src/feature_tracker.h:129-343), float32 where the reference is."""
    f32 = np.float32
    def aff(T, p):
        T = T.astype(f32)
        return np.stack([((T[r, 0] * p[:, 0] + T[r, 1] * p[:, 1]) + T[r, 2] * p[:, 2]) + T[r, 3] for r in range(3)], 1).astype(f32)
    p = aff(T2, aff(T1, cloud[:, :3].astype(f32)))
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        keep = ~((x < 0) | (np.abs(y / x) > 10) | (np.abs(z / x) > 10))
    bin_res = f32(180.0) / f32(nb)
    row_angle = (np.arctan2(z, np.sqrt(x * x + y * y).astype(f32)).astype(f32).astype(np.float64) * 180.0 / np.pi + 90.0).astype(f32)
    col_angle = (np.arctan2(x, y).astype(f32).astype(np.float64) * 180.0 / np.pi).astype(f32)
    def cround(v):
        fl = np.floor(v); return np.where(v - fl >= 0.5, fl + 1, fl).astype(np.int64)
    with np.errstate(invalid="ignore"):
        row = cround((row_angle / bin_res).astype(f32)); col = cround((col_angle / bin_res).astype(f32))
    keep &= (row >= 0) & (row < nb) & (col >= 0) & (col < nb)
    dist = np.sqrt(x * x + y * y + z * z).astype(f32)
    best = {}
    for i in np.nonzero(keep)[0]:
        k = (row[i], col[i])
        if k not in best or dist[i] < dist[best[k]]:
            best[k] = i
    idx = np.array([best[k] for k in sorted(best)], np.int64)
    depth = np.full(len(feat), -1.0, f32)
    if len(idx) < 10:
        return depth
    sel = p[idx]; rng_ = np.sqrt((sel * sel).sum(1)).astype(f32)
    sph = (sel / rng_[:, None]).astype(f32)
    fv = feat.astype(f32); fv = (fv / np.sqrt((fv * fv).sum(1))[:, None]).astype(f32)
    q = np.stack([fv[:, 2], -fv[:, 0], -fv[:, 1]], 1)
    tree = scipy_spatial.cKDTree(sph.astype(np.float64))
    d, nn = tree.query(q.astype(np.float64), k=3)
    thr = f32(np.power(np.sin(float(bin_res) / 180.0 * np.pi) * 5.0, 2))
    for i in range(len(feat)):
        if f32(d[i, 2] ** 2) < thr:
            r = rng_[nn[i]]
            if not (r.max() - r.min() > 2):
                dep = q[i, 0] * ((r[0] + r[1] + r[2]) / f32(3))
                if dep > 3.0:
                    depth[i] = dep
    return depth


def test_depth_register_matches_restatement():
    from mvil_fusion_b200 import lib
    rng = np.random.default_rng(77)
    _, surf = room_map(rng, 4000, 120000)                       # walls of a 16 x 12 x 3 m room, world frame
    cloud = surf.copy(); cloud[:, 2] -= 1.2                     # sensor 1.2 m above the floor
    import cv2
    R1, _ = cv2.Rodrigues(np.array([0.02, -0.01, 0.3])); t1 = np.array([0.5, -0.3, 0.1])
    T1 = np.c_[R1.T, -R1.T @ t1]                                # transNow.inverse()
    R2, _ = cv2.Rodrigues(np.array([0.01, 0.02, -0.015])); T2 = np.c_[R2, np.array([0.05, 0.0, -0.08])]
    feat = np.stack([rng.uniform(-0.9, 0.9, 150), rng.uniform(-0.65, 0.65, 150), np.ones(150)], 1)
    depth, ms = lib.depth_register(cloud, T1, T2, feat, 360)
    ref = depth_oracle(cloud, T1, T2, feat, 360)
    same = ((depth < 0) == (ref < 0))
    assert same.mean() >= 0.98, same.mean()
    both = (depth > 0) & (ref > 0)
    assert both.sum() > 60                                      # most features get a LiDAR depth in this scene
    rel = np.abs(depth[both] - ref[both]) / ref[both]
    assert (rel <= 1e-5).mean() >= 0.98 and np.median(rel) <= 1e-6
    # few points / no points: nothing is assigned
    d2, _ = lib.depth_register(cloud[:5], T1, T2, feat, 360)
    assert (d2 == -1).all()
    d3, _ = lib.depth_register(np.zeros((0, 4), np.float32), T1, T2, feat, 360)
    assert (d3 == -1).all()
