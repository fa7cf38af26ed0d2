"""CPU checks that pin oracle/vgicp_oracle.py (the numpy restatement of fast_gicp::FastVGICP): the gradient convention of linearize
(b = half the derivative of the error sum along the left perturbation used by step_lm), symmetry / definiteness of H, and recovery of a
known rigid motion.  The compiled fast_gicp is not available (prebuilt .a absent from the reference tree): parity unpinned, see the module header."""
import os
import sys

import numpy as np
import pytest

pytest.importorskip("scipy.spatial")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import vgicp_oracle as vo  # noqa: E402


def small_pair(seed=3, n=1200):
    rng = np.random.default_rng(seed)
    tgt = vo.room_scan(rng, n)
    T = np.eye(4); T[:3, :3] = vo.so3_exp(np.array([0.004, -0.006, 0.02])); T[:3, 3] = [0.12, -0.08, 0.03]     # source -> target
    src = vo.room_scan(rng, n, pose=T)
    return src, tgt, T


def test_linearize_gradient_matches_central_differences():
    src, tgt, T = small_pair()
    g = vo.FastVGICP(0.5); g.set_input(src, tgt)
    T0 = np.eye(4)
    err, H, b = g.linearize(T0)
    assert len(g.corr) > 0.5 * len(src)
    assert np.allclose(H, H.T, rtol=0, atol=1e-9 * np.abs(H).max()) and np.linalg.eigvalsh(H).min() > 0
    for k in range(6):
        d = np.zeros(6); d[k] = 1e-6
        Dp = np.eye(4); Dp[:3, :3] = vo.so3_exp(d[:3]); Dp[:3, 3] = d[3:]
        Dm = np.eye(4); Dm[:3, :3] = vo.so3_exp(-d[:3]); Dm[:3, 3] = -d[3:]
        fd = (g.compute_error(Dp @ T0) - g.compute_error(Dm @ T0)) / 2e-6
        assert abs(fd - 2 * b[k]) <= 1e-5 * max(1.0, abs(2 * b[k])), (k, fd, 2 * b[k])


def test_align_recovers_known_motion():
    src, tgt, T = small_pair()
    g = vo.FastVGICP(0.5); g.set_input(src, tgt)
    X = g.align()
    assert g.converged and g.nr_iterations < 30
    assert np.abs(X[:3, 3] - T[:3, 3]).max() < 0.03
    assert np.abs(X[:3, :3] - T[:3, :3]).max() < 5e-3
    f1 = g.fitness_score()
    g.final = np.eye(4)
    assert f1 < g.fitness_score()            # sparse synthetic scans: the score is dominated by the sampling density, but alignment lowers it


def test_voxel_map_is_additive_mean():
    src, tgt, _ = small_pair(n=600)
    covs, idx = vo.calculate_covariances(tgt[:, :3])
    assert (idx[:, 0] == np.arange(len(tgt))).all()
    w = np.linalg.eigvalsh(covs)
    assert np.allclose(w, [1e-3, 1.0, 1.0], atol=1e-9)
    vox = vo.create_voxelmap(tgt[:, :3], covs, 0.5)
    assert sum(v[2] for v in vox.values()) == len(tgt)
    c, v = next(iter(vox.items()))
    members = [i for i, p in enumerate(tgt[:, :3]) if vo.voxel_coord(p, 0.5) == c]
    assert members[0] == v[3] and np.allclose(v[0], tgt[members, :3].astype(np.float64).mean(0))
