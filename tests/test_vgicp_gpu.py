"""vils_vgicp_* (voxelised GICP scan matching, the producer of the LidarICPConstraint measurement) against oracle/vgicp_oracle.py, the
numpy restatement of fast_gicp::FastVGICP as estimator.cpp:263-303 configures it.  Bars: identical neighbour sets (up to exact distance
ties), covariances / voxel statistics 1e-9, H / b / error 1e-9 relative, the same LM iteration count, final transformation 1e-7,
fitness score 1e-6 relative; and recovery of the true motion of a full-size 28.8 k-point scan pair."""
import os
import sys

import numpy as np
import pytest

pytest.importorskip("scipy.spatial")
pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import vgicp_oracle as vo  # noqa: E402


def pair(seed, n, rot=(0.004, -0.006, 0.02), trans=(0.12, -0.08, 0.03)):
    rng = np.random.default_rng(seed)
    tgt = vo.room_scan(rng, n)
    T = np.eye(4); T[:3, :3] = vo.so3_exp(np.array(rot)); T[:3, 3] = trans
    src = vo.room_scan(rng, n, pose=T)
    return src, tgt, T


def test_covariances_match_restatement():
    from mvil_fusion_b200 import lib
    _, tgt, _ = pair(11, 3000)
    cov, nn = lib.vgicp_covariances(tgt)
    cov_o, nn_o = vo.calculate_covariances(tgt[:, :3])
    same = np.array([set(a) == set(b) for a, b in zip(nn, nn_o)])
    assert same.mean() >= 0.999
    assert (nn[:, 0] == np.arange(len(tgt))).all()
    assert np.abs(cov[same] - cov_o[same]).max() <= 1e-9


@pytest.mark.parametrize("search", [1, 7])
def test_linearize_and_voxel_map_match_restatement(search):
    from mvil_fusion_b200 import lib
    src, tgt, _ = pair(12, 3000)
    T0 = np.eye(4); T0[:3, :3] = vo.so3_exp(np.array([0.001, 0.002, -0.003])); T0[:3, 3] = [0.02, 0.01, -0.01]
    g = vo.FastVGICP(0.5, search); g.set_input(src, tgt)
    err, H, b = g.linearize(T0)
    r = lib.vgicp_linearize(src, tgt, T0, lib.vgicp_opts(0.5, search))
    assert r["n_voxels"] == len(g.vox) == len(r["voxels"])
    for v in g.vox.values():
        m, c, num = r["voxels"][v[3]]
        assert num == v[2] and np.abs(m - v[0]).max() <= 1e-12 and np.abs(c - v[1]).max() <= 1e-9
    assert r["n_corr"] == len(g.corr)
    assert abs(r["error"] - err) <= 1e-9 * abs(err)
    assert np.abs(r["H"] - H).max() <= 1e-9 * np.abs(H).max()
    assert np.abs(r["b"] - b).max() <= 1e-9 * max(1.0, np.abs(b).max())


def test_align_matches_restatement_and_recovers_motion():
    from mvil_fusion_b200 import lib
    src, tgt, T = pair(13, 3000)
    g = vo.FastVGICP(0.5); g.set_input(src, tgt)
    X = g.align()
    r = lib.vgicp_align(src, tgt, None, lib.vgicp_opts(0.5))
    assert r["converged"] and g.converged
    assert r["iterations"] == g.nr_iterations and r["n_linearize"] == g.n_linearize
    assert np.abs(r["T"] - X).max() <= 1e-7
    assert np.abs(r["H"] - g.final_hessian).max() <= 1e-7 * np.abs(g.final_hessian).max()
    assert abs(r["error"] - g.last_error) <= 1e-7 * abs(g.last_error)
    f = g.fitness_score()
    assert abs(r["fitness"] - f) <= 1e-6 * f
    assert np.abs(r["T"][:3, 3] - T[:3, 3]).max() < 0.03 and np.abs(r["T"][:3, :3] - T[:3, :3]).max() < 5e-3
    # a guess (as PredictRelative_rt provides, estimator.cpp:285-293) is honoured and gives the same optimum
    r2 = lib.vgicp_align(src, tgt, T.astype(np.float32).astype(np.float64), lib.vgicp_opts(0.5))
    assert r2["converged"] and r2["iterations"] <= r["iterations"]
    assert np.abs(r2["T"][:3, 3] - r["T"][:3, 3]).max() < 5e-3
    # bit-reproducible run to run
    r3 = lib.vgicp_align(src, tgt, None, lib.vgicp_opts(0.5))
    assert (r3["T"] == r["T"]).all() and r3["error"] == r["error"]


def test_full_size_scan_pair():
    from mvil_fusion_b200 import lib
    src, tgt, T = pair(14, 28800, rot=(0.01, -0.004, 0.03), trans=(0.25, 0.1, -0.02))
    r = lib.vgicp_align(src, tgt, None, lib.vgicp_opts(0.5))
    print(f"vgicp 28.8k x 28.8k: {r['elapsed_ms']:.2f} ms, {r['n_linearize']} linearisations, {r['n_voxels']} voxels, {r['n_corr']} correspondences, fitness {r['fitness']:.4g}")
    assert r["converged"]
    assert np.abs(r["T"][:3, 3] - T[:3, 3]).max() < 0.01 and np.abs(r["T"][:3, :3] - T[:3, :3]).max() < 2e-3


def test_bad_arguments():
    from mvil_fusion_b200 import cabi, lib
    src, tgt, _ = pair(15, 600)
    with pytest.raises(lib.VilsError) as e:
        lib.vgicp_align(src[:10], tgt, None, lib.vgicp_opts(0.5))
    assert e.value.code == cabi.VILS_ERR_BAD_ARG
    with pytest.raises(lib.VilsError):
        lib.vgicp_align(src, tgt, None, lib.vgicp_opts(0.5, 5))
    with pytest.raises(lib.VilsError):
        lib.vgicp_align(src, tgt, None, lib.vgicp_opts(0.5, 1, k_correspondences=10))


def test_source_partly_outside_target_grid():
    """A pose that pushes a good part of the source outside the target's bounding box: voxel lookups beyond the grid find nothing (as
    the reference's hash map), DIRECT27 offsets reach back in, and the fitness 1-NN starts from the first shell that touches the grid."""
    from mvil_fusion_b200 import lib
    src, tgt, _ = pair(16, 3000)
    T0 = np.eye(4); T0[:3, :3] = vo.so3_exp(np.array([0.0, 0.0, 0.2])); T0[:3, 3] = [3.3, -2.1, 0.7]
    T0 = T0.astype(np.float32).astype(np.float64)
    g = vo.FastVGICP(0.5, 27); g.set_input(src, tgt)
    err, H, b = g.linearize(T0)
    r = lib.vgicp_linearize(src, tgt, T0, lib.vgicp_opts(0.5, 27))
    assert 0 < len(g.corr) and r["n_corr"] == len(g.corr)
    assert abs(r["error"] - err) <= 1e-9 * abs(err)
    assert np.abs(r["H"] - H).max() <= 1e-9 * np.abs(H).max() and np.abs(r["b"] - b).max() <= 1e-9 * np.abs(b).max()
    g1 = vo.FastVGICP(0.5, 1); g1.src, g1.tgt, g1.src_cov, g1.vox = g.src, g.tgt, g.src_cov, g.vox
    n1 = g1.update_correspondences(T0)
    assert n1 < 0.9 * len(src)                                   # the pose really leaves points without a voxel
    ra = lib.vgicp_align(src, tgt, T0, lib.vgicp_opts(0.5, 1, max_iterations=0))
    assert (ra["T"] == T0).all() and ra["iterations"] == 0 and not ra["converged"] and (ra["H"] == np.eye(6)).all()
    g.final = T0
    f = g.fitness_score()
    assert abs(ra["fitness"] - f) <= 1e-6 * f


def test_capacity_and_non_finite_inputs():
    from mvil_fusion_b200 import cabi, lib
    src, tgt, _ = pair(17, 600)
    far = tgt.copy(); far[5, :3] = [4.0e6, 0.0, 0.0]               # one stray return: the dense voxel grid would not fit
    with pytest.raises(lib.VilsError) as e:
        lib.vgicp_align(src, far, None, lib.vgicp_opts(0.5))
    assert e.value.code == cabi.VILS_ERR_CAPACITY
    nan = src.copy(); nan[7, 1] = np.nan
    with pytest.raises(lib.VilsError) as e:
        lib.vgicp_align(nan, tgt, None, lib.vgicp_opts(0.5))
    assert e.value.code == cabi.VILS_ERR_NOT_FINITE
    r = lib.vgicp_align(src, tgt, None, lib.vgicp_opts(0.5))      # the library is still usable afterwards
    assert r["n_voxels"] > 0
