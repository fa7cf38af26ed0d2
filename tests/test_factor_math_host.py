"""The product's factor arithmetic (mvil_fusion_b200/csrc/factors.cuh, __host__ __device__) compiled for the CPU by
tests/hostcheck and compared element-wise with the oracle's Evaluate restatements.  Catches math errors without a GPU;
the -m gpu tests repeat the comparison through the kernels and the C-ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import helpers
import oracle_lib as ol
from mvil_fusion_b200 import cabi, synth

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "hostcheck", "libhostcheck.so")


@pytest.fixture(scope="module")
def hc():
    src = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", "-fPIC", "-shared", "-o", SO, src])
    lib = C.CDLL(SO)
    lib.hc_plane_eval.restype = C.c_double
    return lib


def d(a):
    a = np.ascontiguousarray(a, np.float64)
    return a.ctypes.data_as(cabi.c_double_p), a


def product_evaluate(hc, cfg, w):
    """Re-create vils_ba_evaluate(apply_loss=0) output with the product's per-factor functions."""
    N = w["pose"].shape[0]
    RLB = np.array(cfg.rlb[:]).reshape(3, 3); TLB = np.array(cfg.tlb[:])
    G = np.array(cfg.gravity[:])
    rs, Js = [], []
    pose = np.ascontiguousarray(w["pose"]); sb = np.ascontiguousarray(w["speedbias"]); ex = np.ascontiguousarray(w["ex_pose"])
    P = lambda a: a.ctypes.data_as(cabi.c_double_p)
    for k in range(len(w["imu"])):
        i = int(w["imu_kf"][k])
        pre = np.ascontiguousarray(w["imu"][k])
        W = np.zeros((15, 15)); r = np.zeros(15); J = np.zeros((15, 30))
        assert hc.hc_imu_sqrt_info(P(pre[242:]), P(W)) == 0
        hc.hc_imu_eval_raw(P(pre), P(G), P(pose[i]), P(sb[i]), P(pose[i + 1]), P(sb[i + 1]), P(r), P(J))
        r2 = np.zeros(15); J2 = np.zeros((15, 30))   # the lane-split variant the solve kernel uses must be identical
        hc.hc_imu_eval_parts(P(pre), P(G), P(pose[i]), P(sb[i]), P(pose[i + 1]), P(sb[i + 1]), P(r2), P(J2))
        assert np.array_equal(r, r2) and np.array_equal(J, J2)
        r3 = np.zeros(15); J3 = np.zeros((15, 30))   # the table-assembled (warp-cooperative) variant the kernels use
        hc.hc_imu_eval_tbl(P(pre), P(G), P(pose[i]), P(sb[i]), P(pose[i + 1]), P(sb[i + 1]), P(r3), P(J3))
        assert np.array_equal(r, r3) and np.array_equal(J, J3)
        rs.append(W @ r); Js.append((W @ J).reshape(-1))
    for k in range(len(w["kf_i"])):
        i, j, f = int(w["kf_i"][k]), int(w["kf_j"][k]), int(w["feat"][k])
        c = np.concatenate([w["pts_i"][k], w["pts_j"][k], w["vel_i"][k], w["vel_j"][k], [w["td_i"][k], w["td_j"][k], w["row_i"][k], w["row_j"][k]]])
        r = np.zeros(2); J = np.zeros(40)
        hc.hc_proj_eval(C.c_double(cfg.focal_length / 2), C.c_double(cfg.tr / cfg.row), C.c_double(cfg.row / 2), int(cfg.estimate_td),
                        P(c), P(pose[i]), P(pose[j]), P(ex), C.c_double(w["inv_depth"][f]), C.c_double(w["td"]), P(r), P(J))
        r2 = np.zeros(2); J2 = np.zeros(40)   # the pair-context variant used by the solve kernel: same formulas, re-associated
        hc.hc_proj_eval_ctx(C.c_double(cfg.focal_length / 2), C.c_double(cfg.tr / cfg.row), C.c_double(cfg.row / 2), int(cfg.estimate_td),
                            P(c), P(pose[i]), P(pose[j]), P(ex), C.c_double(w["inv_depth"][f]), C.c_double(w["td"]), P(r2), P(J2))
        assert np.abs(r2 - r).max() <= 1e-11 * max(1.0, np.abs(r).max()) and np.abs(J2 - J).max() <= 1e-11 * max(1.0, np.abs(J).max())
        # the row-vector variant used by the evaluate kernel (and its folded Cauchy corrector) against the line-by-line one
        for loss_a in (0.0, 1.0):
            r3 = np.zeros(2); J3 = np.zeros(40); r4 = np.zeros(2); J4 = np.zeros(40)
            args = (C.c_double(cfg.focal_length / 2), C.c_double(cfg.tr / cfg.row), C.c_double(cfg.row / 2), int(cfg.estimate_td),
                    P(c), P(pose[i]), P(pose[j]), P(ex), C.c_double(w["inv_depth"][f]), C.c_double(w["td"]), C.c_double(loss_a))
            hc.hc_proj_eval_rows(*args, P(r3), P(J3)); hc.hc_proj_eval_loss(*args, P(r4), P(J4))
            wgt = 1.0 if loss_a == 0.0 else np.sqrt(1.0 / (1.0 + (r @ r) / loss_a ** 2))   # corrector of marginalization_factor.cpp:49-53
            for rr, JJ in ((r3, J3), (r4, J4)):
                assert np.abs(rr - wgt * r).max() <= 1e-11 * max(1.0, np.abs(r).max()) and np.abs(JJ - wgt * J).max() <= 1e-11 * max(1.0, np.abs(J).max())
        rs.append(r); Js.append(J)
    for k in range(len(w.get("plane_kf", []))):
        pb = np.ascontiguousarray(RLB.T @ (w["plane_p"][k] - TLB)); n = np.ascontiguousarray(w["plane_n"][k]); J = np.zeros(6)
        r = hc.hc_plane_eval(P(pose[int(w["plane_kf"][k])]), P(pb), P(n), C.c_double(w["plane_d"][k]), P(J))
        rs.append(np.array([r])); Js.append(J)
    for k in range(len(w.get("edge_kf", []))):
        pb = np.ascontiguousarray(RLB.T @ (w["edge_p"][k] - TLB)); a = np.ascontiguousarray(w["edge_a"][k]); b = np.ascontiguousarray(w["edge_b"][k])
        r = np.zeros(3); J = np.zeros(18)
        hc.hc_edge_eval(P(pose[int(w["edge_kf"][k])]), P(pb), P(a), P(b), P(r), P(J))
        rs.append(r); Js.append(J)
    for cst in w.get("icp") or []:
        c = np.array([*cst["t"], *cst["trans_t"], cst["sqrt_info"]]); r = np.zeros(3); J = np.zeros(72)
        kf = cst["kf"]
        hc.hc_icp_eval(P(c), P(pose[kf[0]]), P(pose[kf[1]]), P(pose[kf[2]]), P(pose[kf[3]]), P(r), P(J))
        rs.append(r); Js.append(J)
    for cst in w.get("lps") or []:
        c = np.array([*cst["t"], *cst["q"]]); r = np.zeros(3); J = np.zeros(36)
        kf = cst["kf"]
        hc.hc_lps_eval(P(c), P(pose[kf[0]]), P(pose[kf[1]]), P(r), P(J))
        rs.append(r); Js.append(J)
    return np.concatenate(rs), np.concatenate(Js)


@pytest.mark.parametrize("use_td", [1, 0])
def test_product_factor_math_matches_oracle(hc, use_td):
    w = synth.make_window(config_id=9, window_idx=7, N=6, M=25, n_lidar=120, n_icp=3, n_lps=3)
    cfg = cabi.default_config(); cfg.estimate_td = use_td
    r_o, J_o, _ = ol.evaluate_window(cfg, w, apply_loss=False)
    r_p, J_p = product_evaluate(hc, cfg, w)
    n = len(r_p)   # the oracle appends the prior residual
    ro = jo = 0
    for fam, k, nr, blocks in helpers.factor_layout(w):
        width = sum(s for _, s in blocks)
        a, b = r_p[ro:ro + nr], r_o[ro:ro + nr]
        Ja, Jb = J_p[jo:jo + nr * width], J_o[jo:jo + nr * width]
        # IMU rows are whitened by sqrt_info (entries up to ~1e5) computed through a 15x15 inverse whose condition
        # number is ~1e9: compare relative to the row scale there, tightly elsewhere.
        tol = 1e-7 if fam == "imu" else 1e-11
        assert np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max()), (fam, k)
        assert np.abs(Ja - Jb).max() <= tol * max(1.0, np.abs(Jb).max()), (fam, k)
        ro += nr; jo += nr * width
    assert ro == n


def test_prior_dx_and_plus(hc):
    rng = np.random.default_rng(5)
    for _ in range(20):
        x0 = np.concatenate([rng.normal(size=3), synth.small_quat(rng.normal(size=3))])
        dl = rng.normal(size=6) * 0.1
        x = x0.copy()
        hc.hc_pose_plus(x.ctypes.data_as(cabi.c_double_p), dl.ctypes.data_as(cabi.c_double_p))
        np.testing.assert_allclose(x, helpers.pose_plus(x0, dl), atol=1e-15)
        dx = np.zeros(6)
        hc.hc_prior_dx_pose(x.ctypes.data_as(cabi.c_double_p), x0.ctypes.data_as(cabi.c_double_p), dx.ctypes.data_as(cabi.c_double_p))
        np.testing.assert_allclose(dx[:3], dl[:3], atol=1e-14)
        # 2 vec(q0^-1 q) of the normalised first-order update
        nrm = np.sqrt(1 + 0.25 * dl[3:] @ dl[3:])
        np.testing.assert_allclose(dx[3:], dl[3:] / nrm, atol=1e-13)
