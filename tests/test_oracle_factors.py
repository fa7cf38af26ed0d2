"""Pins the CPU oracle (the reference ships no tests): finite-difference Jacobian checks modelled on
ProjectionFactor::check (vils_estimator/src/factor/projection_factor.cpp:123-225: right-multiplicative
perturbation Q*deltaQ(eps)), zero-residual fixed points, and an independent numpy re-assembly of the normal equations."""
import numpy as np
import pytest

import helpers
import oracle_lib as ol
from mvil_fusion_b200 import cabi, synth


def small_window(**kw):
    args = dict(config_id=9, window_idx=1, N=5, M=12, n_lidar=40, n_icp=2, n_lps=2)
    args.update(kw)
    return synth.make_window(**args)


def dense_jacobian(w, J):
    N, M = w["pose"].shape[0], w["inv_depth"].shape[0]
    T = 15 * N + 7 + M
    rows = []
    fams = []
    jo = 0
    for fam, k, nr, blocks in helpers.factor_layout(w):
        width = sum(s for _, s in blocks)
        Jf = J[jo:jo + nr * width].reshape(nr, width)
        full = np.zeros((nr, T))
        c = 0
        for o, s in blocks:
            full[:, o:o + s] += Jf[:, c:c + s]
            c += s
        rows.append(full); fams += [fam] * nr
        jo += nr * width
    n = int(w.get("prior_n", 0))
    if n:   # MarginalizationFactor: constant Jacobian = column slices of J_lin (marginalization_factor.cpp:384-398)
        full = np.zeros((n, T))
        full[:, helpers.prior_columns(w)] = np.asarray(w["prior_J"]).reshape(n, n).T
        rows.append(full); fams += ["prior"] * n
    return np.vstack(rows), np.array(fams)


def perturbed(w, col, eps, global_quat=False):
    """Window with tangent coordinate `col` moved by eps."""
    N, M = w["pose"].shape[0], w["inv_depth"].shape[0]
    D = 15 * N + 7
    w2 = dict(w)
    for k in ("pose", "speedbias", "ex_pose", "inv_depth"):
        w2[k] = np.array(w[k], dtype=np.float64, copy=True)
    if col >= D:
        w2["inv_depth"][col - D] += eps
    elif col == 15 * N + 6:
        w2["td"] = w["td"] + eps
    elif col >= 15 * N:
        d = np.zeros(6); d[col - 15 * N] = eps
        w2["ex_pose"] = helpers.pose_plus(w["ex_pose"], d)
    else:
        k, c = divmod(col, 15)
        if c < 6:
            if global_quat and c >= 3:
                w2["pose"][k, c] += eps          # raw quaternion component (x, y or z)
            else:
                d = np.zeros(6); d[c] = eps
                w2["pose"][k] = helpers.pose_plus(w["pose"][k], d)
        else:
            w2["speedbias"][k, c - 6] += eps
    return w2


@pytest.mark.parametrize("use_td", [1, 0])
def test_jacobians_match_finite_differences(use_td):
    w = small_window()
    cfg = cabi.default_config()
    cfg.estimate_td = use_td
    r0, J0, _ = ol.evaluate_window(cfg, w, apply_loss=False)
    Jd, fams = dense_jacobian(w, J0)
    N, M = w["pose"].shape[0], w["inv_depth"].shape[0]
    T = 15 * N + 7 + M
    eps = 1e-6
    autodiff_rows = np.isin(fams, ["icp", "lps"])
    for col in range(T):
        rp = ol.evaluate_window(cfg, perturbed(w, col, eps), apply_loss=False)[0]
        rm = ol.evaluate_window(cfg, perturbed(w, col, -eps), apply_loss=False)[0]
        fd = (rp - rm) / (2 * eps)
        scale = np.maximum(np.abs(Jd).max(axis=1), 1.0)
        err = np.abs(fd - Jd[:, col]) / scale
        is_rot = col < 15 * N and (col % 15) in (3, 4, 5)
        rows = ~autodiff_rows if is_rot else np.ones_like(autodiff_rows)
        assert err[rows].max() < 2e-6, (col, err[rows].max())
        if is_rot and autodiff_rows.any():
            # ICP / LPS are ceres autodiff functors under PoseLocalParameterization, whose ComputeJacobian is [I6; 0]
            # (pose_local_parameterization.cpp:20-27): the kept columns are d r / d (qx, qy, qz) of the raw quaternion.
            rp = ol.evaluate_window(cfg, perturbed(w, col, eps, True), apply_loss=False)[0]
            rm = ol.evaluate_window(cfg, perturbed(w, col, -eps, True), apply_loss=False)[0]
            fd = (rp - rm) / (2 * eps)
            err = np.abs(fd - Jd[:, col]) / scale
            assert err[autodiff_rows].max() < 2e-6, (col, err[autodiff_rows].max())


def test_noise_free_truth_is_a_fixed_point():
    w = synth.make_window(config_id=9, window_idx=2, N=6, M=20, n_lidar=60, n_icp=2, n_lps=2, noise_free=True, perturb=False)
    cfg = cabi.default_config()
    r, J, cost = ol.evaluate_window(cfg, w, apply_loss=False)
    # projection/LiDAR/ICP/LPS residuals vanish exactly at the truth; the IMU residual keeps the mid-point
    # discretisation error (whitened by sqrt_info), so only a loose bound applies there.
    n_imu = 15 * len(w["imu"])
    assert np.abs(r[n_imu:]).max() < 1e-6
    assert cost < 1e3


def test_robust_corrector_matches_closed_form():
    # rho'' <= 0 for Cauchy/Huber => J <- sqrt(rho') J, r <- sqrt(rho') r (marginalization_factor.cpp:49-53,62-67)
    w = small_window(n_icp=0, n_lps=0)
    cfg = cabi.default_config()
    r0, J0, _ = ol.evaluate_window(cfg, w, apply_loss=False)
    r1, J1, cost = ol.evaluate_window(cfg, w, apply_loss=True)
    ro = jo = 0
    total = 0.0
    for fam, k, nr, blocks in helpers.factor_layout(w):
        width = sum(s for _, s in blocks)
        s = float(r0[ro:ro + nr] @ r0[ro:ro + nr])
        if fam == "proj":
            rho1, rho = 1.0 / (1.0 + s), np.log1p(s)
        elif fam in ("plane", "edge"):
            a = cfg.huber_lidar
            rho1, rho = (1.0, s) if s <= a * a else (a / np.sqrt(s), 2 * a * np.sqrt(s) - a * a)
        else:
            rho1, rho = 1.0, s
        np.testing.assert_allclose(r1[ro:ro + nr], np.sqrt(rho1) * r0[ro:ro + nr], rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(J1[jo:jo + nr * width], np.sqrt(rho1) * J0[jo:jo + nr * width], rtol=1e-13, atol=1e-15)
        total += 0.5 * rho
        ro += nr; jo += nr * width
    assert abs(total - cost) < 1e-9 * max(1.0, abs(cost))


def test_linearize_matches_numpy_assembly():
    w = small_window()
    cfg = cabi.default_config()
    r, J, cost = ol.evaluate_window(cfg, w, apply_loss=True)
    H, g = helpers.dense_normal(cfg, w, r, J)
    H, g = helpers.apply_fixed(cfg, w, H, g)
    D = 15 * w["pose"].shape[0] + 7
    S_np, g_np = helpers.schur_reduce(H, g, D)
    S, gr, cost2 = ol.linearize_window(cfg, w)
    assert abs(cost - cost2) < 1e-9 * abs(cost)
    scale = np.abs(S_np).max()
    assert np.abs(S - S_np).max() < 1e-10 * scale
    assert np.abs(gr - g_np).max() < 1e-10 * np.abs(g_np).max()


def test_gauss_newton_and_lm_converge_to_the_same_minimum():
    w = synth.make_window(config_id=9, window_idx=3, N=6, M=30, n_lidar=200)
    cfg = cabi.default_config()
    gn = ol.solve_window(cfg, w, cabi.default_solve_opts(cabi.VILS_MODE_GN, 40, 1e-8))
    # GN without step control converges linearly along the weakly observable td / extrinsic valley
    lm = ol.solve_window(cfg, w, cabi.default_solve_opts(cabi.VILS_MODE_LM, 50, 0.0))
    assert gn["status"] == 0 and lm["status"] == 0
    assert gn["cost_final"] < 1e-3 * gn["cost_initial"]
    assert abs(gn["cost_final"] - lm["cost_final"]) < 1e-4 * lm["cost_final"]
    assert helpers.rel_state_delta(gn, lm) < 2e-2
    # the minimum is near the truth (LiDAR planes fix the gauge)
    assert np.abs(gn["pose"][:, :3] - w["truth"]["pose"][:, :3]).max() < 0.05


def test_solution_is_a_stationary_point_of_scipy_least_squares():
    """Independent check of the converged solution: scipy's trust-region least squares on residuals rescaled so that
    |r~|^2 = rho(|r|^2) (the exact robust objective, finite-difference Jacobian) cannot improve on the oracle's LM."""
    scipy_opt = pytest.importorskip("scipy.optimize")
    w = synth.make_window(config_id=9, window_idx=4, N=4, M=10, n_lidar=80)
    cfg = cabi.default_config()
    o = cabi.default_solve_opts(cabi.VILS_MODE_LM, 100, 0.0)
    o.function_tolerance = 1e-14; o.parameter_tolerance = 1e-14
    sol = ol.solve_window(cfg, w, o)
    ws = dict(w); ws.update({k: sol[k] for k in ("pose", "speedbias", "ex_pose", "inv_depth", "td")})
    N, M = 4, 10
    T = 15 * N + 7 + M
    free = np.array([c for c in range(T) if not (c >= 15 * N + 7 and w["depth_fixed"][c - 15 * N - 7])])
    layout = list(helpers.factor_layout(w))

    def fun(x):
        w2 = ws
        for c, v in zip(free, x):
            if v != 0.0:
                w2 = perturbed(w2, int(c), float(v))
        r = ol.evaluate_window(cfg, w2, apply_loss=False)[0].copy()
        ro = 0
        for fam, k, nr, blocks in layout:
            s = float(r[ro:ro + nr] @ r[ro:ro + nr])
            if s > 0 and fam == "proj":
                r[ro:ro + nr] *= np.sqrt(np.log1p(s) / s)
            elif s > 0 and fam in ("plane", "edge") and s > cfg.huber_lidar ** 2:
                r[ro:ro + nr] *= np.sqrt((2 * cfg.huber_lidar * np.sqrt(s) - cfg.huber_lidar ** 2) / s)
            ro += nr
        return r

    f0 = fun(np.zeros(len(free)))
    assert abs(0.5 * f0 @ f0 - sol["cost_final"]) < 1e-9 * sol["cost_final"]
    res = scipy_opt.least_squares(fun, np.zeros(len(free)), method="trf", xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=6,
                                  diff_step=1e-7)
    assert res.cost <= sol["cost_final"] * (1 + 1e-12)
    assert sol["cost_final"] - res.cost < 1e-7 * sol["cost_final"]
