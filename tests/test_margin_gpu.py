"""GPU marginalization (vils_ba_marginalize) vs the oracle's marginalize() at the same state, and the config-4 chain:
(N+1)-frame window -> solve -> MARGIN_OLD prior -> slide -> N-frame window carrying that real prior (+ ICP/LPS) -> solve.
J_lin itself is only defined up to eigenvector signs/ordering, so parity is judged on J^T J and J^T r
(marginalization_factor.cpp:313-314 uses the same identities as its self-check): 1e-8 relative."""
import numpy as np
import pytest

import helpers
import oracle_lib as ol
from mvil_fusion_b200 import cabi, synth

pytestmark = pytest.mark.gpu


def compare_prior(pg, po):
    assert pg["n"] == po["n"] and pg["m"] == po["m"]
    assert np.array_equal(pg["blk"], po["blk"])
    np.testing.assert_allclose(pg["x0"], po["x0"], rtol=0, atol=0)
    n = pg["n"]
    Jg = pg["J"].reshape(n, n).T; Jo = po["J"].reshape(n, n).T      # stored column-major
    Ag, Ao = Jg.T @ Jg, Jo.T @ Jo
    bg, bo = Jg.T @ pg["r"], Jo.T @ po["r"]
    assert np.abs(Ag - Ao).max() <= 1e-8 * np.abs(Ao).max()
    assert np.abs(bg - bo).max() <= 1e-8 * max(np.abs(bo).max(), 1e-12)


def solved_window(w, s):
    w2 = dict(w)
    w2.update(pose=s["pose"], speedbias=s["speedbias"], ex_pose=s["ex_pose"], inv_depth=s["inv_depth"], td=s["td"])
    return w2


@pytest.mark.parametrize("N,M,nl,icp,lps", [(6, 30, 200, 0, 0), (8, 60, 400, 3, 3)])
def test_margin_old_matches_oracle(N, M, nl, icp, lps):
    from mvil_fusion_b200 import lib
    cfg = cabi.default_config(max_kf=N, max_feat=M, max_proj=M * N, max_lidar=nl)
    w = synth.make_window(config_id=9, window_idx=40 + N, N=N, M=M, n_lidar=nl, n_icp=icp, n_lps=lps)
    if icp:   # make sure one ICP and one LPS constraint start at frame 0 so that they are marginalised
        w["icp"][0] = dict(w["icp"][0]); w["lps"][0] = dict(w["lps"][0])
        tk = w["truth"]["t_kf"]
        w["icp"][0].update(kf=(0, 1, 2, 3), t=(tk[0], tk[1], tk[2], tk[3], tk[0] + 0.03, tk[2] + 0.04))
        w["lps"][0].update(kf=(0, 1), t=(tk[0], tk[1], tk[0] + 0.05))
    ba = lib.BA(cfg, 1)
    ba.set_window(0, w)
    ba.solve(1, cabi.default_solve_opts(cabi.VILS_MODE_GN, 5, 1e-8))
    s = ba.get_state(0)
    assert s["status"] == 0
    pg = ba.marginalize(0, cabi.VILS_MARGIN_OLD)
    po = ol.marginalize_window(cfg, solved_window(w, s), cabi.VILS_MARGIN_OLD)
    compare_prior(pg, po)
    assert pg["n"] == 6 * (N - 1) + 9 + 6 + 1      # poses 1..N-1, sb1, ex, td
    ba.close()


def test_margin_second_new_and_config4_chain():
    from mvil_fusion_b200 import lib
    N = 9
    cfg = cabi.default_config(max_kf=N, max_feat=80, max_proj=80 * N, max_lidar=600)
    big = synth.make_window(config_id=4, window_idx=0, N=N, M=80, n_lidar=600, n_icp=3, n_lps=3)
    ba = lib.BA(cfg, 1)
    ba.set_window(0, big)
    opts = cabi.default_solve_opts(cabi.VILS_MODE_GN, 5, 1e-8)
    ba.solve(1, opts)
    s = ba.get_state(0)
    prior = ba.marginalize(0, cabi.VILS_MARGIN_OLD)
    compare_prior(prior, ol.marginalize_window(cfg, solved_window(big, s), cabi.VILS_MARGIN_OLD))
    # slide: drop frame 0, carry the prior (ids are already re-addressed), solve the (N-1)-frame window
    w = synth.attach_prior(synth.slide_old(solved_window(big, s)), prior)
    w.pop("truth", None)
    ba.set_window(0, w)
    ba.solve(1, opts)
    g = ba.get_state(0)
    o = ol.solve_window(cfg, w, opts)
    assert g["status"] == 0 and o["status"] == 0
    assert abs(g["cost_final"] - o["cost_final"]) <= 1e-7 * max(o["cost_final"], 1e-9)
    assert helpers.rel_state_delta(g, o) <= 1e-5
    # MARGIN_SECOND_NEW on the window that now carries a dense prior containing pose[N-2]
    p2 = ba.marginalize(0, cabi.VILS_MARGIN_SECOND_NEW)
    o2 = ol.marginalize_window(cfg, solved_window(w, g), cabi.VILS_MARGIN_SECOND_NEW)
    compare_prior(p2, o2)
    assert p2["m"] == 6
    # a window whose prior does not contain pose[N-2] skips it (estimator.cpp:1620-1621)
    w0 = synth.make_window(config_id=9, window_idx=50, N=6, M=30, n_lidar=100)
    ba.set_window(0, w0); ba.solve(1, opts)
    assert ba.marginalize(0, cabi.VILS_MARGIN_SECOND_NEW)["n"] == 0
    ba.close()
