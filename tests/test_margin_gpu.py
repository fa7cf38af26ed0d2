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


def reanchored(lib, w, s):
    """double2vector (estimator.cpp:962-1011) applied to a solved state: what the reference's vector2double (:1487) re-packs
    before it builds MarginalizationInfo."""
    pose, sb = lib.double2vector(w["pose"][0], s["pose"], s["speedbias"])
    d = dict(s); d["pose"] = pose; d["speedbias"] = sb
    return d


def test_prior_is_linearised_at_the_reanchored_state():
    """solve -> double2vector -> put_state -> marginalize -> slide -> next solve: the prior's x0 snapshots are the re-anchored (carried)
    state, so at the carried state dx = 0 and the prior residual is exactly r_lin (the raw-gauge prior fought the re-anchoring)."""
    from mvil_fusion_b200 import lib
    N = 8
    cfg = cabi.default_config(max_kf=N, max_feat=60, max_proj=60 * N, max_lidar=400)
    w = synth.make_window(config_id=9, window_idx=61, N=N, M=60, n_lidar=400, n_icp=2, n_lps=2)
    ba = lib.BA(cfg, 1)
    ba.set_window(0, w)
    opts = cabi.default_solve_opts(cabi.VILS_MODE_GN, 5, 1e-8)
    ba.solve(1, opts)
    s = ba.get_state(0)
    sa = reanchored(lib, w, s)
    assert np.abs(sa["pose"] - s["pose"]).max() > 1e-6        # the gauge did move: the test is not vacuous
    ba.put_state(0, sa["pose"], sa["speedbias"], sa["ex_pose"], sa["inv_depth"], sa["td"])
    pg = ba.marginalize(0, cabi.VILS_MARGIN_OLD)
    po = ol.marginalize_window(cfg, solved_window(w, sa), cabi.VILS_MARGIN_OLD)
    compare_prior(pg, po)
    # x0 of the kept pose blocks == the carried (re-anchored) poses 1..N-1, in block order
    nxt = synth.attach_prior(synth.slide_old(solved_window(w, sa)), pg)
    nxt.pop("truth", None)
    off = 0
    for b in pg["blk"]:
        t, i = cabi.blk_type(int(b)), cabi.blk_index(int(b))
        gs = cabi.blk_global_size(t)
        cur = {0: lambda: nxt["pose"][i], 1: lambda: nxt["speedbias"][i], 2: lambda: nxt["ex_pose"], 3: lambda: np.array([nxt["td"]])}[t]()
        np.testing.assert_array_equal(pg["x0"][off:off + gs], cur)
        off += gs
    # prior residual at the carried state is r_lin itself (dx = 0)
    ba.set_window(0, nxt); ba.upload(1)
    r, _ = ba.evaluate(0, True)
    np.testing.assert_allclose(r[-pg["n"]:], pg["r"], rtol=0, atol=1e-12)
    # and the next solve matches the oracle on the same window
    ba.solve(1, opts)
    g = ba.get_state(0); o = ol.solve_window(cfg, nxt, opts)
    assert g["status"] == 0 and o["status"] == 0
    assert helpers.rel_state_delta(g, o) <= 1e-5
    ba.close()


def test_config4_full_size_with_real_marginalization_prior():
    """BASELINE configs[3] exactly as SURVEY.md 8d states it: a 21-frame window is solved and marginalised (MARGIN_OLD, at the re-anchored
    state), slid to N = 20 / M = 300 / 3333 projection + 3750 plane + 1250 edge factors / 3 ICP + 3 LPS carrying the REAL dense prior
    (n = 136 = 6 x 20 poses + sb + ex + td: every new-window pose is seen by a landmark anchored at the dropped frame).
    Bars: prior J^T J / J^T r 1e-8 (GPU marginalization vs oracle), reduced system 1e-9, GN-5 state <= 1e-5 vs the oracle."""
    from mvil_fusion_b200 import lib
    cfg = cabi.default_config(max_kf=21, max_feat=320, max_proj=4000, max_lidar=5000)
    big = synth.make_config4_big(0)
    opts = cabi.default_solve_opts(cabi.VILS_MODE_GN, 5, 1e-8)
    ba = lib.BA(cfg, 1)
    ba.set_window(0, big)
    ba.solve(1, opts)
    s = ba.get_state(0)
    assert s["status"] == 0
    sa = reanchored(lib, big, s)
    ba.put_state(0, sa["pose"], sa["speedbias"], sa["ex_pose"], sa["inv_depth"], sa["td"])
    prior = ba.marginalize(0, cabi.VILS_MARGIN_OLD)
    compare_prior(prior, ol.marginalize_window(cfg, solved_window(big, sa), cabi.VILS_MARGIN_OLD))
    assert prior["n"] == 6 * 20 + 9 + 6 + 1 and prior["m"] == 15 + 18
    w = synth.attach_prior(synth.slide_old(solved_window(big, sa)), prior)
    w.pop("truth", None)
    assert w["pose"].shape[0] == 20 and w["inv_depth"].shape[0] == 300 and len(w["kf_i"]) == 3333
    assert len(w["plane_kf"]) == 3750 and len(w["edge_kf"]) == 1250 and len(w["icp"]) == 3 and len(w["lps"]) == 3
    ba.set_window(0, w); ba.upload(1)
    S, g, cost = ba.linearize(0)
    So, go, co = ol.linearize_window(cfg, w)
    assert abs(cost - co) <= 1e-10 * abs(co)
    assert np.abs(S - So).max() <= 1e-9 * np.abs(So).max()
    assert np.abs(g - go).max() <= 1e-9 * np.abs(go).max()
    ba.solve(1, opts)
    gs = ba.get_state(0)
    o = ol.solve_window(cfg, w, opts)
    assert gs["status"] == 0 and o["status"] == 0
    assert abs(gs["cost_final"] - o["cost_final"]) <= 1e-7 * o["cost_final"]
    ga = reanchored(lib, w, gs); oa = dict(o)
    oa["pose"], oa["speedbias"] = ol.double2vector(w["pose"][0], o["pose"], o["speedbias"])
    assert helpers.rel_state_delta(ga, oa) <= 1e-5
    ba.close()
