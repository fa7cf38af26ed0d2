"""Test helpers: independent numpy assembly of the normal equations from materialised residuals/Jacobians."""
import numpy as np

from mvil_fusion_b200 import cabi


def factor_layout(w):
    """Yield (family, index, nr, [(tangent offset, size), ...]) in the order of vils_ba_evaluate's output."""
    N = w["pose"].shape[0]
    D = 15 * N + 7
    for k in range(0 if w.get("imu") is None else len(w["imu"])):
        i = int(w["imu_kf"][k]); j = i + 1
        yield "imu", k, 15, [(15 * i, 6), (15 * i + 6, 9), (15 * j, 6), (15 * j + 6, 9)]
    for k in range(0 if w.get("kf_i") is None else len(w["kf_i"])):
        i, j, f = int(w["kf_i"][k]), int(w["kf_j"][k]), int(w["feat"][k])
        yield "proj", k, 2, [(15 * i, 6), (15 * j, 6), (15 * N, 6), (D + f, 1), (15 * N + 6, 1)]
    for k in range(0 if w.get("plane_kf") is None else len(w["plane_kf"])):
        yield "plane", k, 1, [(15 * int(w["plane_kf"][k]), 6)]
    for k in range(0 if w.get("edge_kf") is None else len(w["edge_kf"])):
        yield "edge", k, 3, [(15 * int(w["edge_kf"][k]), 6)]
    for k, c in enumerate(w.get("icp") or []):
        yield "icp", k, 3, [(15 * int(x), 6) for x in c["kf"]]
    for k, c in enumerate(w.get("lps") or []):
        yield "lps", k, 3, [(15 * int(x), 6) for x in c["kf"]]


def prior_columns(w):
    N = w["pose"].shape[0]
    cols = []
    for b in w["prior_blk"]:
        t, i = cabi.blk_type(int(b)), cabi.blk_index(int(b))
        off = {0: 15 * i, 1: 15 * i + 6, 2: 15 * N, 3: 15 * N + 6}[t]
        cols += list(range(off, off + cabi.blk_local_size(t)))
    return np.array(cols)


def dense_normal(cfg, w, r, J):
    """H (T x T), g (T) over [camera D | landmarks M] from the flat outputs of evaluate (loss already applied)."""
    N, M = w["pose"].shape[0], w["inv_depth"].shape[0]
    D = 15 * N + 7
    T = D + M
    H = np.zeros((T, T)); g = np.zeros(T)
    ro = jo = 0
    for fam, k, nr, blocks in factor_layout(w):
        width = sum(s for _, s in blocks)
        Jf = J[jo:jo + nr * width].reshape(nr, width); rf = r[ro:ro + nr]
        cols = np.concatenate([np.arange(o, o + s) for o, s in blocks])
        H[np.ix_(cols, cols)] += Jf.T @ Jf
        g[cols] += Jf.T @ rf
        ro += nr; jo += nr * width
    n = int(w.get("prior_n", 0))
    if n:
        Jp = np.asarray(w["prior_J"]).reshape(n, n).T   # stored column-major
        cols = prior_columns(w)
        rp = r[ro:ro + n]
        H[np.ix_(cols, cols)] += Jp.T @ Jp
        g[cols] += Jp.T @ rp
    return H, g


def apply_fixed(cfg, w, H, g):
    N, M = w["pose"].shape[0], w["inv_depth"].shape[0]
    D = 15 * N + 7
    fixed = np.zeros(D + M, bool)
    if not cfg.estimate_extrinsic:
        fixed[15 * N:15 * N + 6] = True
    if not cfg.estimate_td:
        fixed[15 * N + 6] = True
    if w.get("kf_fixed") is not None:
        for k in np.nonzero(w["kf_fixed"])[0]:
            fixed[15 * k:15 * k + 15] = True
    df = w.get("depth_fixed")
    for f in range(M):
        if (df is not None and df[f]) or H[D + f, D + f] == 0.0:
            fixed[D + f] = True
    H = H.copy(); g = g.copy()
    H[fixed, :] = 0; H[:, fixed] = 0
    H[fixed, fixed] = 1.0
    g[fixed] = 0
    return H, g


def schur_reduce(H, g, D):
    C = np.diag(H)[D:]
    E = H[:D, D:]
    S = H[:D, :D] - (E / C) @ E.T
    gr = g[:D] - (E / C) @ g[D:]
    return S, gr


def pose_plus(pose7, d6):
    """PoseLocalParameterization::Plus (pose_local_parameterization.cpp:3-19)."""
    from mvil_fusion_b200.synth import quat_mul
    out = np.array(pose7, dtype=np.float64)
    out[:3] += d6[:3]
    dq = np.array([d6[3] / 2, d6[4] / 2, d6[5] / 2, 1.0])
    q = quat_mul(out[3:], dq)
    out[3:] = q / np.linalg.norm(q)
    return out


def rel_state_delta(a, b):
    """max over blocks of |a-b| / max(|b|, 1) — the 'state delta' of the 1e-5 parity bar."""
    worst = 0.0
    for k in ["pose", "speedbias", "ex_pose", "inv_depth"]:
        x, y = np.asarray(a[k]), np.asarray(b[k])
        if k in ("pose", "ex_pose"):   # quaternion sign
            x = x.reshape(-1, 7).copy(); y = y.reshape(-1, 7)
            sgn = np.sign(np.sum(x[:, 3:] * y[:, 3:], axis=1, keepdims=True)); sgn[sgn == 0] = 1
            x[:, 3:] *= sgn
        worst = max(worst, float(np.max(np.abs(x - y) / np.maximum(np.abs(y), 1.0))))
    worst = max(worst, abs(a["td"] - b["td"]) / max(abs(b["td"]), 1.0))
    return worst
