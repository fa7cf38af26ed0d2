"""Seeded synthetic sliding windows: trajectory -> IMU samples -> mid-point pre-integration ->
landmarks / observations / velocities -> LiDAR plane / edge factor constants -> perturbed initial state.

Pure numpy, no GPU, no oracle: this is the input generator shared by tests/ and bench.py.  The constants come
from the reference's config/mynteye_leishen_indoor.yaml (cited in cabi.py); the pre-integration follows
vils_estimator/src/factor/integration_base.h:54-158 so that covariance / bias Jacobians have the reference's
structure (tests/test_preint.py checks it against the oracle and against the CUDA kernel).
"""
import numpy as np

from . import cabi

KF_DT = 0.1       # freq: 10 (yaml:69)
IMU_DT = 0.005    # 200 Hz (README.md:18)
SAMPLES = 20


# ------------------------------------------------------------------------------------------------
# small rotation helpers (quaternions stored x y z w like para_Pose, estimator.cpp:923-927)
# ------------------------------------------------------------------------------------------------
def skew(v):
    v = np.asarray(v)
    z = np.zeros(v.shape[:-1])
    return np.stack([np.stack([z, -v[..., 2], v[..., 1]], -1),
                     np.stack([v[..., 2], z, -v[..., 0]], -1),
                     np.stack([-v[..., 1], v[..., 0], z], -1)], -2)


def quat_mul(a, b):
    ax, ay, az, aw = np.moveaxis(a, -1, 0)
    bx, by, bz, bw = np.moveaxis(b, -1, 0)
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz], -1)


def quat_conj(q):
    return q * np.array([-1.0, -1.0, -1.0, 1.0])


def quat_to_R(q):
    x, y, z, w = np.moveaxis(q, -1, 0)
    return np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
                     np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
                     np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], -2)


def R_to_quat(R):
    """Single 3x3 -> (x y z w), w >= 0."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[3] = (R[k, j] - R[j, k]) / s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
    if q[3] < 0:
        q = -q
    return q / np.linalg.norm(q)


def quat_rot(q, v):
    return np.einsum("...ij,...j->...i", quat_to_R(q), v)


def small_quat(theta):
    """Unit quaternion of a rotation vector."""
    theta = np.asarray(theta, dtype=np.float64)
    a = np.linalg.norm(theta, axis=-1, keepdims=True)
    h = 0.5 * a
    k = np.where(a > 1e-12, np.sin(h) / np.maximum(a, 1e-300), 0.5)
    return np.concatenate([theta * k, np.cos(h)], -1)


def quat_slerp(a, b, t):
    """Eigen QuaternionBase::slerp semantics (shortest arc)."""
    d = float(np.dot(a, b))
    ad = abs(d)
    if ad >= 1.0 - np.finfo(np.float64).eps:
        s0, s1 = 1.0 - t, t
    else:
        th = np.arccos(ad)
        s0, s1 = np.sin((1 - t) * th) / np.sin(th), np.sin(t * th) / np.sin(th)
    if d < 0:
        s1 = -s1
    return s0 * a + s1 * b


# ------------------------------------------------------------------------------------------------
# trajectory: p(t) = [2 sin .5t, 2 sin(.4t+1), .3 sin .7t]; yaw .3t, pitch/roll .05 sin(1.1t [+.5])
# ------------------------------------------------------------------------------------------------
def traj(t):
    t = np.asarray(t, dtype=np.float64)
    p = np.stack([2 * np.sin(0.5 * t), 2 * np.sin(0.4 * t + 1), 0.3 * np.sin(0.7 * t)], -1)
    v = np.stack([np.cos(0.5 * t), 0.8 * np.cos(0.4 * t + 1), 0.21 * np.cos(0.7 * t)], -1)
    a = np.stack([-0.5 * np.sin(0.5 * t), -0.32 * np.sin(0.4 * t + 1), -0.147 * np.sin(0.7 * t)], -1)
    yaw, dyaw = 0.3 * t, 0.3 * np.ones_like(t)
    pit, dpit = 0.05 * np.sin(1.1 * t), 0.055 * np.cos(1.1 * t)
    rol, drol = 0.05 * np.sin(1.1 * t + 0.5), 0.055 * np.cos(1.1 * t + 0.5)
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pit), np.sin(pit), np.cos(rol), np.sin(rol)
    R = np.stack([np.stack([cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr], -1),
                  np.stack([sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr], -1),
                  np.stack([-sp, cp * sr, cp * cr], -1)], -2)
    # body rates of a ZYX Euler sequence
    w = np.stack([drol - dyaw * sp, dpit * cr + dyaw * sr * cp, -dpit * sr + dyaw * cr * cp], -1)
    return p, v, a, R, w


# ------------------------------------------------------------------------------------------------
# mid-point pre-integration, batched over K intervals (integration_base.h:54-158)
# ------------------------------------------------------------------------------------------------
def preintegrate(dt, acc, gyr, acc0, gyr0, ba, bg, noise=(cabi.ACC_N, cabi.GYR_N, cabi.ACC_W, cabi.GYR_W)):
    """acc, gyr: (K, S, 3) samples pushed after (acc0, gyr0): (K, 3).  Returns (K, 467) vils_preint rows."""
    K, S, _ = acc.shape
    dp = np.zeros((K, 3)); dv = np.zeros((K, 3)); dq = np.tile(np.array([0, 0, 0, 1.0]), (K, 1))
    J = np.tile(np.eye(15), (K, 1, 1)); P = np.zeros((K, 15, 15))
    an, gn, aw, gw = noise
    Q = np.diag(np.repeat([an * an, gn * gn, an * an, gn * gn, aw * aw, gw * gw], 3))
    I3 = np.eye(3)
    a0, g0 = acc0.copy(), gyr0.copy()
    for s in range(S):
        a1, g1 = acc[:, s], gyr[:, s]
        h = dt
        un_gyr = 0.5 * (g0 + g1) - bg
        dq_step = np.concatenate([un_gyr * h / 2, np.ones((K, 1))], -1)       # [1, w dt/2] (not normalised) :66
        rq = quat_mul(dq, dq_step)
        Rd, Rr = quat_to_R(dq), quat_to_R(rq)                                   # toRotationMatrix of unnormalised rq
        un_acc_0 = np.einsum("kij,kj->ki", Rd, a0 - ba)
        un_acc_1 = np.einsum("kij,kj->ki", Rr, a1 - ba)
        un_acc = 0.5 * (un_acc_0 + un_acc_1)
        rp = dp + dv * h + 0.5 * un_acc * h * h
        rv = dv + un_acc * h
        Rwx, Ra0, Ra1 = skew(un_gyr), skew(a0 - ba), skew(a1 - ba)
        F = np.zeros((K, 15, 15)); V = np.zeros((K, 15, 18))
        IRw = I3 - Rwx * h
        F[:, 0:3, 0:3] = I3
        F[:, 0:3, 3:6] = -0.25 * (Rd @ Ra0) * h * h + -0.25 * (Rr @ Ra1 @ IRw) * h * h
        F[:, 0:3, 6:9] = I3 * h
        F[:, 0:3, 9:12] = -0.25 * (Rd + Rr) * h * h
        F[:, 0:3, 12:15] = -0.25 * (Rr @ Ra1) * h * h * -h
        F[:, 3:6, 3:6] = IRw
        F[:, 3:6, 12:15] = -I3 * h
        F[:, 6:9, 3:6] = -0.5 * (Rd @ Ra0) * h + -0.5 * (Rr @ Ra1 @ IRw) * h
        F[:, 6:9, 6:9] = I3
        F[:, 6:9, 9:12] = -0.5 * (Rd + Rr) * h
        F[:, 6:9, 12:15] = -0.5 * (Rr @ Ra1) * h * -h
        F[:, 9:12, 9:12] = I3
        F[:, 12:15, 12:15] = I3
        V[:, 0:3, 0:3] = 0.25 * Rd * h * h
        V[:, 0:3, 3:6] = 0.25 * -(Rr @ Ra1) * h * h * 0.5 * h
        V[:, 0:3, 6:9] = 0.25 * Rr * h * h
        V[:, 0:3, 9:12] = V[:, 0:3, 3:6]
        V[:, 3:6, 3:6] = 0.5 * I3 * h
        V[:, 3:6, 9:12] = 0.5 * I3 * h
        V[:, 6:9, 0:3] = 0.5 * Rd * h
        V[:, 6:9, 3:6] = 0.5 * -(Rr @ Ra1) * h * 0.5 * h
        V[:, 6:9, 6:9] = 0.5 * Rr * h
        V[:, 6:9, 9:12] = V[:, 6:9, 3:6]
        V[:, 9:12, 12:15] = I3 * h
        V[:, 12:15, 15:18] = I3 * h
        J = F @ J
        P = F @ P @ np.transpose(F, (0, 2, 1)) + V @ Q @ np.transpose(V, (0, 2, 1))
        dp, dv = rp, rv
        dq = rq / np.linalg.norm(rq, axis=-1, keepdims=True)
        a0, g0 = a1, g1
    out = np.zeros((K, cabi.PREINT_DOUBLES))
    out[:, 0:3] = dp
    out[:, 3:7] = dq
    out[:, 7:10] = dv
    out[:, 10:13] = ba
    out[:, 13:16] = bg
    out[:, 16] = S * dt
    out[:, 17:242] = np.transpose(J, (0, 2, 1)).reshape(K, 225)    # column-major
    out[:, 242:467] = np.transpose(P, (0, 2, 1)).reshape(K, 225)
    return out


# ------------------------------------------------------------------------------------------------
# window generator
# ------------------------------------------------------------------------------------------------
ROOM = dict(x=(-8.0, 8.0), y=(-6.0, 6.0), z=(-1.5, 3.0))


def _lidar_factors(rng, n_lidar, N, Rk, Pk, RLB, TLB, nz=1.0, kf0=0):
    n_edge = n_lidar // 4
    n_plane = n_lidar - n_edge
    axes = "xyz"
    lo = np.array([ROOM[a][0] for a in axes]); hi = np.array([ROOM[a][1] for a in axes])

    def to_lidar(pw, kf):
        pb = np.einsum("nji,nj->ni", Rk[kf], pw - Pk[kf])
        return pb @ RLB.T + TLB + nz * rng.normal(0, 0.02, pw.shape)

    # planes: axis ax at lo/hi; unit normal +e_ax; n.p + d = 0
    pax = rng.integers(0, 3, n_plane); pside = rng.integers(0, 2, n_plane)
    pw = rng.uniform(lo, hi, (n_plane, 3))
    pval = np.where(pside == 0, lo[pax], hi[pax])
    pw[np.arange(n_plane), pax] = pval
    pn = np.zeros((n_plane, 3)); pn[np.arange(n_plane), pax] = 1.0
    pd = -pval
    pkf = (kf0 + np.arange(n_plane) % (N - kf0)).astype(np.int32)
    plane_p = to_lidar(pw, pkf)
    # edges: intersection of two walls/floor/ceiling -> line along the third axis
    eax = rng.integers(0, 3, n_edge)                 # direction axis
    c = rng.uniform(lo, hi, (n_edge, 3))
    for k in range(3):
        sel = eax != k
        side = rng.integers(0, 2, n_edge)
        c[sel, k] = np.where(side[sel] == 0, lo[k], hi[k])
    d = np.zeros((n_edge, 3)); d[np.arange(n_edge), eax] = 1.0
    ea, eb = c + 0.1 * d, c - 0.1 * d                 # localMapping.cpp:661-662
    pw_e = c + d * rng.uniform(-1.0, 1.0, (n_edge, 1))
    ekf = (kf0 + (np.arange(n_edge) + n_plane) % (N - kf0)).astype(np.int32)
    edge_p = to_lidar(pw_e, ekf)
    return dict(plane_p=plane_p, plane_n=pn, plane_d=pd, plane_kf=pkf, edge_p=edge_p, edge_a=ea, edge_b=eb, edge_kf=ekf)


def make_window(config_id=2, window_idx=0, N=10, M=150, n_lidar=2000, n_icp=0, n_lps=0, noise_free=False,
                perturb=True, seed=None, ex_prior=True, start=None, lidar_kf0=0, icp_anchors=None, lps_anchors=None):
    """One synthetic window as a dict of numpy arrays keyed like vils_window (include/vils_cabi.h).

    Extra keys: 'truth' (dict of true state) and 'cfg' constants are NOT part of the C struct."""
    if seed is None:
        seed = 20240000 + 1000 * config_id + window_idx
    rng = np.random.default_rng(seed)
    nz = 0.0 if noise_free else 1.0
    RLB = cabi._orthonormalize(cabi.GT_RLI); TLB = cabi.GT_TLI
    ric, tic = cabi._orthonormalize(cabi.RIC), cabi.TIC
    G = np.array([0, 0, cabi.G_NORM])
    t0 = rng.uniform(0.0, 20.0)
    # --- IMU samples
    ns = (N - 1) * SAMPLES + 1
    ts = t0 + IMU_DT * np.arange(ns)
    p, v, a, R, w = traj(ts)
    ba_true = rng.normal(0, 0.02, 3); bg_true = rng.normal(0, 0.002, 3)
    acc_m = np.einsum("nji,nj->ni", R, a + G) + ba_true + nz * rng.normal(0, cabi.ACC_N, (ns, 3))
    gyr_m = w + bg_true + nz * rng.normal(0, cabi.GYR_N, (ns, 3))
    kf = np.arange(N) * SAMPLES
    Pk, Vk, Rk = p[kf], v[kf], R[kf]
    Qk = np.stack([R_to_quat(Rk[i]) for i in range(N)])
    # --- initial (perturbed) state
    s = 1.0 if perturb else 0.0
    pose = np.zeros((N, 7)); sb = np.zeros((N, 9))
    pose[:, :3] = Pk + s * rng.normal(0, 0.05, (N, 3))
    q0 = quat_mul(Qk, small_quat(s * rng.normal(0, 0.01, (N, 3))))
    pose[:, 3:] = q0 / np.linalg.norm(q0, axis=-1, keepdims=True)
    sb[:, 0:3] = Vk + s * rng.normal(0, 0.05, (N, 3))
    sb[:, 3:6] = ba_true + s * rng.normal(0, 0.01, (N, 3))
    sb[:, 6:9] = bg_true + s * rng.normal(0, 0.001, (N, 3))
    ex = np.concatenate([tic, R_to_quat(ric)])
    # --- pre-integration, linearised at the initial bias estimate of frame i plus a small offset
    lin_ba = sb[:-1, 3:6] + s * rng.normal(0, 0.002, (N - 1, 3))
    lin_bg = sb[:-1, 6:9] + s * rng.normal(0, 0.0002, (N - 1, 3))
    idx = kf[:-1, None] + 1 + np.arange(SAMPLES)[None, :]
    imu = preintegrate(IMU_DT, acc_m[idx], gyr_m[idx], acc_m[kf[:-1]], gyr_m[kf[:-1]], lin_ba, lin_bg)
    # --- landmarks
    Rc = Rk @ ric                      # camera->world rotation per KF
    Pc = Pk + Rk @ tic
    if start is None:
        start = (np.arange(M) % max(N - 3, 1)).astype(np.int32)   # start_frame < WINDOW_SIZE - 2 (estimator.cpp:1192)
    start = np.asarray(start, np.int32)
    lm_w = np.zeros((M, 3)); depth = np.zeros(M); bear = np.zeros((M, 3))
    for f in range(M):
        for _ in range(100):
            u, vv = rng.uniform(0, cabi.IMG_W), rng.uniform(0, cabi.IMG_H)
            b = np.array([(u - cabi.CX) / cabi.FX, (vv - cabi.CY) / cabi.FY, 1.0])
            d = rng.uniform(3.0, 15.0)
            pw_ = Rc[start[f]] @ (b * d) + Pc[start[f]]
            zs = np.einsum("nji,nj->ni", Rc[start[f]:], pw_ - Pc[start[f]:])[:, 2]
            if np.all(zs > 0.5):
                break
        lm_w[f], depth[f], bear[f] = pw_, d, b
    pix_sigma = nz * 0.5 / cabi.FOCAL_LENGTH
    obs = {}   # (f, j) -> normalised xy
    for f in range(M):
        for j in range(start[f], N):
            pc = Rc[j].T @ (lm_w[f] - Pc[j])
            obs[(f, j)] = pc[:2] / pc[2] + rng.normal(0, 1.0, 2) * pix_sigma
    vel = {}
    for f in range(M):
        for j in range(start[f], N):
            if j > start[f]:
                vel[(f, j)] = (obs[(f, j)] - obs[(f, j - 1)]) / KF_DT
            else:
                vel[(f, j)] = (obs[(f, j + 1)] - obs[(f, j)]) / KF_DT
    kf_i, kf_j, feat, pts_i, pts_j, vel_i, vel_j, row_i, row_j = [], [], [], [], [], [], [], [], []
    for f in range(M):
        i = int(start[f])
        for j in range(i + 1, N):
            kf_i.append(i); kf_j.append(j); feat.append(f)
            pts_i.append([*obs[(f, i)], 1.0]); pts_j.append([*obs[(f, j)], 1.0])
            vel_i.append(vel[(f, i)]); vel_j.append(vel[(f, j)])
            row_i.append(cabi.FY * obs[(f, i)][1] + cabi.CY); row_j.append(cabi.FY * obs[(f, j)][1] + cabi.CY)
    n_proj = len(kf_i)
    depth_fixed = (rng.uniform(size=M) < 0.10).astype(np.uint8)
    lam_true = 1.0 / depth
    lam = np.where(depth_fixed == 1, lam_true, lam_true * (1 + s * rng.normal(0, 0.1, M)))
    w_ = dict(
        pose=pose, speedbias=sb, ex_pose=ex, inv_depth=lam, depth_fixed=depth_fixed, kf_fixed=np.zeros(N, np.uint8),
        td=cabi.TD0 + s * rng.normal(0, 1e-3),
        imu=imu, imu_kf=np.arange(N - 1, dtype=np.int32),
        pts_i=np.array(pts_i).reshape(n_proj, 3), pts_j=np.array(pts_j).reshape(n_proj, 3),
        vel_i=np.array(vel_i).reshape(n_proj, 2), vel_j=np.array(vel_j).reshape(n_proj, 2),
        td_i=np.full(n_proj, cabi.TD0), td_j=np.full(n_proj, cabi.TD0),
        row_i=np.array(row_i, dtype=np.float64), row_j=np.array(row_j, dtype=np.float64),
        kf_i=np.array(kf_i, np.int32), kf_j=np.array(kf_j, np.int32), feat=np.array(feat, np.int32),
        prior_n=0,
    )
    if n_lidar > 0:
        w_.update(_lidar_factors(rng, n_lidar, N, Rk, Pk, RLB, TLB, nz, lidar_kf0))
    tk = t0 + KF_DT * np.arange(N)
    icp, lps = [], []
    # anchors (first keyframe of each constraint): random unless the caller fixes them (config 4 keeps 3 + 3 clear of frame 0)
    icp_a = list(icp_anchors) if icp_anchors is not None else [int(rng.integers(0, N - 3)) for _ in range(n_icp)]
    lps_a = list(lps_anchors) if lps_anchors is not None else [int(rng.integers(0, N - 1)) for _ in range(n_lps)]
    for a_ in icp_a:   # constraint_mode 3 (estimator.cpp:1376-1395): sqrt_info = 100 / fitness, fitness 0.2
        c_ = a_ + 2
        ti, tj = tk[a_] + rng.uniform(0.01, 0.09), tk[c_] + rng.uniform(0.01, 0.09)
        si, sj = (ti - tk[a_]) / KF_DT, (tj - tk[c_]) / KF_DT
        Qi = quat_slerp(Qk[a_], Qk[a_ + 1], si); Qj = quat_slerp(Qk[c_], Qk[c_ + 1], sj)
        Pi = Pk[a_] + (Pk[a_ + 1] - Pk[a_]) * si; Pj = Pk[c_] + (Pk[c_ + 1] - Pk[c_]) * sj
        tm = quat_rot(quat_conj(Qi / np.linalg.norm(Qi)), Pj - Pi) + nz * rng.normal(0, 0.01, 3)
        icp.append(dict(t=(tk[a_], tk[a_ + 1], tk[c_], tk[c_ + 1], ti, tj), trans_t=tm, sqrt_info=100.0 / 0.2,
                        kf=(a_, a_ + 1, c_, c_ + 1)))
    for a_ in lps_a:
        tm_ = tk[a_] + rng.uniform(0.01, 0.09)
        Qi = quat_slerp(Qk[a_], Qk[a_ + 1], (tm_ - tk[a_]) / KF_DT)
        Qm = quat_mul(Qi / np.linalg.norm(Qi), small_quat(nz * rng.normal(0, 0.002, 3)))
        lps.append(dict(t=(tk[a_], tk[a_ + 1], tm_), q=Qm, kf=(a_, a_ + 1)))
    w_["icp"], w_["lps"] = icp, lps
    if ex_prior:
        # A window never runs without a marginalization prior in steady state (estimator.cpp:1171-1177); the camera
        # extrinsic and td are only weakly observable inside one second of motion and it is that prior which holds
        # them.  Stand-in: a diagonal MarginalizationFactor on [para_Ex_Pose, para_Td] linearised at the initial
        # value (sigma 1 cm / 0.005 rad / 1 ms), in the exact r = r_lin + J_lin (x [-] x0) form of
        # marginalization_factor.cpp:352-400.
        wts = np.array([100.0] * 3 + [200.0] * 3 + [1000.0])
        w_["prior_n"] = 7
        w_["prior_J"] = np.diag(wts).reshape(-1)       # column-major 7x7 (diagonal)
        w_["prior_r"] = np.zeros(7)
        w_["prior_blk"] = np.array([cabi.blk_id(cabi.VILS_BLK_EXPOSE, 0), cabi.blk_id(cabi.VILS_BLK_TD, 0)], np.int32)
        w_["prior_x0"] = np.concatenate([ex, [w_["td"]]])
    w_["raw"] = dict(ts=ts, acc=acc_m, gyr=gyr_m, kf=kf, obs=obs, vel=vel, start=start, depth=depth, ric=ric, tic=tic)
    w_["truth"] = dict(pose=np.concatenate([Pk, Qk], 1), speedbias=np.concatenate([Vk, np.tile(ba_true, (N, 1)), np.tile(bg_true, (N, 1))], 1),
                       inv_depth=lam_true, td=cabi.TD0, t_kf=tk)
    return w_


def make_config4_big(window_idx=0, N=20, M=300, n_lidar=5000, M0=18):
    """BASELINE configs[3] as SURVEY.md 8d states it: the (N+1)-frame window whose solve -> MARGIN_OLD marginalization -> slide leaves an
    N-frame window with exactly M landmarks (3333 projection factors at N=20, M=300), n_lidar LiDAR factors (3:1 plane:edge), 3 ICP + 3 LPS
    constraints and the REAL n = 6(N-1)+16 = 130-dim prior.  Frame 0 additionally carries M0 landmarks anchored there, one ICP and one LPS
    constraint (all marginalised, estimator.cpp:1489-1589); LiDAR point factors sit on frames 1..N only."""
    start = np.concatenate([np.zeros(M0, np.int32), 1 + (np.arange(M) % (N - 3))]).astype(np.int32)
    return make_window(4, window_idx, N=N + 1, M=M + M0, n_lidar=n_lidar, start=start, lidar_kf0=1,
                       icp_anchors=[0, 2, 7, 13], lps_anchors=[0, 3, 9, 16])


def attach_prior(w, prior):
    """prior: dict(n, J (n x n col-major flat), r, blk, x0) as produced by a marginalize() call."""
    w = dict(w)
    w["prior_n"] = int(prior["n"])
    w["prior_J"] = np.asarray(prior["J"], np.float64)
    w["prior_r"] = np.asarray(prior["r"], np.float64)
    w["prior_blk"] = np.asarray(prior["blk"], np.int32)
    w["prior_x0"] = np.asarray(prior["x0"], np.float64)
    return w


def slide_old(w):
    """Drop keyframe 0 of a window dict (what slideWindow() does after MARGIN_OLD, estimator.cpp:1689-1760):
    landmarks anchored at frame 0 are re-anchored... the reference moves their anchor to frame 1 through
    removeBackShiftDepth; here they are simply dropped with their factors, which keeps the synthetic window
    self-consistent.  Used only to build config 4 (a window carrying a real prior)."""
    N = w["pose"].shape[0]
    keepf = w["kf_i"] > 0
    feats = np.unique(w["feat"][keepf])
    remap = -np.ones(w["inv_depth"].shape[0], np.int64); remap[feats] = np.arange(len(feats))
    out = dict(w)
    for k in ["pose", "speedbias", "kf_fixed"]:
        out[k] = w[k][1:].copy()
    out["inv_depth"] = w["inv_depth"][feats].copy(); out["depth_fixed"] = w["depth_fixed"][feats].copy()
    out["imu"] = w["imu"][1:].copy(); out["imu_kf"] = (w["imu_kf"][1:] - 1).astype(np.int32)
    for k in ["pts_i", "pts_j", "vel_i", "vel_j", "td_i", "td_j", "row_i", "row_j"]:
        out[k] = w[k][keepf].copy()
    out["kf_i"] = (w["kf_i"][keepf] - 1).astype(np.int32); out["kf_j"] = (w["kf_j"][keepf] - 1).astype(np.int32)
    out["feat"] = remap[w["feat"][keepf]].astype(np.int32)
    if w.get("plane_kf") is not None:
        kp = w["plane_kf"] > 0; ke = w["edge_kf"] > 0
        for k in ["plane_p", "plane_n", "plane_d"]:
            out[k] = w[k][kp].copy()
        out["plane_kf"] = (w["plane_kf"][kp] - 1).astype(np.int32)
        for k in ["edge_p", "edge_a", "edge_b"]:
            out[k] = w[k][ke].copy()
        out["edge_kf"] = (w["edge_kf"][ke] - 1).astype(np.int32)
    out["icp"] = [dict(c, kf=tuple(x - 1 for x in c["kf"])) for c in w.get("icp", []) if min(c["kf"]) > 0]
    out["lps"] = [dict(c, kf=tuple(x - 1 for x in c["kf"])) for c in w.get("lps", []) if min(c["kf"]) > 0]
    out["prior_n"] = 0
    for k in ["prior_J", "prior_r", "prior_blk", "prior_x0"]:
        out.pop(k, None)
    if "truth" in w:
        t = w["truth"]
        out["truth"] = dict(pose=t["pose"][1:], speedbias=t["speedbias"][1:], inv_depth=t["inv_depth"][feats], td=t["td"], t_kf=t["t_kf"][1:])
    return out


def room_scan(rng, n, pose=None, noise=0.01):
    """A LiDAR-like scan of the 16 x 12 x 3 m room of SURVEY §8d seen from `pose` (4 x 4, sensor -> world): n x 4 float32 (x y z intensity)."""
    per = n // 6
    pts = []
    for axis, val in ((0, -8.0), (0, 8.0), (1, -6.0), (1, 6.0), (2, 0.0), (2, 3.0)):
        p = np.stack([rng.uniform(-8, 8, per), rng.uniform(-6, 6, per), rng.uniform(0, 3, per)], 1)
        p[:, axis] = val
        pts.append(p)
    w = np.concatenate(pts) + rng.normal(0, noise, (per * 6, 3))
    if pose is not None:
        w = (w - pose[:3, 3]) @ pose[:3, :3]
    return np.c_[w, rng.uniform(0, 100, len(w))].astype(np.float32)


def rigid(rotvec, trans):
    """4 x 4 rigid transform from a rotation vector (Rodrigues) and a translation."""
    r = np.asarray(rotvec, float); th = np.linalg.norm(r)
    K = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
    R = np.eye(3) if th < 1e-12 else np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * (K @ K)
    T = np.eye(4); T[:3, :3] = R; T[:3, 3] = trans
    return T
