"""ctypes mirror of include/vils_cabi.h.

Only struct layouts and small marshalling helpers live here; they are shared by the product binding
(`mvil_fusion_b200.lib`, which loads libvils_b200.so) and by the test-side oracle binding
binding in tests/, because the CPU checker deliberately reuses the header POD types.
"""
import ctypes as C

import numpy as np

VILS_OK = 0
VILS_ERR_BAD_ARG = 1
VILS_ERR_NO_DEVICE = 2
VILS_ERR_CUDA = 3
VILS_ERR_NOT_FINITE = 4
VILS_ERR_CHOLESKY = 5
VILS_ERR_CAPACITY = 6

VILS_BLK_POSE, VILS_BLK_SPEEDBIAS, VILS_BLK_EXPOSE, VILS_BLK_TD = 0, 1, 2, 3
VILS_MODE_GN, VILS_MODE_LM, VILS_MODE_DOGLEG = 0, 1, 2
VILS_MARGIN_OLD, VILS_MARGIN_SECOND_NEW = 0, 1


def blk_id(type_, idx):
    return (type_ << 16) | idx


def blk_type(bid):
    return bid >> 16


def blk_index(bid):
    return bid & 0xFFFF


def blk_local_size(type_):
    return {0: 6, 1: 9, 2: 6, 3: 1}[type_]


def blk_global_size(type_):
    return {0: 7, 1: 9, 2: 7, 3: 1}[type_]


c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_uint8_p = C.POINTER(C.c_uint8)
c_float_p = C.POINTER(C.c_float)


class VilsConfig(C.Structure):
    _fields_ = [
        ("focal_length", C.c_double),
        ("gravity", C.c_double * 3),
        ("tr", C.c_double),
        ("row", C.c_double),
        ("cauchy_visual", C.c_double),
        ("huber_lidar", C.c_double),
        ("rlb", C.c_double * 9),
        ("tlb", C.c_double * 3),
        ("estimate_extrinsic", C.c_int32),
        ("estimate_td", C.c_int32),
        ("max_kf", C.c_int32),
        ("max_feat", C.c_int32),
        ("max_proj", C.c_int32),
        ("max_lidar", C.c_int32),
        ("device", C.c_int32),
        ("reserved", C.c_int32),
        ("imu_noise", C.c_double * 4),
    ]


class VilsPreint(C.Structure):
    _fields_ = [
        ("delta_p", C.c_double * 3),
        ("delta_q", C.c_double * 4),
        ("delta_v", C.c_double * 3),
        ("lin_ba", C.c_double * 3),
        ("lin_bg", C.c_double * 3),
        ("sum_dt", C.c_double),
        ("jacobian", C.c_double * 225),
        ("covariance", C.c_double * 225),
    ]


PREINT_DOUBLES = 467
assert C.sizeof(VilsPreint) == PREINT_DOUBLES * 8


class VilsIcp(C.Structure):
    _fields_ = [
        ("ta", C.c_double), ("tb", C.c_double), ("tc", C.c_double), ("td", C.c_double),
        ("ti", C.c_double), ("tj", C.c_double),
        ("trans_t", C.c_double * 3),
        ("sqrt_info", C.c_double),
        ("kf", C.c_int32 * 4),
    ]


class VilsLps(C.Structure):
    _fields_ = [
        ("tl", C.c_double), ("tr", C.c_double), ("tk", C.c_double),
        ("q", C.c_double * 4),
        ("kf", C.c_int32 * 2),
    ]


class VilsWindow(C.Structure):
    _fields_ = [
        ("n_kf", C.c_int32), ("n_feat", C.c_int32), ("n_imu", C.c_int32), ("n_proj", C.c_int32),
        ("n_plane", C.c_int32), ("n_edge", C.c_int32), ("n_icp", C.c_int32), ("n_lps", C.c_int32),
        ("pose", c_double_p), ("speedbias", c_double_p), ("ex_pose", c_double_p), ("inv_depth", c_double_p),
        ("depth_fixed", c_uint8_p), ("kf_fixed", c_uint8_p),
        ("td", C.c_double),
        ("imu", C.POINTER(VilsPreint)), ("imu_kf", c_int32_p),
        ("pts_i", c_double_p), ("pts_j", c_double_p), ("vel_i", c_double_p), ("vel_j", c_double_p),
        ("td_i", c_double_p), ("td_j", c_double_p), ("row_i", c_double_p), ("row_j", c_double_p),
        ("kf_i", c_int32_p), ("kf_j", c_int32_p), ("feat", c_int32_p),
        ("plane_p", c_double_p), ("plane_n", c_double_p), ("plane_d", c_double_p), ("plane_kf", c_int32_p),
        ("edge_p", c_double_p), ("edge_a", c_double_p), ("edge_b", c_double_p), ("edge_kf", c_int32_p),
        ("icp", C.POINTER(VilsIcp)), ("lps", C.POINTER(VilsLps)),
        ("prior_n", C.c_int32), ("prior_nblk", C.c_int32),
        ("prior_J", c_double_p), ("prior_r", c_double_p), ("prior_blk", c_int32_p), ("prior_x0", c_double_p),
    ]


class VilsSolveOpts(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("max_iters", C.c_int32),
        ("mu", C.c_double), ("lm_initial_radius", C.c_double), ("function_tolerance", C.c_double),
        ("parameter_tolerance", C.c_double), ("min_relative_decrease", C.c_double), ("max_solver_time", C.c_double),
    ]


class VilsSummary(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("iterations", C.c_int32), ("accepted", C.c_int32), ("reserved", C.c_int32),
        ("cost_initial", C.c_double), ("cost_final", C.c_double),
    ]


class VilsPriorOut(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("nblk", C.c_int32), ("m", C.c_int32), ("capacity_n", C.c_int32),
        ("J", c_double_p), ("r", c_double_p), ("blk", c_int32_p), ("x0", c_double_p),
    ]


# ---------------------------------------------------------------------------------------------
# Constants of config/mynteye_leishen_indoor.yaml + parameters.h (the reference's defaults).
# ---------------------------------------------------------------------------------------------
RIC = np.array([0.99999072, -0.00209387, -0.00376471, -0.00208308, -0.99999371, 0.0028693,
                -0.0037707, -0.00286143, -0.9999888]).reshape(3, 3)   # yaml:31-35 imu^R_cam
TIC = np.array([-0.04571386, 0.01268073, -0.01535602])               # yaml:36-40
GT_RLI = np.array([-0.0320631, 0.000946093, -0.999485, -0.999482, -0.00274554, 0.0320604,
                   -0.0027138, 0.999996, 0.00103363]).reshape(3, 3)  # yaml:43-47
GT_TLI = np.array([0.2, -0.005, -0.1])                                # yaml:48-52
FX, FY, CX, CY = 356.37000498, 354.92225534, 326.87903275, 250.93806883  # yaml:18-22
IMG_W, IMG_H = 640, 480
ACC_N, GYR_N, ACC_W, GYR_W = 0.02065, 0.00519, 0.00667, 0.00088056   # yaml:81-86
G_NORM = 9.795                                                        # yaml:102
TD0 = 0.00003                                                         # yaml:113
FOCAL_LENGTH = 460.0                                                  # parameters.h:11


def _orthonormalize(R):
    """parameters.cpp:143-145 — Quaterniond(R).normalized() back to a rotation matrix (nearest rotation)."""
    u, _, vt = np.linalg.svd(R)
    Rn = u @ vt
    if np.linalg.det(Rn) < 0:
        u[:, -1] *= -1
        Rn = u @ vt
    return Rn


def default_config(max_kf=10, max_feat=150, max_proj=1400, max_lidar=2000, device=0):
    cfg = VilsConfig()
    cfg.focal_length = FOCAL_LENGTH
    cfg.gravity[:] = [0.0, 0.0, G_NORM]
    cfg.tr = 0.0
    cfg.row = float(IMG_H)
    cfg.cauchy_visual = 1.0
    cfg.huber_lidar = 0.1
    cfg.rlb[:] = list(_orthonormalize(GT_RLI).reshape(-1))
    cfg.tlb[:] = list(GT_TLI)
    cfg.estimate_extrinsic = 1
    cfg.estimate_td = 1
    cfg.max_kf, cfg.max_feat, cfg.max_proj, cfg.max_lidar = max_kf, max_feat, max_proj, max_lidar
    cfg.imu_noise[:] = [ACC_N, GYR_N, ACC_W, GYR_W]
    cfg.device = device
    return cfg


def default_solve_opts(mode=VILS_MODE_GN, max_iters=5, mu=1e-8):
    o = VilsSolveOpts()
    o.mode, o.max_iters, o.mu = mode, max_iters, mu
    o.lm_initial_radius = 1e4
    o.function_tolerance = 1e-6
    o.parameter_tolerance = 1e-8
    o.min_relative_decrease = 1e-3
    o.max_solver_time = 0.0
    return o


def _dp(a):
    return a.ctypes.data_as(c_double_p) if a is not None and a.size else None


def _ip(a):
    return a.ctypes.data_as(c_int32_p) if a is not None and a.size else None


def _up(a):
    return a.ctypes.data_as(c_uint8_p) if a is not None and a.size else None


_F64 = ["pose", "speedbias", "ex_pose", "inv_depth", "pts_i", "pts_j", "vel_i", "vel_j", "td_i", "td_j", "row_i",
        "row_j", "plane_p", "plane_n", "plane_d", "edge_p", "edge_a", "edge_b", "prior_J", "prior_r", "prior_x0"]
_I32 = ["imu_kf", "kf_i", "kf_j", "feat", "plane_kf", "edge_kf", "prior_blk"]
_U8 = ["depth_fixed", "kf_fixed"]


def window_struct(w):
    """dict of numpy arrays (see synth.make_window) -> (VilsWindow, keepalive list)."""
    keep = []
    s = VilsWindow()

    def arr(name, dtype):
        a = w.get(name)
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=dtype)
        keep.append(a)
        return a

    for k in _F64:
        setattr(s, k, _dp(arr(k, np.float64)))
    for k in _I32:
        setattr(s, k, _ip(arr(k, np.int32)))
    for k in _U8:
        setattr(s, k, _up(arr(k, np.uint8)))
    s.n_kf = int(w["pose"].shape[0])
    s.n_feat = int(w["inv_depth"].shape[0])
    s.td = float(w["td"])
    imu = arr("imu", np.float64)  # (n_imu, 467)
    s.n_imu = 0 if imu is None else int(imu.shape[0])
    if s.n_imu:
        assert imu.shape[1] == PREINT_DOUBLES
        s.imu = C.cast(imu.ctypes.data, C.POINTER(VilsPreint))
    s.n_proj = 0 if w.get("kf_i") is None else int(np.asarray(w["kf_i"]).shape[0])
    s.n_plane = 0 if w.get("plane_kf") is None else int(np.asarray(w["plane_kf"]).shape[0])
    s.n_edge = 0 if w.get("edge_kf") is None else int(np.asarray(w["edge_kf"]).shape[0])
    icp = w.get("icp") or []
    lps = w.get("lps") or []
    s.n_icp, s.n_lps = len(icp), len(lps)
    if icp:
        a = (VilsIcp * len(icp))()
        for k, c in enumerate(icp):
            a[k].ta, a[k].tb, a[k].tc, a[k].td, a[k].ti, a[k].tj = c["t"]
            a[k].trans_t[:] = list(c["trans_t"])
            a[k].sqrt_info = c["sqrt_info"]
            a[k].kf[:] = list(c["kf"])
        keep.append(a)
        s.icp = a
    if lps:
        a = (VilsLps * len(lps))()
        for k, c in enumerate(lps):
            a[k].tl, a[k].tr, a[k].tk = c["t"]
            a[k].q[:] = list(c["q"])
            a[k].kf[:] = list(c["kf"])
        keep.append(a)
        s.lps = a
    s.prior_n = int(w.get("prior_n", 0))
    s.prior_nblk = 0 if w.get("prior_blk") is None else int(np.asarray(w["prior_blk"]).shape[0])
    return s, keep


def residual_count(w):
    n = lambda k: 0 if w.get(k) is None else len(w[k])
    return (15 * n("imu") + 2 * n("kf_i") + n("plane_kf") + 3 * n("edge_kf") + 3 * len(w.get("icp") or [])
            + 3 * len(w.get("lps") or []) + int(w.get("prior_n", 0)))


def jacobian_count(w):
    n = lambda k: 0 if w.get(k) is None else len(w[k])
    return (450 * n("imu") + 40 * n("kf_i") + 6 * n("plane_kf") + 18 * n("edge_kf") + 72 * len(w.get("icp") or [])
            + 36 * len(w.get("lps") or []))


class VilsVgicpOpts(C.Structure):
    """vils_vgicp_opts (include/vils_cabi.h)."""
    _fields_ = [("resolution", C.c_double), ("rotation_epsilon", C.c_double), ("transformation_epsilon", C.c_double), ("lm_init_lambda_factor", C.c_double),
                ("k_correspondences", C.c_int32), ("neighbor_search", C.c_int32), ("max_iterations", C.c_int32), ("lm_max_iterations", C.c_int32),
                ("compute_fitness", C.c_int32), ("reserved", C.c_int32)]


class VilsVgicpResult(C.Structure):
    """vils_vgicp_result (include/vils_cabi.h)."""
    _fields_ = [("T", C.c_double * 16), ("H", C.c_double * 36), ("error", C.c_double), ("fitness", C.c_double), ("iterations", C.c_int32), ("converged", C.c_int32),
                ("n_corr", C.c_int32), ("n_voxels", C.c_int32), ("n_linearize", C.c_int32), ("elapsed_ms", C.c_float)]
