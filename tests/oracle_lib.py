"""ctypes binding of oracle/_build/libvils_oracle.so — TEST INFRASTRUCTURE ONLY (see oracle/README.md)."""
import ctypes as C
import os
import subprocess

import numpy as np

from mvil_fusion_b200 import cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "_build", "libvils_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
        _lib = C.CDLL(_SO)
    return _lib


def _d(a):
    return a.ctypes.data_as(cabi.c_double_p)


def evaluate_window(cfg, w, apply_loss=True):
    ws, keep = cabi.window_struct(w)
    r = np.zeros(cabi.residual_count(w)); J = np.zeros(max(cabi.jacobian_count(w), 1)); cost = C.c_double()
    lib().vo_evaluate_window(C.byref(cfg), C.byref(ws), int(apply_loss), _d(r), _d(J), C.byref(cost))
    return r, J[:cabi.jacobian_count(w)], cost.value


def linearize_window(cfg, w):
    ws, keep = cabi.window_struct(w)
    D = 15 * ws.n_kf + 7
    S = np.zeros((D, D)); g = np.zeros(D); cost = C.c_double()
    lib().vo_linearize_window(C.byref(cfg), C.byref(ws), _d(S), _d(g), C.byref(cost))
    return S, g, cost.value


def solve_window(cfg, w, opts):
    ws, keep = cabi.window_struct(w)
    N, M = ws.n_kf, ws.n_feat
    pose = np.zeros((N, 7)); sb = np.zeros((N, 9)); ex = np.zeros(7); lam = np.zeros(max(M, 1)); td = C.c_double()
    summ = cabi.VilsSummary()
    st = lib().vo_solve_window(C.byref(cfg), C.byref(ws), C.byref(opts), _d(pose), _d(sb), _d(ex), _d(lam), C.byref(td), C.byref(summ))
    return dict(status=st, pose=pose, speedbias=sb, ex_pose=ex, inv_depth=lam[:M], td=td.value,
                iterations=summ.iterations, accepted=summ.accepted, cost_initial=summ.cost_initial, cost_final=summ.cost_final)


def marginalize_window(cfg, w, flag, capacity_n=512):
    ws, keep = cabi.window_struct(w)
    out = cabi.VilsPriorOut()
    J = np.zeros(capacity_n * capacity_n); r = np.zeros(capacity_n); blk = np.zeros(64, np.int32); x0 = np.zeros(64 * 9)
    out.capacity_n = capacity_n
    out.J, out.r, out.x0 = _d(J), _d(r), _d(x0)
    out.blk = blk.ctypes.data_as(cabi.c_int32_p)
    st = lib().vo_marginalize_window(C.byref(cfg), C.byref(ws), int(flag), C.byref(out))
    n, nb = out.n, out.nblk
    gs = sum(cabi.blk_global_size(cabi.blk_type(int(b))) for b in blk[:nb])
    return dict(status=st, n=n, m=out.m, J=J[:n * n].copy(), r=r[:n].copy(), blk=blk[:nb].copy(), x0=x0[:gs].copy())


def double2vector(pose0_before, pose, sb):
    pose = np.ascontiguousarray(pose, np.float64).copy(); sb = np.ascontiguousarray(sb, np.float64).copy()
    p0 = np.ascontiguousarray(pose0_before, np.float64)
    lib().vo_double2vector(int(pose.shape[0]), _d(p0), _d(pose), _d(sb))
    return pose, sb


def preintegrate(off, dt, acc, gyr, acc0, gyr0, ba, bg, noise):
    K = len(off) - 1
    out = np.zeros((K, cabi.PREINT_DOUBLES))
    off = np.ascontiguousarray(off, np.int32)
    args = [np.ascontiguousarray(a, np.float64) for a in (dt, acc, gyr, acc0, gyr0, ba, bg, noise)]
    lib().vo_preintegrate(K, off.ctypes.data_as(cabi.c_int32_p), *[_d(a) for a in args], C.cast(out.ctypes.data, C.POINTER(cabi.VilsPreint)))
    return out


def deskew(xyzi, stride, q, t, time_factor, min_r, max_r):
    a = np.ascontiguousarray(xyzi, np.float32).copy()
    q = np.ascontiguousarray(q, np.float32); t = np.ascontiguousarray(t, np.float32)
    lib().vo_deskew.argtypes = [cabi.c_float_p, C.c_int, C.c_int, cabi.c_float_p, cabi.c_float_p, C.c_float, C.c_double, C.c_double]
    lib().vo_deskew(a.ctypes.data_as(cabi.c_float_p), a.size // stride, stride, q.ctypes.data_as(cabi.c_float_p),
                    t.ctypes.data_as(cabi.c_float_p), float(time_factor), float(min_r), float(max_r))
    return a


def stamp_rings(xyzi, stride, lower_deg=-15.0, upper_deg=15.0, n_rings=16, scan_period=0.1):
    a = np.ascontiguousarray(xyzi, np.float32).copy()
    n = a.size // stride
    ring = np.zeros(n, np.int32)
    lib().vo_stamp_rings.argtypes = [cabi.c_float_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_float, cabi.c_int32_p]
    lib().vo_stamp_rings(a.ctypes.data_as(cabi.c_float_p), n, stride, lower_deg, upper_deg, n_rings, scan_period,
                         ring.ctypes.data_as(cabi.c_int32_p))
    return a, ring
