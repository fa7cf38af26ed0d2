"""Python binding of libvils_b200.so (the C-ABI of include/vils_cabi.h) — what tests/ and bench.py call.

The reference is a C++ system; its host side is mirrored in C++ (csrc/host/estimator.h, feature_tracker.h).  This
module is the thin ctypes view of the same C-ABI used for parity tests and measurement.  It never falls back to a CPU
implementation: if the shared library is missing or no sm_100 device is present, calls raise.
"""
import ctypes as C
import os

import numpy as np

from . import cabi

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("VILS_SO") or os.path.join(HERE, "libvils_b200.so")   # VILS_SO: experiment builds (tools/ only)
_lib = None


class VilsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libvils_b200 error {code}: {msg}")
        self.code = code


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(f"{SO_PATH} is missing: run `python -m mvil_fusion_b200.build` (there is no CPU fallback)")
    L = C.CDLL(SO_PATH)
    L.vils_last_error.restype = C.c_char_p
    vp = C.c_void_p
    dp, ip, up, fp = cabi.c_double_p, cabi.c_int32_p, cabi.c_uint8_p, cabi.c_float_p
    L.vils_ba_create.argtypes = [C.POINTER(cabi.VilsConfig), C.c_int32, C.POINTER(vp)]
    L.vils_ba_destroy.argtypes = [vp]
    L.vils_ba_destroy.restype = None
    L.vils_ba_set_window.argtypes = [vp, C.c_int32, C.POINTER(cabi.VilsWindow)]
    L.vils_ba_set_windows.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(cabi.VilsWindow)]
    L.vils_ba_solve_windows.argtypes = [vp, C.c_int32, C.POINTER(cabi.VilsWindow), C.POINTER(cabi.VilsSolveOpts)]
    L.vils_ba_put_state.argtypes = [vp, C.c_int32, dp, dp, dp, dp, C.c_double]
    L.vils_ba_upload.argtypes = [vp, C.c_int32]
    L.vils_ba_solve_device.argtypes = [vp, C.c_int32, C.POINTER(cabi.VilsSolveOpts)]
    L.vils_ba_download.argtypes = [vp, C.c_int32]
    L.vils_ba_solve.argtypes = [vp, C.c_int32, C.POINTER(cabi.VilsSolveOpts)]
    L.vils_ba_get_state.argtypes = [vp, C.c_int32, dp, dp, dp, dp, dp, C.POINTER(cabi.VilsSummary)]
    L.vils_double2vector.argtypes = [C.c_int32, dp, dp, dp]
    L.vils_ba_evaluate.argtypes = [vp, C.c_int32, C.c_int32, dp, dp]
    L.vils_ba_evaluate_device.argtypes = [vp, C.c_int32, C.c_int32]
    L.vils_ba_linearize.argtypes = [vp, C.c_int32, dp, dp, dp]
    L.vils_ba_marginalize.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(cabi.VilsPriorOut)]
    L.vils_ba_last_device_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.vils_ba_set_cluster.argtypes = [vp, C.c_int32]
    L.vils_ba_last_cluster.argtypes = [vp, C.POINTER(C.c_int32)]
    L.vils_ba_last_launches.argtypes = [vp, C.POINTER(C.c_int32)]
    L.vils_ba_last_transfer_bytes.argtypes = [vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.vils_ba_sharded_buffer.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.vils_ba_sharded_linearize.argtypes = [vp, C.c_int32, C.POINTER(cabi.VilsSolveOpts)]
    L.vils_ba_sharded_read.argtypes = [vp, dp]
    L.vils_ba_sharded_write.argtypes = [vp, dp]
    L.vils_ba_sharded_update.argtypes = [vp, C.POINTER(cabi.VilsSolveOpts)]
    L.vils_nccl_unique_id.argtypes = [up]
    L.vils_ba_sharded_init.argtypes = [vp, C.c_int32, C.c_int32, up]
    L.vils_ba_sharded_solve.argtypes = [vp, C.POINTER(cabi.VilsSolveOpts), C.POINTER(cabi.VilsSummary)]
    L.vils_preintegrate.argtypes = [C.c_int32, ip, dp, dp, dp, dp, dp, dp, dp, dp, C.POINTER(cabi.VilsPreint), C.c_int32]
    L.vils_klt_create.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(vp)]
    L.vils_klt_destroy.argtypes = [vp]
    L.vils_klt_destroy.restype = None
    L.vils_klt_track.argtypes = [vp, up, up, C.c_int32, fp, C.c_int32, fp, up, fp]
    L.vils_klt_upload.argtypes = [vp, up, up, C.c_int32, fp, C.c_int32]
    L.vils_klt_track_device.argtypes = [vp]
    L.vils_klt_download.argtypes = [vp, fp, up, fp]
    L.vils_klt_last_device_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.vils_klt_advance.argtypes = [vp, vp, C.c_int32, vp, fp, C.c_int32, fp, up, fp]
    L.vils_frontend_load.argtypes = [vp, up, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_int32]
    L.vils_frontend_current.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_int32), C.POINTER(vp)]
    L.vils_good_features_resident.argtypes = [vp, C.c_int32, C.c_double, C.c_double, C.c_int32, fp, ip]
    L.vils_frontend_create.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(vp)]
    L.vils_frontend_destroy.argtypes = [vp]
    L.vils_frontend_destroy.restype = None
    L.vils_clahe.argtypes = [vp, up, C.c_int32, C.c_double, C.c_int32, C.c_int32, up, C.c_int32]
    L.vils_set_mask.argtypes = [vp, fp, ip, C.c_int32, C.c_int32, ip, ip]
    L.vils_get_mask.argtypes = [vp, up, C.c_int32]
    L.vils_good_features.argtypes = [vp, up, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32, fp, ip]
    L.vils_frontend_get_eig.argtypes = [vp, fp]
    L.vils_lift_projective.argtypes = [vp, dp, fp, C.c_int32, dp]
    L.vils_frontend_last_device_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.vils_reject_with_f.argtypes = [vp, fp, fp, C.c_int32, C.c_double, up, dp]
    L.vils_triangulate.argtypes = [C.c_int32, ip, ip, dp, C.c_int32, dp, dp, dp, dp, C.c_double, dp, C.c_int32]
    L.vils_lidar_associate.argtypes = [fp, C.c_int32, fp, C.c_int32, dp, dp, C.c_int32, dp, up, ip, C.POINTER(C.c_float), C.c_int32]
    L.vils_depth_register.argtypes = [fp, C.c_int32, fp, fp, C.c_int32, fp, C.c_int32, fp, C.POINTER(C.c_float), C.c_int32]
    L.vils_vgicp_default_opts.argtypes = [C.POINTER(cabi.VilsVgicpOpts)]
    L.vils_vgicp_default_opts.restype = None
    L.vils_vgicp_covariances.argtypes = [fp, C.c_int32, dp, ip, C.c_int32]
    L.vils_vgicp_linearize.argtypes = [fp, C.c_int32, fp, C.c_int32, dp, C.POINTER(cabi.VilsVgicpOpts), dp, dp, dp, ip, ip, dp, C.c_int32]
    L.vils_vgicp_align.argtypes = [fp, C.c_int32, fp, C.c_int32, dp, C.POINTER(cabi.VilsVgicpOpts), C.POINTER(cabi.VilsVgicpResult), C.c_int32]
    L.vils_deskew.argtypes = [fp, C.c_int32, C.c_int32, fp, fp, C.c_float, C.c_float, C.c_float, C.c_int32]
    L.vils_stamp_rings.argtypes = [fp, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int32, C.c_float, ip, C.c_int32]
    L.vils_point_to_ring.argtypes = [fp, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int32, C.c_float, fp, ip, C.c_int32]
    L.vils_lidar_dev_alloc.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(vp)]
    L.vils_lidar_dev_upload.argtypes = [vp, fp]
    L.vils_lidar_dev_deskew.argtypes = [vp, fp, fp, C.c_float, C.c_float, C.c_float, C.POINTER(C.c_float)]
    L.vils_lidar_dev_download.argtypes = [vp, fp]
    L.vils_lidar_dev_free.argtypes = [vp]
    L.vils_lidar_dev_free.restype = None
    _lib = L
    return L


def _check(code):
    if code != 0:
        raise VilsError(code, load().vils_last_error().decode())


def _d(a):
    return a.ctypes.data_as(cabi.c_double_p)


class BA:
    """One vils_ba handle: up to max_windows sliding windows resident on one GPU."""

    def __init__(self, cfg=None, max_windows=1):
        self.L = load()
        self.cfg = cfg or cabi.default_config()
        self.h = C.c_void_p()
        _check(self.L.vils_ba_create(C.byref(self.cfg), max_windows, C.byref(self.h)))
        self.max_windows = max_windows
        self._dims = {}

    def close(self):
        if self.h:
            self.L.vils_ba_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_window(self, slot, w):
        ws, keep = cabi.window_struct(w)
        _check(self.L.vils_ba_set_window(self.h, slot, C.byref(ws)))
        self._dims[slot] = (ws.n_kf, ws.n_feat, cabi.residual_count(w), cabi.jacobian_count(w))

    @staticmethod
    def window_array(ws):
        """list of window dicts -> (contiguous VilsWindow array, keepalive): the caller-side arrays handed to the batched calls."""
        arr = (cabi.VilsWindow * len(ws))()
        keep = []
        for k, w in enumerate(ws):
            s, kp = cabi.window_struct(w)
            arr[k] = s
            keep.append((s, kp))
        return arr, keep

    def _note_dims(self, slot0, ws):
        # lazily: the per-window residual / Jacobian counts are only needed by get_state / evaluate, and computing them for hundreds of windows
        # costs more Python time than the library call itself
        for k, w in enumerate(ws):
            self._dims[slot0 + k] = w

    def set_windows(self, slot0, ws, arr=None):
        """vils_ba_set_windows: pack len(ws) windows on all host threads."""
        if arr is None:
            arr, keep = self.window_array(ws)
        _check(self.L.vils_ba_set_windows(self.h, slot0, len(ws), arr))
        self._note_dims(slot0, ws)

    def solve_windows(self, ws, opts, arr=None):
        """vils_ba_solve_windows: caller arrays -> solved states (pack | H2D | solve | D2H pipelined)."""
        if arr is None:
            arr, keep = self.window_array(ws)
        _check(self.L.vils_ba_solve_windows(self.h, len(ws), arr, C.byref(opts)))
        self._note_dims(0, ws)

    def put_state(self, slot, pose, sb, ex, lam, td):
        a = [np.ascontiguousarray(v, np.float64) for v in (pose, sb, ex, lam)]
        _check(self.L.vils_ba_put_state(self.h, slot, _d(a[0]), _d(a[1]), _d(a[2]), _d(a[3]) if a[3].size else None, float(td)))

    def set_cluster(self, size):
        """Latency mode: 0 auto, 1 off, 2 / 4 / 8 / 16 SMs per window (vils_ba_set_cluster)."""
        _check(self.L.vils_ba_set_cluster(self.h, int(size)))

    @property
    def last_cluster(self):
        n = C.c_int32()
        self.L.vils_ba_last_cluster(self.h, C.byref(n))
        return n.value

    def upload(self, n):
        _check(self.L.vils_ba_upload(self.h, n))

    def solve_device(self, n, opts):
        _check(self.L.vils_ba_solve_device(self.h, n, C.byref(opts)))

    def download(self, n):
        _check(self.L.vils_ba_download(self.h, n))

    def solve(self, n, opts):
        _check(self.L.vils_ba_solve(self.h, n, C.byref(opts)))

    def _dim(self, slot):
        d = self._dims[slot]
        if isinstance(d, dict):
            d = (int(d["pose"].shape[0]), int(d["inv_depth"].shape[0]), cabi.residual_count(d), cabi.jacobian_count(d))
            self._dims[slot] = d
        return d

    def get_state(self, slot):
        N, M, _, _ = self._dim(slot)
        pose = np.zeros((N, 7)); sb = np.zeros((N, 9)); ex = np.zeros(7); lam = np.zeros(max(M, 1)); td = C.c_double()
        s = cabi.VilsSummary()
        st = self.L.vils_ba_get_state(self.h, slot, _d(pose), _d(sb), _d(ex), _d(lam), C.byref(td), C.byref(s))
        return dict(status=st, pose=pose, speedbias=sb, ex_pose=ex, inv_depth=lam[:M], td=td.value, iterations=s.iterations,
                    accepted=s.accepted, cost_initial=s.cost_initial, cost_final=s.cost_final)

    def evaluate(self, slot, apply_loss=True):
        _, _, nr, nj = self._dim(slot)
        r = np.zeros(nr); J = np.zeros(max(nj, 1))
        _check(self.L.vils_ba_evaluate(self.h, slot, int(apply_loss), _d(r), _d(J)))
        return r, J[:nj]

    def evaluate_device(self, n, apply_loss=True):
        _check(self.L.vils_ba_evaluate_device(self.h, n, int(apply_loss)))

    def linearize(self, slot):
        N = self._dim(slot)[0]
        D = 15 * N + 7
        S = np.zeros((D, D)); g = np.zeros(D); cost = C.c_double()
        _check(self.L.vils_ba_linearize(self.h, slot, _d(S), _d(g), C.byref(cost)))
        return S, g, cost.value

    def marginalize(self, slot, flag, capacity_n=512):
        out = cabi.VilsPriorOut()
        J = np.zeros(capacity_n * capacity_n); r = np.zeros(capacity_n); blk = np.zeros(80, np.int32); x0 = np.zeros(80 * 9)
        out.capacity_n = capacity_n
        out.J, out.r, out.x0 = _d(J), _d(r), _d(x0)
        out.blk = blk.ctypes.data_as(cabi.c_int32_p)
        _check(self.L.vils_ba_marginalize(self.h, slot, int(flag), C.byref(out)))
        n, nb = out.n, out.nblk
        gs = sum(cabi.blk_global_size(cabi.blk_type(int(b))) for b in blk[:nb])
        return dict(n=n, m=out.m, J=J[:n * n].copy(), r=r[:n].copy(), blk=blk[:nb].copy(), x0=x0[:gs].copy())

    # ---- factor-sharded mode (one window over several GPUs) ----
    def sharded_buffer(self):
        """(device pointer, n_doubles) of the partial system [H | g | hd | cost] the caller all-reduces."""
        ptr = C.c_void_p(); n = C.c_size_t()
        _check(self.L.vils_ba_sharded_buffer(self.h, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def sharded_buffer_tensor(self):
        """The same buffer as a torch CUDA tensor view (for torch.distributed.all_reduce over NCCL)."""
        import torch
        ptr, n = self.sharded_buffer()

        class _Arr:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}
        return torch.as_tensor(_Arr(), device=f"cuda:{self.cfg.device}")

    def sharded_init(self, rank, nranks, uid):
        """uid: the 128 bytes of nccl_unique_id() from rank 0."""
        a = np.frombuffer(bytes(uid), np.uint8).copy()
        _check(self.L.vils_ba_sharded_init(self.h, rank, nranks, a.ctypes.data_as(cabi.c_uint8_p)))

    def sharded_solve(self, opts):
        """All GN iterations natively: linearise -> ncclAllReduce -> update on the library stream, one host sync."""
        s = cabi.VilsSummary()
        st = self.L.vils_ba_sharded_solve(self.h, C.byref(opts), C.byref(s))
        if st not in (0, cabi.VILS_ERR_CHOLESKY, cabi.VILS_ERR_NOT_FINITE):
            _check(st)
        return s

    def sharded_linearize(self, iteration, opts):
        _check(self.L.vils_ba_sharded_linearize(self.h, iteration, C.byref(opts)))

    def sharded_update(self, opts):
        _check(self.L.vils_ba_sharded_update(self.h, C.byref(opts)))

    def sharded_read(self):
        _, n = self.sharded_buffer()
        a = np.zeros(n)
        _check(self.L.vils_ba_sharded_read(self.h, _d(a)))
        return a

    def sharded_write(self, a):
        a = np.ascontiguousarray(a, np.float64)
        _check(self.L.vils_ba_sharded_write(self.h, _d(a)))

    @property
    def last_ms(self):
        ms = C.c_float()
        self.L.vils_ba_last_device_ms(self.h, C.byref(ms))
        return ms.value

    @property
    def last_transfer_bytes(self):
        a, b = C.c_size_t(), C.c_size_t()
        self.L.vils_ba_last_transfer_bytes(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    @property
    def last_launches(self):
        n = C.c_int32()
        self.L.vils_ba_last_launches(self.h, C.byref(n))
        return n.value


def nccl_unique_id():
    a = np.zeros(128, np.uint8)
    _check(load().vils_nccl_unique_id(a.ctypes.data_as(cabi.c_uint8_p)))
    return a.tobytes()


def double2vector(pose0_before, pose, sb):
    pose = np.ascontiguousarray(pose, np.float64).copy(); sb = np.ascontiguousarray(sb, np.float64).copy()
    p0 = np.ascontiguousarray(pose0_before, np.float64)
    _check(load().vils_double2vector(int(pose.shape[0]), _d(p0), _d(pose), _d(sb)))
    return pose, sb


def preintegrate(off, dt, acc, gyr, acc0, gyr0, ba, bg, noise, device=0):
    K = len(off) - 1
    out = np.zeros((K, cabi.PREINT_DOUBLES))
    off = np.ascontiguousarray(off, np.int32)
    a = [np.ascontiguousarray(x, np.float64) for x in (dt, acc, gyr, acc0, gyr0, ba, bg, noise)]
    _check(load().vils_preintegrate(K, off.ctypes.data_as(cabi.c_int32_p), *[_d(x) for x in a],
                                    C.cast(out.ctypes.data, C.POINTER(cabi.VilsPreint)), device))
    return out


def deskew(xyzi, stride, q, t, time_factor, min_r, max_r, device=0):
    a = np.ascontiguousarray(xyzi, np.float32).copy()
    q = np.ascontiguousarray(q, np.float32); t = np.ascontiguousarray(t, np.float32)
    fp = cabi.c_float_p
    _check(load().vils_deskew(a.ctypes.data_as(fp), a.size // stride, stride, q.ctypes.data_as(fp), t.ctypes.data_as(fp),
                              float(time_factor), float(min_r), float(max_r), device))
    return a


def stamp_rings(xyzi, stride, lower_deg=-15.0, upper_deg=15.0, n_rings=16, scan_period=0.1, device=0):
    a = np.ascontiguousarray(xyzi, np.float32).copy()
    n = a.size // stride
    ring = np.zeros(n, np.int32)
    _check(load().vils_stamp_rings(a.ctypes.data_as(cabi.c_float_p), n, stride, lower_deg, upper_deg, n_rings, scan_period,
                                   ring.ctypes.data_as(cabi.c_int32_p), device))
    return a, ring


def point_to_ring(xyzi, stride, lower_deg=-15.0, upper_deg=15.0, n_rings=16, scan_period=0.1, device=0):
    """vils_point_to_ring: returns (ring-major cloud of the kept points, ring_start[n_rings + 1])."""
    a = np.ascontiguousarray(xyzi, np.float32)
    n = a.size // stride
    out = np.zeros((max(n, 1), stride), np.float32); start = np.zeros(n_rings + 1, np.int32)
    _check(load().vils_point_to_ring(a.ctypes.data_as(cabi.c_float_p), n, stride, lower_deg, upper_deg, n_rings, scan_period,
                                     out.ctypes.data_as(cabi.c_float_p), start.ctypes.data_as(cabi.c_int32_p), device))
    return out[:start[-1]], start


class KLT:
    """cv::calcOpticalFlowPyrLK replacement (feature_tracker_/src/feature_tracker.cpp:113)."""

    def __init__(self, rows, cols, max_pts=512, win=21, max_level=3, device=0):
        self.L = load()
        self.h = C.c_void_p()
        _check(self.L.vils_klt_create(rows, cols, max_pts, win, max_level, device, C.byref(self.h)))
        self.rows, self.cols = rows, cols

    def close(self):
        if self.h:
            self.L.vils_klt_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _rows(img):
        """u8 image whose rows are contiguous; the row pitch may exceed the width (cv::Mat::step > cols): no copy is made for such views."""
        img = np.asarray(img)
        if img.dtype != np.uint8 or img.ndim != 2 or img.strides[1] != 1 or img.strides[0] < img.shape[1]:
            img = np.ascontiguousarray(img, np.uint8)
        return img

    def track(self, prev, nxt, pts):
        prev = self._rows(prev); nxt = self._rows(nxt)
        if prev.strides[0] != nxt.strides[0]:
            prev = np.ascontiguousarray(prev); nxt = np.ascontiguousarray(nxt)
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
        n = pts.shape[0]
        out = np.zeros((n, 2), np.float32); status = np.zeros(n, np.uint8); err = np.zeros(n, np.float32)
        up, fp = cabi.c_uint8_p, cabi.c_float_p
        _check(self.L.vils_klt_track(self.h, prev.ctypes.data_as(up), nxt.ctypes.data_as(up), prev.strides[0], pts.ctypes.data_as(fp), n,
                                     out.ctypes.data_as(fp), status.ctypes.data_as(up), err.ctypes.data_as(fp)))
        return out, status, err

    def advance(self, image_dev, pitch, pts, ready_event=None):
        """vils_klt_advance: the new frame is already on the device (Frontend.load / Frontend.current); its pyramid becomes the next call's prev."""
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
        n = pts.shape[0]
        out = np.zeros((max(n, 1), 2), np.float32); status = np.zeros(max(n, 1), np.uint8); err = np.zeros(max(n, 1), np.float32)
        up, fp = cabi.c_uint8_p, cabi.c_float_p
        _check(self.L.vils_klt_advance(self.h, C.c_void_p(image_dev), int(pitch), C.c_void_p(ready_event), pts.ctypes.data_as(fp) if n else None, n, out.ctypes.data_as(fp),
                                       status.ctypes.data_as(up), err.ctypes.data_as(fp)))
        return out[:n], status[:n], err[:n]

    def upload(self, prev, nxt, pts):
        prev = np.ascontiguousarray(prev, np.uint8); nxt = np.ascontiguousarray(nxt, np.uint8)
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
        self._n = pts.shape[0]
        up, fp = cabi.c_uint8_p, cabi.c_float_p
        _check(self.L.vils_klt_upload(self.h, prev.ctypes.data_as(up), nxt.ctypes.data_as(up), prev.strides[0], pts.ctypes.data_as(fp), self._n))

    def track_device(self):
        _check(self.L.vils_klt_track_device(self.h))

    def download(self):
        n = self._n
        out = np.zeros((n, 2), np.float32); status = np.zeros(n, np.uint8); err = np.zeros(n, np.float32)
        _check(self.L.vils_klt_download(self.h, out.ctypes.data_as(cabi.c_float_p), status.ctypes.data_as(cabi.c_uint8_p), err.ctypes.data_as(cabi.c_float_p)))
        return out, status, err

    @property
    def last_ms(self):
        ms = C.c_float()
        self.L.vils_klt_last_device_ms(self.h, C.byref(ms))
        return ms.value


def triangulate(start, off, pts, Ps, Rs, tic, ric, init_depth=5.0, device=0):
    start = np.ascontiguousarray(start, np.int32); off = np.ascontiguousarray(off, np.int32); pts = np.ascontiguousarray(pts, np.float64)
    Ps = np.ascontiguousarray(Ps, np.float64); Rs = np.ascontiguousarray(Rs, np.float64); tic = np.ascontiguousarray(tic, np.float64); ric = np.ascontiguousarray(ric, np.float64)
    depth = np.zeros(max(len(start), 1))
    ip = cabi.c_int32_p
    _check(load().vils_triangulate(len(start), start.ctypes.data_as(ip), off.ctypes.data_as(ip), _d(pts), len(Ps), _d(Ps), _d(Rs), _d(tic), _d(ric), init_depth, _d(depth), device))
    return depth[:len(start)]


def lidar_associate(map_xyzi, scan_xyzi, q, t, mode, device=0):
    """vils_lidar_associate: returns (out n x 10, valid n, nn n x 5, device_ms)."""
    m = np.ascontiguousarray(map_xyzi, np.float32).reshape(-1, 4); s = np.ascontiguousarray(scan_xyzi, np.float32).reshape(-1, 4)
    q = np.ascontiguousarray(q, np.float64); t = np.ascontiguousarray(t, np.float64)
    out = np.zeros((max(len(s), 1), 10)); valid = np.zeros(max(len(s), 1), np.uint8); nn = np.zeros((max(len(s), 1), 5), np.int32); ms = C.c_float()
    fp = cabi.c_float_p
    _check(load().vils_lidar_associate(m.ctypes.data_as(fp), len(m), s.ctypes.data_as(fp), len(s), _d(q), _d(t), mode, _d(out),
                                       valid.ctypes.data_as(cabi.c_uint8_p), nn.ctypes.data_as(cabi.c_int32_p), C.byref(ms), device))
    return out[:len(s)], valid[:len(s)].astype(bool), nn[:len(s)], ms.value


def depth_register(cloud_xyzi, T1, T2, feat_xyz, num_bins=360, device=0):
    """vils_depth_register: returns (depth m, device_ms)."""
    c = np.ascontiguousarray(cloud_xyzi, np.float32).reshape(-1, 4); f = np.ascontiguousarray(feat_xyz, np.float32).reshape(-1, 3)
    T1 = np.ascontiguousarray(T1, np.float32).reshape(-1)[:12].copy(); T2 = np.ascontiguousarray(T2, np.float32).reshape(-1)[:12].copy()
    depth = np.zeros(max(len(f), 1), np.float32); ms = C.c_float()
    fp = cabi.c_float_p
    _check(load().vils_depth_register(c.ctypes.data_as(fp), len(c), T1.ctypes.data_as(fp), T2.ctypes.data_as(fp), num_bins, f.ctypes.data_as(fp), len(f),
                                      depth.ctypes.data_as(fp), C.byref(ms), device))
    return depth[:len(f)], ms.value


def vgicp_opts(resolution=0.5, neighbor_search=1, **kw):
    """vils_vgicp_default_opts + the estimator's resolution (estimator.cpp:270)."""
    o = cabi.VilsVgicpOpts()
    load().vils_vgicp_default_opts(C.byref(o))
    o.resolution = resolution; o.neighbor_search = neighbor_search
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def vgicp_covariances(xyzi, device=0):
    """vils_vgicp_covariances: returns (cov n x 3 x 3, nn n x 20)."""
    p = np.ascontiguousarray(xyzi, np.float32).reshape(-1, 4)
    c6 = np.zeros((max(len(p), 1), 6)); nn = np.zeros((max(len(p), 1), 20), np.int32)
    _check(load().vils_vgicp_covariances(p.ctypes.data_as(cabi.c_float_p), len(p), _d(c6), nn.ctypes.data_as(cabi.c_int32_p), device))
    return _sym6(c6[:len(p)]), nn[:len(p)]


def _sym6(c6):
    out = np.zeros((len(c6), 3, 3))
    for e, (r, k) in enumerate(((0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2))):
        out[:, r, k] = c6[:, e]; out[:, k, r] = c6[:, e]
    return out


def vgicp_linearize(src_xyzi, tgt_xyzi, T, opts=None, device=0):
    """vils_vgicp_linearize: returns dict(H 6x6, b 6, error, n_corr, n_voxels, voxels {first point index: (mean, cov 3x3, num)})."""
    s = np.ascontiguousarray(src_xyzi, np.float32).reshape(-1, 4); t = np.ascontiguousarray(tgt_xyzi, np.float32).reshape(-1, 4)
    T = np.ascontiguousarray(T, np.float64).reshape(16); opts = opts or vgicp_opts()
    H = np.zeros((6, 6)); b = np.zeros(6); err = np.zeros(1); nc = np.zeros(1, np.int32); nv = np.zeros(1, np.int32); vox = np.zeros((max(len(t), 1), 10))
    fp = cabi.c_float_p; ip = cabi.c_int32_p
    _check(load().vils_vgicp_linearize(s.ctypes.data_as(fp), len(s), t.ctypes.data_as(fp), len(t), _d(T), C.byref(opts), _d(H), _d(b), _d(err), nc.ctypes.data_as(ip),
                                       nv.ctypes.data_as(ip), _d(vox), device))
    rows = np.nonzero(vox[:, 9] > 0)[0]
    covs = _sym6(vox[rows, 3:9])
    voxels = {int(r): (vox[r, :3].copy(), covs[k], int(vox[r, 9])) for k, r in enumerate(rows)}
    return dict(H=H, b=b, error=float(err[0]), n_corr=int(nc[0]), n_voxels=int(nv[0]), voxels=voxels)


def vgicp_align(src_xyzi, tgt_xyzi, guess=None, opts=None, device=0):
    """vils_vgicp_align (FastVGICP::align + getFitnessScore): returns dict(T 4x4, H 6x6, error, fitness, iterations, converged, ...)."""
    s = np.ascontiguousarray(src_xyzi, np.float32).reshape(-1, 4); t = np.ascontiguousarray(tgt_xyzi, np.float32).reshape(-1, 4)
    opts = opts or vgicp_opts(); res = cabi.VilsVgicpResult()
    g = None if guess is None else np.ascontiguousarray(guess, np.float64).reshape(16)
    fp = cabi.c_float_p
    _check(load().vils_vgicp_align(s.ctypes.data_as(fp), len(s), t.ctypes.data_as(fp), len(t), None if g is None else _d(g), C.byref(opts), C.byref(res), device))
    return dict(T=np.array(res.T).reshape(4, 4), H=np.array(res.H).reshape(6, 6), error=res.error, fitness=res.fitness, iterations=res.iterations,
                converged=bool(res.converged), n_corr=res.n_corr, n_voxels=res.n_voxels, n_linearize=res.n_linearize, elapsed_ms=res.elapsed_ms)


class Frontend:
    """vils_frontend handle: CLAHE, setMask, goodFeaturesToTrack, liftProjective (the rest of FeatureTracker::readImage)."""

    def __init__(self, rows, cols, max_pts=1024, device=0):
        self.L = load(); self.rows, self.cols, self.max_pts = rows, cols, max_pts
        self.h = C.c_void_p()
        _check(self.L.vils_frontend_create(rows, cols, max_pts, device, C.byref(self.h)))

    def close(self):
        if self.h:
            self.L.vils_frontend_destroy(self.h); self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def clahe(self, img, clip=3.0, tiles=(8, 8)):
        img = np.ascontiguousarray(img, np.uint8); out = np.zeros_like(img)
        _check(self.L.vils_clahe(self.h, img.ctypes.data_as(cabi.c_uint8_p), img.strides[0], clip, tiles[0], tiles[1], out.ctypes.data_as(cabi.c_uint8_p), out.strides[0]))
        return out

    def load(self, img, equalize=True, clip=3.0, tiles=(8, 8)):
        """vils_frontend_load: one upload, CLAHE on the device when asked; returns (device pointer, pitch) of the resident frame."""
        img = np.ascontiguousarray(img, np.uint8)
        _check(self.L.vils_frontend_load(self.h, img.ctypes.data_as(cabi.c_uint8_p), img.strides[0], int(equalize), clip, tiles[0], tiles[1]))
        p = C.c_void_p(); pitch = C.c_int32(); ev = C.c_void_p()
        _check(self.L.vils_frontend_current(self.h, C.byref(p), C.byref(pitch), C.byref(ev)))
        self.ready_event = ev.value
        return p.value, pitch.value

    def good_features_resident(self, max_corners, quality, min_distance, use_mask=False):
        out = np.zeros((max(max_corners, 1), 2), np.float32); n = C.c_int32()
        _check(self.L.vils_good_features_resident(self.h, max_corners, quality, min_distance, int(use_mask), out.ctypes.data_as(cabi.c_float_p),
                                                  C.cast(C.byref(n), cabi.c_int32_p)))
        return out[:n.value].copy()

    def set_mask(self, xy, track_cnt, radius):
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2); tc = np.ascontiguousarray(track_cnt, np.int32)
        keep = np.zeros(max(len(xy), 1), np.int32); nk = C.c_int32()
        _check(self.L.vils_set_mask(self.h, xy.ctypes.data_as(cabi.c_float_p), tc.ctypes.data_as(cabi.c_int32_p), len(xy), radius,
                                    keep.ctypes.data_as(cabi.c_int32_p), C.cast(C.byref(nk), cabi.c_int32_p)))
        return keep[:nk.value]

    def get_mask(self):
        m = np.zeros((self.rows, self.cols), np.uint8)
        _check(self.L.vils_get_mask(self.h, m.ctypes.data_as(cabi.c_uint8_p), m.strides[0]))
        return m

    def good_features(self, img, max_corners, quality, min_distance, use_mask=False):
        img = np.ascontiguousarray(img, np.uint8)
        out = np.zeros((max(max_corners, 1), 2), np.float32); n = C.c_int32()
        _check(self.L.vils_good_features(self.h, img.ctypes.data_as(cabi.c_uint8_p), img.strides[0], max_corners, quality, min_distance, int(use_mask),
                                         out.ctypes.data_as(cabi.c_float_p), C.cast(C.byref(n), cabi.c_int32_p)))
        return out[:n.value].copy()

    def eig(self):
        e = np.zeros((self.rows, self.cols), np.float32)
        _check(self.L.vils_frontend_get_eig(self.h, e.ctypes.data_as(cabi.c_float_p)))
        return e

    def lift_projective(self, cam, uv):
        cam = np.ascontiguousarray(cam, np.float64); uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
        rays = np.zeros((len(uv), 3))
        _check(self.L.vils_lift_projective(self.h, _d(cam), uv.ctypes.data_as(cabi.c_float_p), len(uv), _d(rays)))
        return rays

    def reject_with_f(self, p1, p2, threshold=1.0):
        p1 = np.ascontiguousarray(p1, np.float32).reshape(-1, 2); p2 = np.ascontiguousarray(p2, np.float32).reshape(-1, 2)
        st = np.zeros(max(len(p1), 1), np.uint8); F = np.zeros(9)
        _check(self.L.vils_reject_with_f(self.h, p1.ctypes.data_as(cabi.c_float_p), p2.ctypes.data_as(cabi.c_float_p), len(p1), threshold,
                                         st.ctypes.data_as(cabi.c_uint8_p), _d(F)))
        return st[:len(p1)].astype(bool), F.reshape(3, 3)

    @property
    def last_ms(self):
        ms = C.c_float(); self.L.vils_frontend_last_device_ms(self.h, C.byref(ms)); return ms.value
