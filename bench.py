#!/usr/bin/env python
"""bench.py — sliding-window BA solves/sec (BASELINE.json metric) on N B200s of one node.

Workload (configs[1]): 10-KF window, 150 features (906 projection factors), 2000 LiDAR plane/edge factors, 9 IMU factors,
a 7-dim prior on [extrinsic, td]; 5 Gauss-Newton iterations per solve, FP64.  One "step" = one batched solve of
`--windows` independent windows per GPU (weak scaling: windows shard across ranks with no data-path collective).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--windows B] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (inputs already in HBM), `e2e` = the same metric
through vils_ba_solve_windows() from the CALLER's vils_window arrays to solved states: host pack (all host threads) -> pinned
staging -> H2D -> solve -> D2H, all inside the timed region.  `configs` carries the other single-GPU BASELINE configs
(configs[2] KLT, configs[3] 20-KF window with the real marginalization prior) measured in the same run.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from mvil_fusion_b200 import cabi, synth  # noqa: E402

METRIC = "sliding-window BA solves/sec (10 KF, 150 feat, 2k LiDAR pts)"
UNIT = "solves/s"
WORKLOAD = "configs[1]: 10-KF window, 150 feats (906 proj factors) + 2000 LiDAR edge/plane factors, 5 GN iters, FP64"
UNIQUE_WINDOWS = 74   # distinct seeded windows generated on the host; tiled to fill the batch (distinct HBM slots)

# Algorithmic bytes (restated in DESIGN.md §4): fused lower bound per linearisation of one config-2 window
#   reads  906*124 + 9*2296 + 1500*60 + 500*76 + 2.5 K state = 265,572 B ; writes (D^2 + D)*8 = 198,448 B  (D = 157)
BYTES_PER_LINEARISATION = 906 * 124 + 9 * 2296 + 1500 * 60 + 500 * 76 + 2544 + (157 * 157 + 157) * 8
BYTES_PER_COST_EVAL = 906 * 124 + 9 * 2296 + 1500 * 60 + 500 * 76 + 2544
BYTES_PER_SOLVE = 5 * BYTES_PER_LINEARISATION + BYTES_PER_COST_EVAL
# dram__bytes_read.sum + dram__bytes_write.sum of solve_kernel per window, from the committed ncu capture
SOLVE_DRAM_TRAFFIC_PER_WINDOW = 3273324   # (1026.30 + 911.50) MB / 592 windows, ncu --set full capture of solve_kernel (profiles/r2_ncu_solve_v2.txt)
# Materialised Evaluate() traffic per window (what the CPU reference moves per linearisation; §8d first table)
BYTES_PER_EVAL_WINDOW = 906 * 460 + 9 * 6016 + 1500 * 116 + 500 * 244 + 2544
# Algorithmic FP64 work per linearisation (Schur SYRK 2 D^2 M = 7.4 MFLOP, Cholesky D^3/3 = 1.3 MFLOP at D = 157, M = 150;
# factor evaluation + J^T J: ~1500 flop per projection factor, ~200 per LiDAR factor, ~34 k per IMU factor), 5 linearisations per solve.
FLOPS_PER_SOLVE = 5 * (2 * 157 * 157 * 150 + 157 ** 3 // 3 + 906 * 1500 + 2000 * 200 + 9 * 34000)
FP64_PEAK_TFLOPS = 36.0      # measured DFMA peak of this part (profiles/r1_fp64_microbench.txt); B200 has no FP64 tensor-core path worth the name
# configs[3] (N = 20, M = 300, 3333 projection + 3750 plane + 1250 edge factors, 3 ICP + 3 LPS, n = 136 prior), same accounting, D = 307
CFG4_PRIOR_N = 136
CFG4_READS = 3333 * 124 + 19 * 2296 + 3750 * 60 + 1250 * 76 + (CFG4_PRIOR_N * CFG4_PRIOR_N + 2 * CFG4_PRIOR_N) * 8 + 6 * 200 + (16 * 20 + 8 + 300) * 8
CFG4_BYTES_PER_SOLVE = 5 * (CFG4_READS + (307 * 307 + 307) * 8) + CFG4_READS
KLT_BYTES_PER_FRAME_PAIR = 816000   # both 4-level u8 pyramids read once


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region: NVML in-process every 10 ms (the whole default
    run lasts well under a second, so an `nvidia-smi -lms` child process would deliver its first line too late)."""

    REASONS = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
               ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"))

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.mx, self.reasons, self.stop_flag, self.src = index, [], None, set(), False, None

    def _visible_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self._visible_index())
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.src = "nvml"
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop_flag:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = get_reasons(h)
                for name, attr in self.REASONS:
                    if r & getattr(nv, attr, 0):
                        self.reasons.add(name)
                time.sleep(0.01)
            return
        except Exception:
            pass
        # fallback: one-shot nvidia-smi queries in a loop
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        self.src = "nvidia-smi"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self._visible_index()}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.sm.append(float(f[0])); self.mx = float(f[1])
                for n, v in zip([r[0] for r in self.REASONS], f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                time.sleep(0.05)

    def stop(self):
        self.stop_flag = True
        self.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "source": self.src}


def oracle_lib():
    """CPU oracle — used ONLY for the cpu_baseline leg and --impl reference (see oracle/README.md)."""
    so = os.path.join(ROOT, "oracle", "_build", "libvils_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return ctypes.CDLL(so)


def cpu_solves_per_sec(windows, cores, per_core, opts):
    """GN-5 solves of the same windows by the CPU restatement on `cores` host threads (ctypes releases the GIL)."""
    L = oracle_lib()
    cfg = cabi.default_config()
    structs = [cabi.window_struct(w) for w in windows]
    done = [0] * cores

    def work(t):
        N, M = 10, 150
        pose = np.zeros((N, 7)); sb = np.zeros((N, 9)); ex = np.zeros(7); lam = np.zeros(M); td = ctypes.c_double(); s = cabi.VilsSummary()
        dp = lambda a: a.ctypes.data_as(cabi.c_double_p)
        for k in range(per_core):
            ws, _ = structs[(t * per_core + k) % len(structs)]
            L.vo_solve_window(ctypes.byref(cfg), ctypes.byref(ws), ctypes.byref(opts), dp(pose), dp(sb), dp(ex), dp(lam), ctypes.byref(td), ctypes.byref(s))
            done[t] += 1

    th = [threading.Thread(target=work, args=(t,)) for t in range(cores)]
    t0 = time.perf_counter()
    for x in th:
        x.start()
    for x in th:
        x.join()
    dt = time.perf_counter() - t0
    return sum(done) / dt, dt


def cpu_solves_generic(windows, cfg, opts, cores, total):
    """`total` solves of arbitrary windows by the CPU restatement on `cores` threads -> solves/s."""
    L = oracle_lib()
    structs = [cabi.window_struct(w) for w in windows]
    per = max(1, total // cores)

    def work(t):
        ws0, _ = structs[0]
        N, M = ws0.n_kf, ws0.n_feat
        pose = np.zeros((N, 7)); sb = np.zeros((N, 9)); ex = np.zeros(7); lam = np.zeros(max(M, 1)); td = ctypes.c_double(); s = cabi.VilsSummary()
        dp = lambda a: a.ctypes.data_as(cabi.c_double_p)
        for k in range(per):
            ws, _ = structs[(t * per + k) % len(structs)]
            L.vo_solve_window(ctypes.byref(cfg), ctypes.byref(ws), ctypes.byref(opts), dp(pose), dp(sb), dp(ex), dp(lam), ctypes.byref(td), ctypes.byref(s))

    th = [threading.Thread(target=work, args=(t,)) for t in range(cores)]
    t0 = time.perf_counter()
    for x in th:
        x.start()
    for x in th:
        x.join()
    return per * cores / (time.perf_counter() - t0)


def build_config4_windows(lib, cfg4, n, opts):
    """configs[3] windows with the REAL marginalization prior, produced by the product path only (no oracle): solve the 21-frame
    window -> double2vector -> put_state -> vils_ba_marginalize(MARGIN_OLD) -> slide -> attach the prior."""
    out = []
    h = lib.BA(cfg4, 1)
    for k in range(n):
        big = synth.make_config4_big(k)
        h.set_window(0, big); h.solve(1, opts)
        s = h.get_state(0)
        assert s["status"] == 0, s
        pose, sb = lib.double2vector(big["pose"][0], s["pose"], s["speedbias"])
        h.put_state(0, pose, sb, s["ex_pose"], s["inv_depth"], s["td"])
        prior = h.marginalize(0, cabi.VILS_MARGIN_OLD)
        solved = dict(big); solved.update(pose=pose, speedbias=sb, ex_pose=s["ex_pose"], inv_depth=s["inv_depth"], td=s["td"])
        w = synth.attach_prior(synth.slide_old(solved), prior)
        w.pop("truth", None); w.pop("raw", None)
        out.append(w)
    h.close()
    return out


def other_configs(lib, device, cores, opts):
    """configs[2] (KLT) and configs[3] (20-KF window with the real prior) on this GPU: device ms, e2e ms, roofline fraction, CPU figure."""
    peak, _ = read_peaks()
    rec = {}
    # ---- configs[3]
    try:
        cfg4 = cabi.default_config(max_kf=21, max_feat=320, max_proj=4000, max_lidar=5000, device=device)
        ws4 = build_config4_windows(lib, cfg4, 4, opts)
        B4 = 148
        h = lib.BA(cfg4, B4)
        batch = [ws4[k % len(ws4)] for k in range(B4)]
        arr, keep = lib.BA.window_array(batch)
        h.set_windows(0, batch, arr); h.upload(B4)
        ms = []
        for _ in range(6):
            h.solve_device(B4, opts); ms.append(h.last_ms)
        dev = float(np.mean(ms[2:]))
        h.set_cluster(1); h.solve_device(1, opts); h.solve_device(1, opts); single_one = h.last_ms
        h.set_cluster(0); h.solve_device(1, opts); h.solve_device(1, opts); single = h.last_ms; single_cluster = h.last_cluster
        for _ in range(2):
            h.solve_windows(batch, opts, arr)
        t0 = time.perf_counter()
        for _ in range(4):
            h.solve_windows(batch, opts, arr)
        e2e = 1e3 * (time.perf_counter() - t0) / 4
        t0 = time.perf_counter(); h.solve_windows(batch[:1], opts, arr); e2e_single = 1e3 * (time.perf_counter() - t0)
        st = h.get_state(0)
        h.close()
        cpu = cpu_solves_generic(ws4, cfg4, opts, cores, 2 * cores)
        cpu1 = cpu_solves_generic(ws4, cfg4, opts, 1, 2)
        ach = CFG4_BYTES_PER_SOLVE * B4 / (dev * 1e-3) / 1e9
        rec["config4"] = {"workload": "configs[3]: 20-KF window, 300 feats (3333 proj factors), 5000 LiDAR factors, 3 ICP + 3 LPS, real marginalization prior n=%d, GN x5" % int(ws4[0]["prior_n"]),
                          "windows": B4, "device_ms": dev, "solves_per_s_device": B4 / dev * 1e3, "single_window_device_ms": single, "single_window_cluster_size": single_cluster,
                          "single_window_device_ms_one_cta": single_one,
                          "e2e_ms_from_caller_arrays": e2e, "solves_per_s_e2e": B4 / e2e * 1e3, "single_window_e2e_ms": e2e_single,
                          "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "algorithmic_bytes_per_solve": CFG4_BYTES_PER_SOLVE},
                          "cpu_baseline": {"value": cpu, "unit": UNIT, "cores": cores, "kind": "port", "single_thread_value": cpu1},
                          "status": int(st["status"]), "cost_initial": st["cost_initial"], "cost_final": st["cost_final"]}
    except Exception as e:   # a sub-record must never take the headline line down
        rec["config4"] = {"error": repr(e)}
    # ---- configs[2]: pyramidal KLT 640x480, 150 corners, 3 pyramid levels above the base
    try:
        rng = np.random.default_rng(1)
        try:
            import cv2
        except Exception:
            cv2 = None
        f = lib.Frontend(480, 640, 512, device=device)
        if cv2 is not None:
            img = cv2.GaussianBlur(rng.uniform(0, 255, (480, 640)).astype(np.float32), (0, 0), 2.0)
            img = cv2.normalize(img, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)
        else:
            base = rng.uniform(0, 255, (480 // 4, 640 // 4)); img = np.kron(base, np.ones((4, 4))).astype(np.uint8)
        img = f.clahe(img)                                    # readImage applies CLAHE before LK (feature_tracker.cpp:87-93)
        nxt = np.roll(np.roll(img, 3, axis=0), -4, axis=1)
        if cv2 is not None:
            M = np.float64([[1, 0, -4.3], [0, 1, 3.2]])
            nxt = cv2.warpAffine(img, M, (640, 480), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
        pts = f.good_features(img, 150, 0.01, 30.0)
        k = lib.KLT(480, 640, 512, 21, 3, device=device)
        k.upload(img, nxt, pts)
        ms = []
        for _ in range(30):
            k.track_device(); ms.append(k.last_ms)
        dev = float(np.mean(ms[5:]))
        for _ in range(3):
            k.track(img, nxt, pts)
        t0 = time.perf_counter()
        for _ in range(20):
            out, stt, err = k.track(img, nxt, pts)
        e2e = 1e3 * (time.perf_counter() - t0) / 20
        # the readImage chain: raw frame uploaded once, CLAHE + ONE pyramid + LK on the device, previous pyramid reused (vils_frontend_load + vils_klt_advance)
        raw0, raw1 = (np.roll(img, 0, axis=0), nxt)
        k2 = lib.KLT(480, 640, 512, 21, 3, device=device)
        d0, pitch = f.load(raw0, True); k2.advance(d0, pitch, np.zeros((0, 2), np.float32), f.ready_event)
        for _ in range(3):
            d1, pitch = f.load(raw1, True); k2.advance(d1, pitch, pts, f.ready_event)
        t0 = time.perf_counter()
        for i in range(20):
            d1, pitch = f.load(raw1 if i % 2 == 0 else raw0, True); k2.advance(d1, pitch, pts, f.ready_event)
        chain = 1e3 * (time.perf_counter() - t0) / 20
        k2.close()
        r = {"workload": "configs[2]: pyramidal LK 640x480 mono, %d corners, win 21, 3 pyramid levels above the base" % len(pts), "device_ms": dev,
             "e2e_ms_host_buffers": e2e, "frame_pairs_per_s_device": 1e3 / dev, "frame_pairs_per_s_e2e": 1e3 / e2e, "tracked": int(stt.sum()),
             "read_image_chain_ms_per_frame": chain, "read_image_chain": "one 307 KB upload + CLAHE + one pyramid + LK per frame, previous pyramid resident (vils_frontend_load + vils_klt_advance)",
             "roofline": {"bound": "hbm", "achieved": KLT_BYTES_PER_FRAME_PAIR / dev / 1e6, "peak": peak, "unit": "GB/s", "frac": KLT_BYTES_PER_FRAME_PAIR / dev / 1e6 / peak,
                          "algorithmic_bytes_per_frame_pair": KLT_BYTES_PER_FRAME_PAIR, "note": "one 0.8 MB frame pair cannot load HBM: launch/latency-bound"}}
        if cv2 is not None:
            def best(fn, n=10):
                b = 1e9
                for _ in range(n):
                    t = time.perf_counter(); fn(); b = min(b, time.perf_counter() - t)
                return b * 1e3
            p3 = pts.reshape(-1, 1, 2)
            cv2.setNumThreads(1); c1 = best(lambda: cv2.calcOpticalFlowPyrLK(img, nxt, p3, None, winSize=(21, 21), maxLevel=3))
            cv2.setNumThreads(0); cn = best(lambda: cv2.calcOpticalFlowPyrLK(img, nxt, p3, None, winSize=(21, 21), maxLevel=3))
            r["cpu_baseline"] = {"kind": "reference (cv2 %s, the library the reference calls)" % cv2.__version__, "ms_1_thread": c1, "ms_all_threads": cn, "cores": cores}
        k.close(); f.close()
        rec["klt"] = r
    except Exception as e:
        rec["klt"] = {"error": repr(e)}
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--windows", type=int, default=592, help="windows per GPU per step (4 x 148 SMs; 592 x 281 KB = 166 MB > 126 MB L2)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-configs", action="store_true", help="skip the configs[2] (KLT) and configs[3] (20-KF window) sub-records")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    opts = cabi.default_solve_opts(cabi.VILS_MODE_GN, 5, 1e-8)
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        # The reference (ROS + Ceres + Eigen C++) cannot be built in this image, so the reference arm is the CPU restatement
        # of its arithmetic (oracle/, "port"), on all host cores, same windows / metric / GN-5 schedule.
        if rank != 0:
            return
        wins = [synth.make_window(2, k) for k in range(min(16, cores * 2))]
        per_core = 2
        vals = []
        for s in range(args.warmup + args.steps):
            v, dt = cpu_solves_per_sec(wins, cores, per_core, opts)
            if s >= args.warmup:
                vals.append((v, dt))
        total = sum(cores * per_core for _ in vals); T = sum(dt for _, dt in vals)
        value = total / T
        print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 * T / max(len(vals), 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                          "data": "synthetic", "impl": "reference", "config": {"workload": WORKLOAD},
                          "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": f"{cores * per_core} windows per step, GN-5, {cores} threads"},
                          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return

    import torch
    import torch.distributed as dist
    from mvil_fusion_b200 import lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libvils_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.windows
    cfg = cabi.default_config(device=local)
    uniq = [synth.make_window(2, (rank * UNIQUE_WINDOWS + k)) for k in range(UNIQUE_WINDOWS)]
    ba = lib.BA(cfg, B)
    batch = [uniq[k % UNIQUE_WINDOWS] for k in range(B)]
    warr, wkeep = lib.BA.window_array(batch)          # the CALLER's arrays: B vils_window structs over host numpy buffers
    # C-side pack alone (vils_ba_set_windows, no per-window ctypes): all host threads, and one thread on a second handle
    ba.set_windows(0, batch, warr)
    t0 = time.perf_counter()
    for _ in range(3):
        ba.set_windows(0, batch, warr)
    pack_ms = 1e3 * (time.perf_counter() - t0) / (3 * B)
    os.environ["VILS_PACK_THREADS"] = "1"
    ba1 = lib.BA(cfg, 64)
    ba1.set_windows(0, batch[:64], warr)
    t0 = time.perf_counter(); ba1.set_windows(0, batch[:64], warr); pack1_ms = 1e3 * (time.perf_counter() - t0) / 64
    del os.environ["VILS_PACK_THREADS"]
    ba1.close()
    ba.upload(B)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local); sampler.start()
    # ---- device-resident throughput
    for _ in range(args.warmup):
        ba.solve_device(B, opts)
    barrier()
    w0 = time.perf_counter(); dev_ms = 0.0
    for _ in range(args.steps):
        ba.solve_device(B, opts)          # CUDA events on the launching stream bracket the kernel (vils_ba_last_device_ms)
        dev_ms += ba.last_ms
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - w0)
    # ---- end to end through the C-ABI: caller arrays -> host pack -> H2D -> solve -> D2H -> states in host memory
    for _ in range(max(1, args.warmup // 2)):
        ba.solve_windows(batch, opts, warr)
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        ba.solve_windows(batch, opts, warr)   # pack (host threads) | chunked H2D | solve kernels | D2H, pipelined; one synchronisation
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - e0)
    e2e_launches = ba.last_launches
    h2d, d2h = ba.last_transfer_bytes          # counted by the library from the copies it issued in that vils_ba_solve_windows call (all B windows)
    # the same with the blobs already packed in pinned staging (round-1 definition, kept for comparison)
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        ba.solve(B, opts)
    barrier()
    e2e_packed_ms = 1e3 * (time.perf_counter() - e0)
    # ---- Jacobian evaluation kernel (materialised Evaluate of every factor), the HBM-roofline kernel north_star names
    for _ in range(3):
        ba.evaluate_device(B, True)
    ev_ms = 0.0
    for _ in range(args.steps):
        ba.evaluate_device(B, True); ev_ms += ba.last_ms
    # ---- single-window latency (the drop-in case: one optimization() per frame): cluster kernel (latency mode) vs one CTA
    lat = {}
    for mode, name in ((1, "one_cta"), (0, "cluster")):
        ba.set_cluster(mode)
        ms = []
        for _ in range(8):
            ba.solve_device(1, opts); ms.append(ba.last_ms)
        lat[name + "_device_ms"] = float(np.mean(ms[3:])); lat[name + "_size"] = ba.last_cluster
        for _ in range(3):
            ba.solve_windows(batch[:1], opts, warr)
        t0 = time.perf_counter()
        for _ in range(10):
            ba.solve_windows(batch[:1], opts, warr)
        lat[name + "_e2e_ms_from_caller_arrays"] = 1e3 * (time.perf_counter() - t0) / 10
    ba.set_cluster(0)
    # what the reference's own solver options give (estimator.cpp:1400-1411: DOGLEG, max_num_iterations = NUM_ITERATIONS = 8): on one CTA, and with the
    # trust-region loop on CTA 0 of a cluster whose CTAs all serve the linearisations
    dl = cabi.default_solve_opts(cabi.VILS_MODE_DOGLEG, 8, 0.0)
    for mode, name in ((1, "dogleg8_one_cta"), (0, "dogleg8_cluster")):
        ba.set_cluster(mode)
        ms = []
        for _ in range(6):
            ba.solve_device(1, dl); ms.append(ba.last_ms)
        lat[name + "_device_ms"] = float(np.mean(ms[2:])); lat[name + "_size"] = ba.last_cluster
    ba.set_cluster(0)
    ba.download(1); lat["dogleg8_iterations"] = int(ba.get_state(0)["iterations"])
    sampler.stop()
    st = ba.get_state(0)
    assert st["status"] == 0, st
    extra = other_configs(lib, local, cores, opts) if (rank == 0 and not args.no_configs) else {}
    sharded = None
    if world > 1 and not args.no_configs:
        # configs[4] names "NCCL Hessian allreduce": ONE config-4 window factor-sharded over all ranks by vils_ba_sharded_solve
        # (library-owned communicator, linearise -> ncclAllReduce -> update on the library stream), next to the single-GPU solve of it
        try:
            from mvil_fusion_b200.sharding import native_sharded_solve_benchmark
            cfg4 = cabi.default_config(max_kf=21, max_feat=320, max_proj=4000, max_lidar=5000, device=local)
            w4 = build_config4_windows(lib, cfg4, 1, opts)[0]
            sharded, _ = native_sharded_solve_benchmark(lib, cfg4, w4, opts, rank, world)
            f4 = lib.BA(cfg4, 1); f4.set_window(0, w4); f4.upload(1)
            for _ in range(3):
                f4.solve_device(1, opts)
            sharded["single_gpu_device_ms"] = f4.last_ms; f4.close()
            sharded["workload"] = "one configs[3] window (D = 307) factor-sharded over %d GPUs, one ncclAllReduce of the partial reduced system per GN iteration" % world
        except Exception as e:
            sharded = {"error": repr(e)}
    from mvil_fusion_b200.sharding import max_over_ranks
    dev_ms, wall_ms, e2e_ms, ev_ms, e2e_packed_ms = max_over_ranks([dev_ms, wall_ms, e2e_ms, ev_ms, e2e_packed_ms], device="cuda")   # timing = max over ranks
    if rank == 0:
        peak, peak_src = read_peaks()
        total = B * args.steps * world
        value = total / (dev_ms * 1e-3)
        kern_ms = dev_ms / args.steps
        achieved = BYTES_PER_SOLVE * B / (kern_ms * 1e-3) / 1e9
        ev_achieved = BYTES_PER_EVAL_WINDOW * B / (ev_ms / args.steps * 1e-3) / 1e9
        # CPU baseline: the oracle ("port") on the host cores of this box, bounded sample of the same windows
        per_core = 2
        v1, dt1 = cpu_solves_per_sec(uniq[:16], cores, per_core, opts)
        reps = max(1, int(args.cpu_seconds / max(dt1, 1e-3)) - 1)
        tot, T = cores * per_core, dt1
        for _ in range(min(reps, 8)):
            v, dt = cpu_solves_per_sec(uniq[:16], cores, per_core, opts); tot += cores * per_core; T += dt
        # the reference itself never sets options.num_threads (estimator.cpp:1400-1411): one thread is what a single optimization() gets
        v_single, _ = cpu_solves_per_sec(uniq[:4], 1, 4, opts)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": kern_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "windows_per_gpu_per_step": B, "unique_seeded_windows_per_gpu": UNIQUE_WINDOWS,
                       "l2_policy": "inputs larger than L2 (592 window blobs x 281 KB = 166 MB per GPU, each step re-reads all of them)",
                       "solver": "GN x5, mu=1e-8 Jacobi damping, Cauchy(1) visual, Huber(0.1) LiDAR",
                       "host_pack_ms_per_window_all_threads": pack_ms, "host_pack_ms_per_window_one_thread": pack1_ms, "host_threads": cores,
                       "wall_ms_per_step": wall_ms / args.steps,
                       "e2e_definition": "vils_ba_solve_windows: caller vils_window arrays -> pack on host threads -> pinned staging -> H2D -> solve -> D2H -> states"},
            "e2e": {"value": total / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
                    "launches_per_step": e2e_launches, "from": "caller arrays (host pack inside the timed region)",
                    "prepacked_value": total / (e2e_packed_ms * 1e-3),
                    "pipeline": "chunks of n_sm/8, n_sm/4, n_sm/2, rest (repeating): host pack (thread pool) | H2D on a copy stream | solve_kernel (prep folded in) | D2H on round-robin streams"},
            "gpu_launches": args.steps,
            "roofline": {"kernel": "solve_kernel (fused GN loop, one CTA per window)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": SOLVE_DRAM_TRAFFIC_PER_WINDOW * B if SOLVE_DRAM_TRAFFIC_PER_WINDOW else None, "peak_source": peak_src,
                         "algorithmic_bytes_per_solve": BYTES_PER_SOLVE, "note": "latency-bound FP64 kernel; see DESIGN.md §4 and profiles/"},
            "roofline_fp64": {"kernel": "solve_kernel", "bound": "fp64", "achieved": FLOPS_PER_SOLVE * B / (kern_ms * 1e-3) / 1e12, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                              "frac": FLOPS_PER_SOLVE * B / (kern_ms * 1e-3) / 1e12 / FP64_PEAK_TFLOPS, "flops_per_solve": FLOPS_PER_SOLVE,
                              "peak_source": "measured DFMA peak, profiles/r1_fp64_microbench.txt"},
            "roofline_eval": {"kernel": "eval_imu + eval_proj + eval_lidar + eval_prior kernels (materialised residual+Jacobian of every factor)", "bound": "hbm", "achieved": ev_achieved, "peak": peak,
                              "unit": "GB/s", "frac": ev_achieved / peak, "ms_per_launch": ev_ms / args.steps, "algorithmic_bytes_per_window": BYTES_PER_EVAL_WINDOW},
            "cpu_baseline": {"value": tot / T, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{tot} GN-5 solves of the same config-2 windows by oracle/ (CPU restatement) on {cores} threads",
                             "single_thread_value": v_single},
            "latency": dict(lat, workload="ONE configs[1] window per call, GN x5; cluster = thread-block cluster per window (vils_ba_set_cluster); dogleg8 = the reference's ceres options (DOGLEG, max 8 iterations), on one CTA and on a cluster"),
            "clocks": sampler.summary(),
            "configs": extra,
        }
        if sharded is not None:
            out["sharded"] = sharded
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
