#!/usr/bin/env python
"""bench.py — sliding-window BA solves/sec (BASELINE.json metric) on N B200s of one node.

Workload (configs[1]): 10-KF window, 150 features (906 projection factors), 2000 LiDAR plane/edge factors, 9 IMU factors,
a 7-dim prior on [extrinsic, td]; 5 Gauss-Newton iterations per solve, FP64.  One "step" = one batched solve of
`--windows` independent windows per GPU (weak scaling: windows shard across ranks with no data-path collective).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--windows B] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (inputs already in HBM), `e2e` = the same metric
through vils_ba_solve() with host buffers (pinned staging -> H2D -> solve -> D2H inside the timed region).
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
# dram__bytes_read.sum + dram__bytes_write.sum of solve_kernel per window, from the committed ncu capture (profiles/r1_ncu_solve_kernel.txt)
SOLVE_DRAM_TRAFFIC_PER_WINDOW = 5265906   # (335.82 + 443.53) MB / 148 windows, ncu --set full capture of solve_kernel (profiles/r1_ncu_solve_s3.txt)
# Materialised Evaluate() traffic per window (what the CPU reference moves per linearisation; §8d first table)
BYTES_PER_EVAL_WINDOW = 906 * 460 + 9 * 6016 + 1500 * 116 + 500 * 244 + 2544


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--windows", type=int, default=592, help="windows per GPU per step (4 x 148 SMs; 592 x 281 KB = 166 MB > 126 MB L2)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
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
    t0 = time.perf_counter()
    for k in range(B):
        ba.set_window(k, uniq[k % UNIQUE_WINDOWS])
    pack_ms = 1e3 * (time.perf_counter() - t0) / B
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
    # ---- end to end through the C-ABI with host buffers
    for _ in range(max(1, args.warmup // 2)):
        ba.solve(B, opts)
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        ba.solve(B, opts)                 # pinned staging -> chunked H2D | solve kernels | D2H of states + summaries, pipelined on 3 streams
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - e0)
    e2e_launches = ba.last_launches
    # ---- Jacobian evaluation kernel (materialised Evaluate of every factor), the HBM-roofline kernel north_star names
    for _ in range(3):
        ba.evaluate_device(B, True)
    ev_ms = 0.0
    for _ in range(args.steps):
        ba.evaluate_device(B, True); ev_ms += ba.last_ms
    sampler.stop()
    st = ba.get_state(0)
    assert st["status"] == 0, st
    from mvil_fusion_b200.sharding import max_over_ranks
    dev_ms, wall_ms, e2e_ms, ev_ms = max_over_ranks([dev_ms, wall_ms, e2e_ms, ev_ms], device="cuda")   # timing = max over ranks
    if rank == 0:
        peak, peak_src = read_peaks()
        total = B * args.steps * world
        value = total / (dev_ms * 1e-3)
        kern_ms = dev_ms / args.steps
        achieved = BYTES_PER_SOLVE * B / (kern_ms * 1e-3) / 1e9
        ev_achieved = BYTES_PER_EVAL_WINDOW * B / (ev_ms / args.steps * 1e-3) / 1e9
        h2d, d2h = ba.last_transfer_bytes      # counted by the library from the copies it issued in the last vils_ba_solve
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
                       "solver": "GN x5, mu=1e-8 Jacobi damping, Cauchy(1) visual, Huber(0.1) LiDAR", "host_pack_ms_per_window": pack_ms,
                       "wall_ms_per_step": wall_ms / args.steps},
            "e2e": {"value": total / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
                    "launches_per_step": e2e_launches, "pipeline": "chunks of n_sm/2 windows round-robin on 3 streams: H2D | solve_kernel (prep folded in) | D2H"},
            "gpu_launches": args.steps,
            "roofline": {"kernel": "solve_kernel (fused GN loop, one CTA per window)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": SOLVE_DRAM_TRAFFIC_PER_WINDOW * B if SOLVE_DRAM_TRAFFIC_PER_WINDOW else None, "peak_source": peak_src,
                         "algorithmic_bytes_per_solve": BYTES_PER_SOLVE, "note": "latency-bound FP64 kernel; see DESIGN.md §4 and profiles/"},
            "roofline_eval": {"kernel": "eval_imu + eval_proj + eval_lidar + eval_prior kernels (materialised residual+Jacobian of every factor)", "bound": "hbm", "achieved": ev_achieved, "peak": peak,
                              "unit": "GB/s", "frac": ev_achieved / peak, "ms_per_launch": ev_ms / args.steps, "algorithmic_bytes_per_window": BYTES_PER_EVAL_WINDOW},
            "cpu_baseline": {"value": tot / T, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{tot} GN-5 solves of the same config-2 windows by oracle/ (CPU restatement) on {cores} threads",
                             "single_thread_value": v_single},
            "clocks": sampler.summary(),
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
