"""Window sharding across ranks: independent windows, no data-path collective; only the timing
reduction (max over ranks) and the final gather of solved states cross ranks."""


def shard_range(total, world, rank):
    """Contiguous, balanced [lo, hi) of `total` windows for `rank` (first `total % world` ranks get one more)."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(values, device=None):
    """Element-wise max of a list of floats over all ranks (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return list(values)
    t = torch.tensor(list(values), dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def gather_states(local_states, device=None):
    """all_gather of per-rank (n_local, X) solved-state arrays -> list over ranks (numpy)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [np.asarray(local_states)]
    t = torch.as_tensor(np.ascontiguousarray(local_states), dtype=torch.float64, device=device or "cpu")
    sizes = [torch.zeros(1, dtype=torch.int64, device=t.device) for _ in range(dist.get_world_size())]
    dist.all_gather(sizes, torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device))
    mx = int(max(int(s.item()) for s in sizes))
    pad = torch.zeros((mx, t.shape[1]), dtype=torch.float64, device=t.device); pad[:t.shape[0]] = t
    outs = [torch.zeros_like(pad) for _ in sizes]
    dist.all_gather(outs, pad)
    return [o[:int(s.item())].cpu().numpy() for o, s in zip(outs, sizes)]


def split_window(w, world, rank):
    """Factor shard of ONE window for `rank`: the full state, landmarks (with all their projection
    factors) by feature % world, LiDAR plane/edge factors round-robin, IMU / prior / ICP / LPS on rank 0 only."""
    import numpy as np
    out = dict(w)
    if w.get("kf_i") is not None and len(w["kf_i"]):
        keep = (np.asarray(w["feat"]) % world) == rank
        for k in ["pts_i", "pts_j", "vel_i", "vel_j", "td_i", "td_j", "row_i", "row_j", "kf_i", "kf_j", "feat"]:
            out[k] = np.asarray(w[k])[keep]
    for fam, keys in (("plane_kf", ["plane_p", "plane_n", "plane_d", "plane_kf"]), ("edge_kf", ["edge_p", "edge_a", "edge_b", "edge_kf"])):
        if w.get(fam) is not None and len(w[fam]):
            keep = (np.arange(len(w[fam])) % world) == rank
            for k in keys:
                out[k] = np.asarray(w[k])[keep]
    if rank != 0:
        out["imu"] = np.asarray(w["imu"])[:0]; out["imu_kf"] = np.asarray(w["imu_kf"])[:0]
        out["icp"], out["lps"] = [], []
        out["prior_n"] = 0
        for k in ["prior_J", "prior_r", "prior_blk", "prior_x0"]:
            out.pop(k, None)
    return out


def native_sharded_solve_benchmark(lib, cfg, w, opts, rank, world, reps=5):
    """Factor-sharded solve of ONE window over `world` GPUs through vils_ba_sharded_solve (library-owned NCCL communicator, all
    Gauss-Newton iterations enqueued on the library stream with the all-reduce between the linearise and update kernels, one host
    synchronisation).  torch.distributed must be initialised (only used to hand the NCCL unique id around and for the max over ranks).
    Returns (record dict on every rank, solved state of this rank)."""
    import time
    import numpy as np
    import torch
    import torch.distributed as dist
    uid = [lib.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    h = lib.BA(cfg, 1)
    h.sharded_init(rank, world, uid[0])
    h.set_window(0, split_window(w, world, rank)); h.upload(1)
    dev_ms, wall_ms = [], []
    for _ in range(reps + 2):
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        s = h.sharded_solve(opts)
        wall_ms.append(1e3 * (time.perf_counter() - t0)); dev_ms.append(h.last_ms)
    dev = max_over_ranks([float(np.mean(dev_ms[2:]))], device="cuda")[0]
    wall = max_over_ranks([float(np.mean(wall_ms[2:]))], device="cuda")[0]
    st = h.get_state(0)
    D = 15 * w["pose"].shape[0] + 7
    rec = {"n_gpus": world, "D": D, "allreduce_bytes_per_iteration": (D * D + 2 * D + 1) * 8, "iterations": int(opts.max_iters),
           "device_ms_per_solve": dev, "wall_ms_per_solve": wall, "status": int(s.status), "launches_per_solve": h.last_launches}
    h.close()
    return rec, st
