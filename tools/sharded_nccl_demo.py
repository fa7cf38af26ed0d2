"""torchrun --nproc-per-node G tools/sharded_nccl_demo.py — one window factor-sharded over G GPUs with an NCCL all-reduce of
the partial reduced system per Gauss-Newton iteration; rank 0 checks against its own single-GPU solve of the whole window."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import torch.distributed as dist
from mvil_fusion_b200 import cabi, synth, lib
from mvil_fusion_b200.sharding import split_window

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
big = len(sys.argv) > 1 and sys.argv[1] == "big"
N, M, NL = (20, 300, 5000) if big else (10, 150, 2000)
cfg = cabi.default_config(max_kf=N, max_feat=M, max_proj=M * 12, max_lidar=NL, device=local)
w = synth.make_window(4 if big else 2, 3, N=N, M=M, n_lidar=NL)
opts = cabi.default_solve_opts(cabi.VILS_MODE_GN, 5, 1e-8)
h = lib.BA(cfg, 1); h.set_window(0, split_window(w, world, rank)); h.upload(1)
buf = h.sharded_buffer_tensor()
for rep in range(3):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for it in range(5):
        h.sharded_linearize(it, opts)
        dist.all_reduce(buf)                  # NCCL sum over NVLink: D^2 + 2D + 1 doubles
        torch.cuda.synchronize()
        h.sharded_update(opts)
    dist.barrier(); dt = time.perf_counter() - t0
h.download(1); s = h.get_state(0)
lam = torch.tensor(s["inv_depth"], device="cuda"); own = torch.tensor((np.arange(M) % world) == rank, device="cuda")
lam = torch.where(own, lam, torch.zeros_like(lam)); dist.all_reduce(lam)
if rank == 0:
    import helpers
    full = lib.BA(cfg, 1); full.set_window(0, w); full.upload(1); full.solve_device(1, opts); t1 = full.last_ms; full.download(1); ref = full.get_state(0)
    s["inv_depth"] = lam.cpu().numpy()
    d = helpers.rel_state_delta(s, ref)
    print(f"world {world} N {N}: sharded GN-5 {dt * 1e3:.3f} ms vs single-GPU kernel {t1:.3f} ms; state delta {d:.2e}; buffer {buf.numel() * 8 / 1e3:.0f} KB")
    assert s["status"] == 0 and d <= 1e-9, d
    print("SHARDED_OK")
dist.destroy_process_group()
