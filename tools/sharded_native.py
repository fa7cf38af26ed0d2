"""torchrun --nproc-per-node G tools/sharded_native.py [big] — one window factor-sharded over G GPUs by the library itself
(vils_ba_sharded_solve: linearise -> ncclAllReduce -> update, all iterations on the library stream, one host sync); every rank ends
with the full state; rank 0 checks it against its own single-GPU solve of the whole window and prints one JSON line."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
from mvil_fusion_b200 import cabi, synth, lib
from mvil_fusion_b200.sharding import native_sharded_solve_benchmark

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
opts = cabi.default_solve_opts(cabi.VILS_MODE_GN, 5, 1e-8)
out = {}
for name in (["config2", "config4"] if "big" in sys.argv else ["config2"]):
    if name == "config4":
        import bench
        cfg = cabi.default_config(max_kf=21, max_feat=320, max_proj=4000, max_lidar=5000, device=local)
        w = bench.build_config4_windows(lib, cfg, 1, opts)[0]
    else:
        cfg = cabi.default_config(device=local)
        w = synth.make_window(2, 3)
    rec, st = native_sharded_solve_benchmark(lib, cfg, w, opts, rank, world)
    # every rank holds the full state: compare all of them with rank 0's single-GPU solve
    full = lib.BA(cfg, 1); full.set_window(0, w); full.upload(1)
    for _ in range(3):
        full.solve_device(1, opts)
    t1 = full.last_ms; full.download(1); ref = full.get_state(0)
    import helpers
    d = helpers.rel_state_delta(st, ref)
    dt = torch.tensor([d], dtype=torch.float64, device="cuda"); dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    rec.update(single_gpu_device_ms=t1, max_state_delta_vs_single_gpu=float(dt.item()))
    out[name] = rec
    assert rec["status"] == 0 and float(dt.item()) <= 1e-9, rec
if rank == 0:
    print(json.dumps(out))
    print("SHARDED_NATIVE_OK")
dist.destroy_process_group()
