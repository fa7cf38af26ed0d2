"""The N>1 host-side path on CPU: world_size-2 gloo processes shard windows, reduce timings with MAX and gather states."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from mvil_fusion_b200.sharding import shard_range


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 256, 592):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                lo, hi = shard_range(total, world, r)
                assert 0 <= lo <= hi <= total and hi - lo in (total // world, total // world + 1)
                got += list(range(lo, hi))
            assert got == list(range(total))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from mvil_fusion_b200.sharding import gather_states, max_over_ranks, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(11, world, rank)
    states = np.arange(lo, hi, dtype=np.float64)[:, None] * np.ones((1, 5))     # stand-in for solved states of windows lo..hi
    t = max_over_ranks([10.0 + rank, 3.0 - rank])
    g = gather_states(states)
    dist.barrier()
    if rank == 0:
        q.put((t, [x.tolist() for x in g]))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    t, g = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert t == [11.0, 3.0]
    allw = np.concatenate([np.asarray(x) for x in g])
    assert allw.shape == (11, 5) and np.array_equal(allw[:, 0], np.arange(11))
