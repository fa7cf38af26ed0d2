"""Factor-sharded solve of ONE window: partial reduced systems summed across ranks, identical Cholesky on
every rank.  (1) two handles on one GPU with a host-mediated sum; (2) two GPUs with torch.distributed all_reduce over NCCL
(skipped when fewer than 2 GPUs are visible).  Reference: the single-handle solve of the whole window, GN x5."""
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers
from mvil_fusion_b200 import cabi, synth
from mvil_fusion_b200.sharding import split_window

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_shards_one_gpu_host_sum():
    from mvil_fusion_b200 import lib
    cfg = cabi.default_config()
    w = synth.make_window(2, 7)
    opts = cabi.default_solve_opts(cabi.VILS_MODE_GN, 5, 1e-8)
    full = lib.BA(cfg, 1); full.set_window(0, w); full.solve(1, opts); ref = full.get_state(0)
    G = 2
    hs = []
    for r in range(G):
        h = lib.BA(cfg, 1); h.set_window(0, split_window(w, G, r)); h.upload(1); hs.append(h)
    for it in range(5):
        for h in hs:
            h.sharded_linearize(it, opts)
        total = sum(h.sharded_read() for h in hs)          # the all-reduce
        for h in hs:
            h.sharded_write(total); h.sharded_update(opts)
    states = []
    for h in hs:
        h.download(1); states.append(h.get_state(0))
    assert all(s["status"] == 0 for s in states)
    for k in ("pose", "speedbias", "ex_pose"):
        assert np.array_equal(states[0][k], states[1][k])     # identical Cholesky on identical reduced systems
    merged = dict(states[0]); lam = states[0]["inv_depth"].copy()
    owner1 = (np.arange(len(lam)) % G) == 1
    lam[owner1] = states[1]["inv_depth"][owner1]
    merged["inv_depth"] = lam
    assert helpers.rel_state_delta(merged, ref) <= 1e-9


@pytest.mark.skipif("__import__('torch').cuda.device_count() < 2")
def test_two_gpus_nccl_allreduce():
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29611", os.path.join(ROOT, "tools", "sharded_nccl_demo.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "SHARDED_OK" in out.stdout


@pytest.mark.skipif("__import__('torch').cuda.device_count() < 2")
def test_two_gpus_native_nccl_sharded_solve():
    """vils_ba_sharded_solve: the library's own NCCL communicator, all iterations on the library stream; every rank ends with the full
    state, identical to the single-GPU solve to 1e-9 (config 2 and config 4 with the real prior)."""
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29613", os.path.join(ROOT, "tools", "sharded_native.py"), "big"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "SHARDED_NATIVE_OK" in out.stdout


def test_native_sharded_solve_single_rank_communicator():
    """World size 1 exercises the whole native path (lazy libnccl binding, communicator, all-reduce on the library stream, depth
    exchange) on one GPU: the result must equal the ordinary solve."""
    from mvil_fusion_b200 import lib
    cfg = cabi.default_config()
    w = synth.make_window(2, 9)
    opts = cabi.default_solve_opts(cabi.VILS_MODE_GN, 5, 1e-8)
    full = lib.BA(cfg, 1); full.set_window(0, w); full.solve(1, opts); ref = full.get_state(0)
    h = lib.BA(cfg, 1)
    h.sharded_init(0, 1, lib.nccl_unique_id())
    h.set_window(0, w); h.upload(1)
    s = h.sharded_solve(opts)
    assert s.status == 0 and s.iterations == 5
    st = h.get_state(0)
    assert helpers.rel_state_delta(st, ref) <= 1e-9
    full.close(); h.close()
