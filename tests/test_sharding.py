"""Host-side logic of the N>1 path on CPU: world_size-2 gloo processes shard frames round-robin, compute a
deterministic per-frame payload, gather to rank 0 and re-order by frame index (rtpose.cpp:1214-1273)."""
import os
import socket
import subprocess
import sys

from caffe_rtpose_b200.sharding import Orderer, shard_indices

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import torch.distributed as dist
from caffe_rtpose_b200.sharding import shard_indices, gather_ordered
from caffe_rtpose_b200.engine import write_json
import numpy as np
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 11
local = []
for i in reversed(shard_indices(n, rank, world)):      # finish out of order on purpose
    joints = np.full((1, 18, 3), float(i), np.float32)
    local.append((i, write_json(joints, 18)))           # host-only ABI call, no GPU needed
out = gather_ordered(local, rank, world, dist)
if rank == 0:
    assert [i for i, _ in out] == list(range(n)), out
    assert all(('"joints":[%%d,%%d,%%d' %% (i, i, i)) in js for i, js in out)
    print("ORDER_OK", len(out))
dist.barrier()
dist.destroy_process_group()
'''


def test_shard_indices_cover_all_frames_once():
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in shard_indices(37, r, world))
        assert seen == list(range(37))


def test_orderer_skips_dropped_frames():
    o = Orderer(dropped={2})
    assert o.push(1, "b") == []
    assert o.push(0, "a") == [(0, "a"), (1, "b")]
    assert o.push(4, "e") == []
    assert o.push(3, "d") == [(3, "d"), (4, "e")]


def test_two_rank_gloo_gather_ordered(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "ORDER_OK 11" in r.stdout
