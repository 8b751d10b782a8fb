"""Multi-GPU host logic on CPU: world_size-2 gloo.  Clips are sharded round-robin by rank, processed independently
(no data-path collective) and gathered to rank 0, which writes one CSV per clip with the reference's schema."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_clip_rows(c):
    rng = np.random.default_rng(c)
    return rng.normal((412.0, 290.0), 1.0, (5 + c, 2))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from lk_optical_flow_method_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_clips = 7
    mine = shard.assign(n_clips, world, rank)
    assert mine == list(range(rank, n_clips, world))
    local = {c: _fake_clip_rows(c) for c in mine}
    merged = shard.gather_to_root(local)
    if rank == 0:
        assert sorted(merged) == list(range(n_clips))
        paths = shard.write_vp_csvs(merged, [f"clip{c:02d}" for c in range(n_clips)], out_dir)
        assert len(paths) == n_clips
    else:
        assert merged is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for c in range(7):
        with open(tmp_path / f"vps_clip{c:02d}.csv", newline="") as f:
            text = f.read()
        assert text.startswith("x,y\r\n")
        got = np.loadtxt(tmp_path / f"vps_clip{c:02d}.csv", delimiter=",", skiprows=1).reshape(-1, 2)
        assert np.array_equal(got, _fake_clip_rows(c))            # full-precision round trip


def test_assignment_is_a_partition():
    from lk_optical_flow_method_b200 import shard
    for world in (1, 2, 4, 8):
        parts = [shard.assign(64, world, r) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(64))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert shard.gather_to_root({0: 1}) == {0: 1}                 # no process group: identity
