"""N>1 host-side plumbing on CPU: world_size-2 gloo process group; sharding, ragged gather, reductions, and the
row partition the SVGD bandwidth uses across ranks."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "ocbnn-public_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ocbnn_b200 import dist as D
    out = {}
    assert D.active() and D.rank() == rank and D.world() == world
    # contiguous shards of 7 chains over 2 ranks: 4 + 3
    lo, hi = D.shard(7)
    out["shard"] = (lo, hi)
    full = torch.arange(7 * 3, dtype=torch.float32).reshape(7, 3)
    mine = full[lo:hi].clone()
    got = D.gather_cat(mine, dim=0)                       # ragged: 4 and 3 rows
    out["gather0"] = torch.equal(got, full)
    out["gather_rows"] = torch.equal(D.gather_rows(mine, 7), full) and torch.equal(D.gather_rows(full[D.shard(6)[0]:D.shard(6)[1]].clone(), 6), full[:6])
    samples = full.reshape(1, 7, 3).repeat(2, 1, 1)       # (rounds, chains, D) sharded along dim 1
    got = D.gather_cat(samples[:, lo:hi].contiguous(), dim=1)
    out["gather1"] = torch.equal(got, samples)
    t = torch.tensor([float(rank + 1)])
    out["all_sum"] = D.all_sum(t)
    h = torch.full((4,), rank + 1, dtype=torch.int32)
    D.all_sum_(h)
    out["all_sum_"] = h.tolist()
    # cyclic row partition of the i<j triangle used by the multi-GPU median select
    n = 11
    rows = list(range(rank, n, world))
    out["pairs"] = sum(n - 1 - i for i in rows)
    q.put((rank, out))
    dist.destroy_process_group()


def test_gloo_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0]["shard"] == (0, 4) and res[1]["shard"] == (4, 7)
    for r in range(world):
        assert res[r]["gather0"] and res[r]["gather1"] and res[r]["gather_rows"]
        assert res[r]["all_sum"] == 3.0 and res[r]["all_sum_"] == [3, 3, 3, 3]
    assert res[0]["pairs"] + res[1]["pairs"] == 11 * 10 // 2


def test_single_process_is_a_noop():
    from ocbnn_b200 import dist as D
    assert not D.active() and D.rank() == 0 and D.world() == 1
    assert D.shard(10) == (0, 10)
    t = torch.ones(3)
    assert D.gather_cat(t) is t and D.gather_rows(t, 3) is t and D.all_sum(torch.tensor([2.0])) == 2.0
    assert [D.shard(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
