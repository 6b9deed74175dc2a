"""The N>1 host logic on CPU: world_size-2 gloo processes shard the units, exchange compressed sizes and
agree on global offsets and on the max-over-ranks time."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from zpaqfranz_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(total, rank, world)
    local = [1000 + 7 * u for u in range(lo, hi)]          # stand-in for this rank's compressed block sizes
    sizes, offs = sharding.exchange_sizes(local, total)
    t = sharding.max_over_ranks(10.0 + rank)
    q.put((rank, lo, hi, sizes.tolist(), offs.tolist(), t))
    dist.destroy_process_group()


def test_two_rank_size_exchange():
    world, total = 2, 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
    want = [1000 + 7 * u for u in range(total)]
    assert res[0][1:3] == (0, 6) and res[1][1:3] == (6, 11)
    for r in res:
        assert r[3] == want
        assert r[4] == np.concatenate([[0], np.cumsum(want)[:-1]]).tolist()
        assert r[5] == 11.0


def test_shard_helpers():
    for total in (0, 1, 7, 64, 10000):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    costs = np.array([5, 1, 9, 3, 3, 7, 2, 8], dtype=float)
    parts = sharding.shard_by_cost(costs, 3)
    assert sorted(np.concatenate(parts).tolist()) == list(range(8))
    loads = [costs[p].sum() for p in parts]
    assert max(loads) - min(loads) <= costs.max()
