"""The N>1 host logic of the C ABI (zq_dist_*, zpaqfranz_b200/csrc/zq_dist.cpp) on CPUs: world_size-2 gloo processes
shard the units, exchange compressed sizes for the archive offsets and fragment digests for the dedup decision --
the same C++ code that runs over NCCL on the GPUs, with the all-gather supplied by the test."""
import hashlib
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import zpaqfranz_b200 as zq


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_allgather(world):
    def ag(data):
        t = torch.frombuffer(bytearray(data), dtype=torch.uint8) if len(data) else torch.zeros(0, dtype=torch.uint8)
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [bytes(o.numpy().tobytes()) for o in out]
    return ag


def _digests(rank, n):
    # fragments 0..n-1 of each rank; every third fragment of rank 1 repeats one of rank 0, every fifth repeats locally
    out = []
    for k in range(n):
        key = ("r0-%d" % (k % 7)) if (rank == 1 and k % 3 == 0) else ("r%d-%d" % (rank, k // 2 if k % 5 == 0 else k))
        out.append(hashlib.sha1(key.encode()).digest())
    return np.frombuffer(b"".join(out), dtype=np.uint8).reshape(-1, 20)


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = zq.Dist(-1, rank, world, allgather=_gloo_allgather(world))
    lo, hi = zq.shard_range(total, rank, world)
    local = [1000 + 7 * u for u in range(lo, hi)]          # stand-in for this rank's compressed block sizes
    sizes, offs = d.exchange_sizes(local, total)
    first, uniq = d.dedup(_digests(rank, 9 + rank))
    q.put((rank, lo, hi, sizes.tolist(), offs.tolist(), first.tolist(), uniq, d.bytes_exchanged()))
    d.close()
    dist.destroy_process_group()


def test_two_rank_exchanges():
    world, total = 2, 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
    want = [1000 + 7 * u for u in range(total)]
    assert res[0][1:3] == (0, 6) and res[1][1:3] == (6, 11)
    for r in res:
        assert r[3] == want
        assert r[4] == np.concatenate([[0], np.cumsum(want)[:-1]]).tolist()
    # dedup against a plain Python model of "first occurrence in (rank, index) order"
    seen, firsts = set(), []
    for rank in range(world):
        f = []
        for dg in _digests(rank, 9 + rank):
            f.append(bytes(dg) not in seen)
            seen.add(bytes(dg))
        firsts.append(f)
    assert res[0][5] == firsts[0] and res[1][5] == firsts[1]
    assert res[0][6] == res[1][6] == len(seen)
    assert not all(firsts[1])          # the case is not vacuous: rank 1 holds duplicates of rank 0
    assert res[0][7] > 0


def _stitch_worker(rank, world, port, q):
    import oracle_bindings
    from zpaqfranz_b200 import corpus
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = oracle_bindings.load_oracle()
    data = corpus.text_bytes(17, 300_000).tobytes() + bytes(250_000) + corpus.text_bytes(18, 250_000).tobytes()
    fragment, total = 1, 800_000
    d = zq.Dist(-1, rank, world, allgather=_gloo_allgather(world))
    lo, hi = zq.shard_range(total, rank, world)
    avail = min(total, hi + 4 * (8128 << fragment)) if rank < world - 1 else total
    start, rounds = lo, 0
    while True:
        lens, _ = orc.fragment(data[start:avail], fragment)
        while True:
            st = d.stitch_fragments(total, lo, hi, start, avail, lens)
            rounds += 1
            if not st["again"] or st["restart"]:
                break
        if not st["again"]:
            break
        start = st["restart_at"]
    kept = lens[st["first_keep"]:st["first_keep"] + st["n_keep"]].tolist()
    q.put((rank, kept, st, rounds, orc.fragment(data, fragment)[0].tolist() if rank == 0 else None))
    d.close()
    dist.destroy_process_group()


def test_two_rank_stream_split():
    """One stream cut across two gloo processes (zq_dist_stitch_fragments): the cut lies inside a run of zeros, so the
    right rank is told where to start again; the kept fragments together are the whole stream's."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_stitch_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
    want = res[0][4]
    assert res[0][1] + res[1][1] == want
    assert res[1][2]["global_first"] == len(res[0][1]) and res[1][2]["global_total"] == len(want)
    assert res[0][2]["end"] == res[1][2]["begin"] == sum(res[0][1])
    assert res[1][3] == 2 and res[0][3] == 2           # one repetition: the restart of rank 1


def test_shard_helpers_and_single_rank():
    for total in (0, 1, 7, 64, 10000):
        for world in (1, 2, 3, 8):
            spans = [zq.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    costs = np.array([5, 1, 9, 3, 3, 7, 2, 8], dtype=np.uint64)
    owner = zq.shard_lpt(costs, 3)
    loads = [int(costs[owner == r].sum()) for r in range(3)]
    assert sorted(np.concatenate([np.nonzero(owner == r)[0] for r in range(3)]).tolist()) == list(range(8))
    assert max(loads) - min(loads) <= 3
    d = zq.Dist(-1, 0, 1)            # world == 1: no transport at all
    sizes, offs = d.exchange_sizes([5, 6, 7], 3)
    assert sizes.tolist() == [5, 6, 7] and offs.tolist() == [0, 5, 11]
    first, uniq = d.dedup(np.frombuffer(b"a" * 20 + b"b" * 20 + b"a" * 20, dtype=np.uint8))
    assert first.tolist() == [True, True, False] and uniq == 2
    d.close()
