"""Parity of the hash and fragmenter kernels with the reference's own implementations (oracle/_ref)
and its known-answer vectors (autotest "ABCDE" values, Z:77129-77160)."""
import os

import numpy as np
import pytest

from zpaqfranz_b200 import corpus

pytestmark = pytest.mark.gpu

LENS = [0, 1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 32, 33, 55, 56, 63, 64, 65, 96, 97, 127, 128, 129, 160, 239, 240, 241, 255,
        256, 511, 512, 1023, 1024, 1025, 1087, 1088, 2047, 2048, 2049, 3000, 4096, 5000, 65536, 100000, 300001, 1 << 20, (1 << 20) + 77]


def _bufs():
    bufs = [corpus.random_unit(1000 + n, n) for n in LENS] + [b"ABCDE", bytes(70000), corpus.text_unit(3, 40000)]
    lens = np.array([len(b) for b in bufs], dtype=np.uint64)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
    arena = np.frombuffer(b"".join(bufs) + b"\0", dtype=np.uint8)
    return bufs, arena, offs, lens


def test_known_answers(ctx):
    a = np.frombuffer(b"ABCDE", dtype=np.uint8)
    assert ctx.sha1(a, [0], [5])[0].tobytes().hex().upper() == "7BE07AAF460D593A323D0DB33DA05B64BFDCB3A5"
    assert ctx.sha256(a, [0], [5])[0].tobytes().hex().upper().startswith("F0393FEB")
    assert ctx.xxh3_128(a, [0], [5])[0].tobytes().hex().upper() == "1C8288B6013152D97B4A5D7E6C7893D4"
    assert ctx.blake3(a, [0], [5])[0].tobytes().hex().upper().startswith("61274278")


@pytest.mark.parametrize("algo", ["sha1", "sha256", "xxh3_128", "blake3"])
def test_hash_vs_reference(ctx, ref, algo):
    bufs, arena, offs, lens = _bufs()
    got = getattr(ctx, algo)(arena, offs, lens)
    want = getattr(ref, algo)
    for i, b in enumerate(bufs):
        assert got[i].tobytes() == want(b), (algo, len(b))


def test_fragmenter_matches_oracle_and_reference(ctx, oracle, ref):
    files = [corpus.text_unit(9, 900000) + corpus.random_unit(9, 700000) + bytes(1500000) + corpus.repeats_unit(9, 400000),
             corpus.text_unit(1, 5000), b"", corpus.random_unit(2, 65536 * 3 + 11), bytes(8128 * 64 * 2 + 5), b"x",
             corpus.text_unit(4, 2_000_000)]
    lens = np.array([len(f) for f in files], dtype=np.uint64)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
    arena = np.frombuffer(b"".join(files) + b"\0", dtype=np.uint8)
    for frag in (6, 4, 8, 0):
        fl, fh, fs, first = ctx.fragment(arena, offs, lens, fragment=frag)
        for f, data in enumerate(files):
            a, b = int(first[f]), int(first[f + 1])
            ol, oh = oracle.fragment(data, frag)
            rl, rh = ref.fragment(data, frag)
            assert (ol == rl).all() and (oh == rh).all()
            assert b - a == len(ol), (frag, f)
            assert (fl[a:b] == ol).all() and (fh[a:b] == oh).all(), (frag, f)
            pos = 0
            for k in range(a, min(b, a + 40)):
                assert fs[k].tobytes() == oracle.sha1(data[pos:pos + int(fl[k])]), (frag, f, k)
                pos += int(fl[k])


def test_crc32_xxh64_match_reference(ctx, ref):
    import zlib
    bufs = [b"", b"ABCDE", bytes(5000), corpus.random_unit(3, 4096).tobytes() if hasattr(corpus.random_unit(3, 4096), "tobytes") else bytes(corpus.random_unit(3, 4096)),
            bytes(corpus.text_unit(4, 70001)), bytes(corpus.random_unit(5, 1 << 20))]
    lens = np.array([len(b) for b in bufs], dtype=np.uint64)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
    arena = np.frombuffer(b"".join(bufs) + b"\0", dtype=np.uint8)
    crc = ctx.crc32(arena, offs, lens)
    xx = ctx.xxh64(arena, offs, lens)
    for i, b in enumerate(bufs):
        assert crc[i].tobytes() == zlib.crc32(b).to_bytes(4, "little")
        if ref is not None:
            assert xx[i].tobytes() == ref.xxh64(b)


def test_md5_sha3_match_reference(ctx, ref):
    import hashlib
    bufs = [b"", b"ABCDE", bytes(5000), bytes(corpus.text_unit(4, 70001)), bytes(corpus.random_unit(5, 1 << 20))] + \
           [bytes(corpus.random_unit(80, k)) for k in (55, 56, 64, 119, 135, 136, 137)]
    lens = np.array([len(b) for b in bufs], dtype=np.uint64)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
    arena = np.frombuffer(b"".join(bufs) + b"\0", dtype=np.uint8)
    m, s3 = ctx.md5(arena, offs, lens), ctx.sha3_256(arena, offs, lens)
    for i, b in enumerate(bufs):
        assert m[i].tobytes() == hashlib.md5(b).digest()
        assert s3[i].tobytes() == hashlib.sha3_256(b).digest()
        if ref is not None:
            assert m[i].tobytes() == ref.md5(b) and s3[i].tobytes() == ref.sha3_256(b)


def test_fragment_index_first_occurrence(ctx):
    """zq_dedup_first over 1 000 000 digests drawn from 200 000 (plus a run sharing one home slot): the earliest equal
    fragment, whatever order the threads insert in -- the HTIndex answer of Jidac::add (zpaqfranz.cpp:71567-71604)."""
    rng = np.random.default_rng(11)
    pool = rng.integers(0, 256, size=(200000, 20), dtype=np.uint8)
    pool[1000:1400, :8] = pool[0, :8]
    pick = rng.integers(0, len(pool), size=1000000)
    dg = np.ascontiguousarray(pool[pick])
    first = ctx.dedup_first(dg)
    _, idx = np.unique(pick, return_index=True)          # first position of every pool entry that was drawn
    want = np.zeros(len(pool), dtype=np.int64)
    want[np.unique(pick)] = idx
    assert np.array_equal(first.astype(np.int64), want[pick])
    assert ctx.dedup_first(np.zeros((0, 20), dtype=np.uint8)).size == 0
