"""The -m2 DEVICE pipeline (k_suffix_sort -> k_lz77_sa, zpaqfranz_b200/csrc/zq_sufsort.cuh / zq_lz77.cuh) compiled for the
host through the SIMT emulator (tests/emu/simt_emu.h): suffix array and LZ77 stream must equal the oracle's, no GPU
needed.  Test infrastructure only -- nothing on the product path uses the emulator."""
import ctypes as C
import os
import subprocess

import pytest

import zpaqfranz_b200 as zq
from zpaqfranz_b200 import corpus

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
CSRC = os.path.join(ROOT, "zpaqfranz_b200", "csrc")


@pytest.fixture(scope="module")
def emu():
    out = os.path.join(EMU, "_build")
    os.makedirs(out, exist_ok=True)
    lib = os.path.join(out, "liblzemu.so")
    deps = [os.path.join(EMU, "lz_emu.cpp"), os.path.join(EMU, "simt_emu.h")] + [os.path.join(CSRC, f) for f in
                                                                                   ("zq_lz77.cuh", "zq_lz77_scan.cuh", "zq_sufsort.cuh", "zq_sufsort16.cuh", "zq_common.cuh", "zq_frame.cuh")]
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-O1", "-std=c++17", "-w", "-I" + os.path.join(EMU, "shim"), "-I" + CSRC,
                        "-I" + os.path.join(ROOT, "include"), "-shared", "-fPIC", "-o", lib, deps[0]], check=True)
    h = C.CDLL(lib)
    h.emu_lz_sa.restype = C.c_long
    h.emu_lz_scan.restype = C.c_long
    h.emu_lz_hash.restype = C.c_long
    h.emu_bwt.restype = C.c_long
    return h


CASES = [("2", corpus.text_unit(3, 3000)), ("2", b"abracadabra" * 100), ("2", bytes(2000)), ("2", corpus.mixed_unit(4, 2500)),
         ("2", corpus.random_unit(6, 1500)), ("2", b""), ("2", b"a"), ("2", b"ab" * 700 + b"c" + b"ab" * 700),
         ("3", corpus.mixed_unit(9, 3000)),          # byte-aligned codes, minMatch 12
         ("2", corpus.text_unit(8, 9000))]


@pytest.mark.parametrize("k", range(len(CASES)))
def test_suffix_array_and_stream_match_oracle(emu, oracle, k):
    method, data = CASES[k]
    plan = zq.plan_block(method, data)
    assert (plan["args"][1] & 3) in (1, 2) and plan["args"][5] - plan["args"][0] >= 21, plan["method"]   # SA search variant
    n = len(data)
    sa = (C.c_uint32 * max(n, 1))()
    cap = 2 * n + 4096
    out = (C.c_uint8 * cap)()
    r = emu.emu_lz_sa(data, n, (C.c_int * 9)(*plan["args"]), sa, out, cap)
    assert r >= 0
    assert list(sa[:n]) == list(oracle.suffix_array(data))
    assert bytes(out[:r]) == oracle.lz_stream(data, plan["args"])


# the position-parallel pipeline (zq_lz77_scan.cuh: k_lz_scan<0> -> k_lz_scan<1> -> k_lz_walk -> k_lz_emit) on the same
# cases plus: several tiles, look-ahead 0, small buckets, byte codes with matches, capped LCPs (deferred to the walk)
SCAN_CASES = CASES + [("x0,1,4,0,3,21,0", corpus.text_unit(3, 6000)), ("x0,2,5,0,6,21,1", corpus.text_unit(4, 7000)),
                      ("x0,2,5,0,6,21,1", corpus.repeats_unit(3, 14000)), ("x0,1,6,0,7,21,1", corpus.repeats_unit(3, 9000)),
                      ("x0,1,4,0,5,21,0", corpus.mixed_unit(6, 8000)),
                      ("2", corpus.text_unit(11, 9000) + bytes(3000) + corpus.text_unit(11, 5000)), ("2", bytes(range(256)) * 40)]


@pytest.mark.parametrize("k", range(len(SCAN_CASES)))
def test_scan_pipeline_matches_oracle(emu, oracle, k):
    method, data = SCAN_CASES[k]
    plan = zq.plan_block(method, data)
    a = plan["args"]
    assert (a[1] & 3) in (1, 2) and a[5] - a[0] >= 21 and a[4] <= 7 and a[6] <= 1, plan["method"]
    n = len(data)
    cap = n + n // 32 + 64
    out = (C.c_uint8 * (cap + 64))()
    ntok = C.c_uint32(0)
    r = emu.emu_lz_scan(data, n, (C.c_int * 9)(*a), out, cap, C.byref(ntok))
    assert r >= 0
    assert bytes(out[:r]) == oracle.lz_stream(data, a)


@pytest.mark.parametrize("n", [65536, 66000])
def test_scan_pipeline_full_block(emu, oracle, n):
    # a whole 64 KiB text block (16 tiles, 16-bit indices) and one just past it (32-bit indices, stream OR-ed in place)
    data = corpus.text_unit(21, n)
    a = zq.plan_block("2", data)["args"]
    cap = n + n // 32 + 64
    out = (C.c_uint8 * (cap + 64))()
    ntok = C.c_uint32(0)
    r = emu.emu_lz_scan(data, n, (C.c_int * 9)(*a), out, cap, C.byref(ntok))
    assert r >= 0 and bytes(out[:r]) == oracle.lz_stream(data, a)


SORT16_CASES = [corpus.text_unit(3, 3000), corpus.random_unit(6, 5000), b"abracadabra" * 100, bytes(2000), b"a", b"ab", corpus.text_unit(8, 20000),
                corpus.mixed_unit(6, 9000), bytes(range(256)) * 40, corpus.text_unit(5, 30000)[:-3] + b"\0\0\0", corpus.repeats_unit(2, 12000),
                corpus.text_unit(21, 65536)]


def _skewed(seed, k, n):
    """n bytes over k byte values, Zipf-like: the dense form of the sort packs them into 1..8 bits per character"""
    import numpy as np
    rng = np.random.default_rng(seed)
    vals = rng.choice(256, size=k, replace=False)
    p = 1.0 / np.arange(1, k + 1) ** 1.2
    return vals[rng.choice(k, size=n, p=p / p.sum())].astype(np.uint8).tobytes()


SORT16_CASES += [_skewed(1, 2, 2180), _skewed(2, 3, 9000), _skewed(3, 5, 8471), _skewed(4, 9, 13591), _skewed(5, 17, 20000), _skewed(6, 33, 10392),
                 _skewed(7, 65, 7468), _skewed(8, 129, 14255), _skewed(9, 256, 11135), corpus.text_unit(77, 12000) + corpus.random_unit(3, 6000),
                 _skewed(10, 17, 65536)]
# block ends and the zero padding of the character stream: runs of the smallest byte value at the end, NULs, tiny blocks
_T9 = corpus.text_unit(5, 9000)
SORT16_CASES += [_T9 + bytes([min(_T9)]) * 7, _T9 + bytes([min(_T9)]) * 70, b"\0" * 3 + _T9 + b"\0" * 9, _T9[:4000] + b"\0" * 40 + _T9[4000:] + b"\0",
                 corpus.text_unit(6, 15), corpus.text_unit(6, 17), corpus.text_unit(6, 33), corpus.text_unit(6, 1025), corpus.text_unit(6, 16385)]


@pytest.mark.parametrize("k", range(len(SORT16_CASES)))
def test_shared_memory_sort_matches_oracle(emu, oracle, k):
    # k_suffix_sort16 (zq_sufsort16.cuh): bins -> bitonic sort on 6-byte prefixes -> direct comparison of ties, with the hand-over
    # to k_suffix_sort for blocks with long repeats; suffix array, inverse, capped LCP, BWT bytes and packed rows
    import numpy as np
    data = bytes(SORT16_CASES[k])
    n = len(data)
    sa, isa, lcp = (np.zeros(max(n, 1), np.uint16) for _ in range(3))
    bwt, pk = np.zeros(max(n, 1), np.uint8), np.zeros(max(n, 1), np.uint32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731
    emu.emu_sort16(data, n, vp(sa), vp(isa), vp(lcp), vp(bwt), vp(pk))
    want = oracle.suffix_array(data)
    assert (sa[:n] == want).all()
    inv = np.zeros(n, np.int64)
    inv[want] = np.arange(n)
    assert (isa[:n] == inv).all()
    wl = np.zeros(n, np.int64)
    for j in range(1, n):
        a, b, l = int(want[j - 1]), int(want[j]), 0
        while a + l < n and b + l < n and l < 256 and data[a + l] == data[b + l]:
            l += 1
        wl[j] = l
    assert (lcp[:n] == wl).all()
    wb = np.array([data[int(x) - 1] if x > 0 else 0 for x in want], dtype=np.uint8) if n else np.zeros(0, np.uint8)
    assert (bwt[:n] == wb).all()
    assert (pk[:n] == (want.astype(np.uint32) | (np.minimum(wl, 255).astype(np.uint32) << 16) | (wb.astype(np.uint32) << 24))).all()


HASH_CASES = [("1", corpus.text_unit(3, 3000)), ("1", b"abracadabra" * 100), ("1", bytes(2000)), ("1", corpus.mixed_unit(4, 4000)),
              ("1", corpus.random_unit(6, 1500)), ("1", b""), ("1", b"abc"), ("24,40,0", corpus.text_unit(2, 3000)),
              ("14,200,3", corpus.text_unit(5, 2500))]


@pytest.mark.parametrize("k", range(len(HASH_CASES)))
def test_hash_table_parse_matches_oracle(emu, oracle, k):
    method, data = HASH_CASES[k]
    plan = zq.plan_block(method, data)
    a = plan["args"]
    if (a[1] & 3) not in (1, 2) or a[5] - a[0] >= 21:
        pytest.skip("not the hash-table variant: " + plan["method"])
    work = bytearray(data)
    if a[1] & 4:
        emu.emu_e8e9((C.c_uint8 * max(len(work), 1)).from_buffer(work), len(work)) if work else None
        assert bytes(work) == oracle.e8e9(data)
    cap = 2 * len(data) + 4096
    out = (C.c_uint8 * cap)()
    r = emu.emu_lz_hash(bytes(work), len(work), (C.c_int * 9)(*a), out, cap)
    assert r >= 0
    assert bytes(out[:r]) == oracle.lz_stream(data, a)     # the oracle filters E8E9 itself when args[1] says so


@pytest.mark.parametrize("data", [corpus.text_unit(3, 3000), b"abracadabra" * 50, bytes(700), b"", b"z", corpus.random_unit(2, 999)])
def test_bwt_stream_matches_oracle(emu, oracle, data):
    plan = zq.plan_block("36,200,1", data)
    assert (plan["args"][1] & 3) == 3
    out = (C.c_uint8 * (len(data) + 16))()
    r = emu.emu_bwt(data, len(data), out)
    assert bytes(out[:r]) == oracle.lz_stream(data, plan["args"])


def test_e8e9_filter_matches_oracle(emu, oracle):
    import random
    rnd = random.Random(5)
    data = bytearray(corpus.random_unit(9, 6000))
    for _ in range(300):                       # plant call/jump opcodes with small displacements
        p = rnd.randrange(0, len(data) - 5)
        data[p] = rnd.choice([0xE8, 0xE9]); data[p + 4] = rnd.choice([0, 255])
    work = bytearray(data)
    emu.emu_e8e9((C.c_uint8 * len(work)).from_buffer(work), len(work))
    assert bytes(work) == oracle.e8e9(bytes(data))
