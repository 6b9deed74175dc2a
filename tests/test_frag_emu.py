"""The dedup fragmenter's DEVICE code (zpaqfranz_b200/csrc/zq_fragment.cuh: speculate-and-stitch rounds, the closed form
for constant runs) on the host through the SIMT emulator; fragment lengths and predictor hits must equal the oracle's.
tests/emu/frag_emu.cpp mirrors the round loop of zq_fragment_ex.  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import pytest

from zpaqfranz_b200 import corpus

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
CSRC = os.path.join(ROOT, "zpaqfranz_b200", "csrc")


@pytest.fixture(scope="module")
def emu():
    out = os.path.join(EMU, "_build")
    os.makedirs(out, exist_ok=True)
    lib = os.path.join(out, "libfragemu.so")
    deps = [os.path.join(EMU, "frag_emu.cpp"), os.path.join(EMU, "simt_emu.h"), os.path.join(CSRC, "zq_fragment.cuh"),
            os.path.join(CSRC, "zq_common.cuh")]
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-O1", "-std=c++17", "-w", "-I" + os.path.join(EMU, "shim"), "-I" + CSRC,
                        "-I" + os.path.join(ROOT, "include"), "-shared", "-fPIC", "-o", lib, deps[0]], check=True)
    h = C.CDLL(lib)
    h.emu_fragment.restype = C.c_long
    return h


TEXT = bytes(corpus.text_bytes(1, 60000))
MIX = TEXT[:50000] + bytes(120000) + TEXT[:30000] + b"\xff" * 90000 + b"\x07" * 50000 + TEXT[:1000]
CASES = [
    (bytes(300000), 0, 8192, 4),            # one long zero run: settles in a few rounds, not one per segment
    (bytes(300000), 2, 16384, 4),
    (MIX, 0, 8192, 8), (MIX, 1, 4096, 8), (MIX, 3, 8192, 8),
    (bytes(70000) + b"\x01" + bytes(70000), 0, 8192, 8),
    (TEXT, 0, 4096, 4), (bytes(corpus.random_unit(3, 100000)), 0, 8192, 4),
    (b"ab" * 100000, 0, 8192, 64),          # periodic but not constant: still one segment per round
    (b"", 0, 8192, 1), (b"x", 0, 8192, 2), (bytes(5000), 6, 131072, 3),
]


@pytest.mark.parametrize("k", range(len(CASES)))
def test_fragments_match_oracle(emu, oracle, k):
    data, fragment, seg, max_rounds = CASES[k]
    n = len(data)
    cap = n // 32 + 100
    fl, fh, rounds = (C.c_uint32 * cap)(), (C.c_uint32 * cap)(), C.c_uint32(0)
    r = emu.emu_fragment(data, C.c_uint64(n), fragment, C.c_uint32((1 << 26) - 4096), C.c_uint64(seg), fl, fh, C.c_uint64(cap),
                         C.byref(rounds))
    ol, oh = oracle.fragment(data, fragment)
    assert r == len(ol)
    assert list(fl[:r]) == list(ol) and list(fh[:r]) == list(oh)
    assert rounds.value <= max_rounds, rounds.value
