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
                                                                                   ("zq_lz77.cuh", "zq_sufsort.cuh", "zq_common.cuh")]
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-O1", "-std=c++17", "-w", "-I" + os.path.join(EMU, "shim"), "-I" + CSRC,
                        "-I" + os.path.join(ROOT, "include"), "-shared", "-fPIC", "-o", lib, deps[0]], check=True)
    h = C.CDLL(lib)
    h.emu_lz_sa.restype = C.c_long
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
