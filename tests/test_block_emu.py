"""One whole libzpaq::compressBlock through the DEVICE code on the host (SIMT emulator): SHA-1, E8E9, suffix sort, LZ77 /
BWT pre-pass, context-mixing coder, framing -- in the order zq_compress_blocks launches the kernels -- and the block bytes
must equal the reference's (oracle/_ref).  No GPU needed; test infrastructure only (tests/emu/block_emu.cpp)."""
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
    lib = os.path.join(out, "libblockemu.so")
    srcs = [os.path.join(EMU, "block_emu.cpp"), os.path.join(CSRC, "zq_cm_host.cpp"), os.path.join(CSRC, "zq_config.cpp")]
    deps = srcs + [os.path.join(EMU, "simt_emu.h")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-O1", "-std=c++17", "-w", "-I" + os.path.join(EMU, "shim"), "-I" + CSRC,
                        "-I" + os.path.join(ROOT, "include"), "-shared", "-fPIC", "-o", lib] + srcs, check=True)
    h = C.CDLL(lib)
    h.emu_block.restype = C.c_long
    return h


CASES = [("0", corpus.text_unit(1, 700)), ("1", corpus.text_unit(2, 3000)), ("2", corpus.text_unit(3, 3000)), ("2", bytes(2500)),
         ("2", b""), ("2", corpus.random_unit(4, 1200)), ("3", corpus.mixed_unit(5, 2500)), ("36,200,1", corpus.text_unit(6, 1500)),
         ("4", corpus.text_unit(7, 900)), ("24,60,2", corpus.random_unit(8, 2000)), ("14,200,3", corpus.text_unit(9, 2000))]


@pytest.mark.parametrize("k", range(len(CASES)))
def test_whole_block_equals_reference(emu, ref, k):
    if ref is None:
        pytest.skip("oracle/_ref not built")
    method, data = CASES[k]
    plan = zq.plan_block(method, data)
    if plan["method"].startswith("0"):
        pytest.skip("stored blocks have no pre-pass or model")
    header, pcomp = bytes(plan["header"]), bytes(plan["pcomp"])
    for fn, cm, sha in ((b"", b"", 1), (b"dir/file.txt", b"jDC\x01", 1), (b"x", b"", 0)):
        cap = 2 * len(data) + 4096
        out = (C.c_uint8 * cap)()
        r = emu.emu_block(data, len(data), (C.c_int * 9)(*plan["args"]), header, len(header), pcomp, len(pcomp), fn, cm if cm else None, sha,
                          out, cap)
        assert r > 0, r
        want = ref.compress_block(data, method, fn.decode(), cm.decode("latin1") if cm else None, dosha1=bool(sha))
        assert bytes(out[:r]) == want, (method, plan["method"], fn, len(data))
