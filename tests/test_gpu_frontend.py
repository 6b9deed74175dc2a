"""Front-end link proof: the reference's OWN archiver (its main(), compressThread, the journal writers) built with
libzpaq::compressBlock bound to libzqb200.so (oracle/_ref/libzpaqref_gpu.so: the reference TU piped through sed with
oracle/frontend_gpu_tail.cpp appended, the stub of INTEGRATION.md section 1) writes byte-identical archives to the
stock build, and its `x` restores the tree from them.  Test infrastructure drives both builds in child processes."""
import os
import subprocess
import sys

import pytest

from zpaqfranz_b200 import corpus

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")

RUNNER = ("import ctypes as C, sys\n"
          "lib = C.CDLL(sys.argv[1])\n"
          "a = [x.encode() for x in sys.argv[2:]]\n"
          "arr = (C.c_char_p * len(a))(*a)\n"
          "lib.zref_main.restype = C.c_int\n"
          "rc = lib.zref_main(len(a), arr)\n"
          "if hasattr(lib, 'zref_gpu_blocks'):\n"
          "    lib.zref_gpu_blocks.restype = C.c_ulonglong\n"
          "    sys.stderr.write('GPU_BLOCKS %d\\n' % lib.zref_gpu_blocks())\n"
          "sys.exit(rc)\n")


def zpaqfranz(libname, *args):
    r = subprocess.run([sys.executable, "-c", RUNNER, os.path.join(REF, libname), "zpaqfranz"] + [str(a) for a in args],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r


def make_tree(root):
    spec = {"a.txt": corpus.text_unit(1, 300000), "b.bin": corpus.random_unit(2, 100000), "sub/c.txt": corpus.text_unit(1, 300000),
            "empty.dat": b"", "z.dat": bytes(200000), "sub/deep/d.TXT": corpus.text_unit(5, 70000) + corpus.text_unit(1, 300000),
            "noext": corpus.repeats_unit(3, 50000), "e.exe": corpus.mixed_unit(9, 150000)}
    for rel, data in spec.items():
        p = os.path.join(root, rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "wb") as f:
            f.write(data)
        os.utime(p, (1700000000, 1700000000))
    return spec


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "libzpaqref_gpu.so")), reason="front-end link build absent (oracle/Makefile: make frontend)")
@pytest.mark.parametrize("method", ["-m1", "-m2", "-m3", "-m4"])
def test_reference_front_end_on_the_device_compressor(tmp_path, method):
    src = tmp_path / "src"
    spec = make_tree(str(src))
    common = [str(src), method, "-timestamp", "20260101000000", "-t2"]
    zpaqfranz("libzpaqref.so", "a", tmp_path / "cpu.zpaq", *common)
    r = zpaqfranz("libzpaqref_gpu.so", "a", tmp_path / "gpu.zpaq", *common)
    nblocks = int(r.stderr.split("GPU_BLOCKS")[1].split()[0])
    assert nblocks >= 3                                     # data block(s) + fragment table + index went through libzqb200.so
    cpu, gpu = open(tmp_path / "cpu.zpaq", "rb").read(), open(tmp_path / "gpu.zpaq", "rb").read()
    assert gpu == cpu, "archives differ (%d vs %d bytes)" % (len(gpu), len(cpu))
    # and the device-written archive extracts with the front end's own `x`
    out = tmp_path / "out"
    zpaqfranz("libzpaqref_gpu.so", "x", tmp_path / "gpu.zpaq", "-to", out, "-space")
    restored = 0
    for dirpath, _, files in os.walk(out):
        for f in files:
            rel = os.path.relpath(os.path.join(dirpath, f), out)
            key = next((k for k in spec if rel.endswith(k)), None)
            assert key is not None, rel
            assert open(os.path.join(dirpath, f), "rb").read() == spec[key], rel
            restored += 1
    assert restored == len(spec)
