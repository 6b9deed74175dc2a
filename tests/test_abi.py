"""The C-ABI library loads without a GPU and exports every symbol include/zq_b200.h declares."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_all_declared_symbols_exported(zq):
    hdr = open(os.path.join(ROOT, "include", "zq_b200.h")).read()
    names = set(re.findall(r"\b(zq_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 15
    for n in names:
        assert hasattr(zq.lib, n), n


def test_no_device_fails_loudly(zq):
    import torch
    if torch.cuda.is_available():
        return
    try:
        zq.Context(0)
    except zq.ZqError as e:
        assert e.code == zq.ZQ_E_NODEVICE
    else:
        raise AssertionError("Context() must fail without a CUDA device (no CPU fallback)")


def test_compress_bound_is_generous(zq):
    for n in (0, 1, 65536, 1 << 20):
        assert zq.lib.zq_compress_bound(n) >= n + n // 32 + 4 * (n // 65536 + 2) + 200


def test_cpp_mirror_links_and_refuses_to_run_without_a_device(zq, tmp_path):
    """include/libzpaq_b200.h (Reader/Writer/StringBuffer/SHA1/SHA256/Compressor/Decompresser/compressBlock/decompress)
    compiles as a plain C++ caller and links against the shared library; without a GPU the first device call throws
    (there is no CPU path to fall back to)."""
    import subprocess
    import torch
    exe = tmp_path / "facade_driver"
    subprocess.run(["g++", "-O0", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests/cpp/facade_driver.cpp"),
                    "-o", str(exe), "-L" + os.path.join(ROOT, "zpaqfranz_b200"), "-lzqb200",
                    "-Wl,-rpath," + os.path.join(ROOT, "zpaqfranz_b200")], check=True)
    if torch.cuda.is_available():
        return
    (tmp_path / "in.bin").write_bytes(b"hello")
    r = subprocess.run([str(exe), str(tmp_path / "in.bin"), str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "FAILED" in r.stdout and "CUDA" in r.stdout
