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
