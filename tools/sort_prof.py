"""Phase breakdown of k_suffix_sort (needs a library built with ZQ_EXTRA_FLAGS=-DZQ_SORT_PROF):
python tools/sort_prof.py [--units U] [--method 2]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import zpaqfranz_b200 as zq  # noqa: E402
from zpaqfranz_b200 import corpus  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--units", type=int, default=3000)
ap.add_argument("--unit", type=int, default=65536)
ap.add_argument("--method", default="2")
a = ap.parse_args()
h = corpus.text_corpus(a.units, a.unit)
arena = torch.from_numpy(h).cuda()
cap = int(zq.lib.zq_compress_bound(a.unit)) * a.units
out = torch.empty(cap, dtype=torch.uint8, device="cuda")
offs = np.arange(a.units, dtype=np.uint64) * a.unit
lens = np.full(a.units, a.unit, dtype=np.uint32)
ctx = zq.Context(0)
prof = (C.c_ulonglong * 8)()
for s in range(2):
    ctx.compress_blocks_device(arena.data_ptr(), offs, lens, out.data_ptr(), cap, method=a.method, filename="", comment="")
    zq.lib.zq_debug_sort_profile(prof)
v = list(prof)
cyc = sum(v[:6])
names = ["initial keys + 4-byte sort", "first ranking/compaction", "round key generation", "round sort", "round re-ranking", "lcp/bwt/isa output"]
print("sufsort ms", ctx.last_timings()["sufsort"], "units", a.units)
for n, c in zip(names, v[:6]):
    print("%-28s %5.1f %%   %8.0f cycles/unit" % (n, 100.0 * c / cyc, c / a.units))
print("rounds/unit %.2f   sum(m)/n %.3f" % (v[6] / a.units, v[7] / a.units / a.unit))
