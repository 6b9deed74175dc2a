"""Runs a few device-resident -m2 steps (for ncu captures): python tools/prof_step.py --units U --steps K"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import zpaqfranz_b200 as zq  # noqa: E402
from zpaqfranz_b200 import corpus  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--units", type=int, default=2000)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--method", default="2")
ap.add_argument("--unit", type=int, default=65536)
a = ap.parse_args()
arena = torch.from_numpy(corpus.text_corpus(a.units, a.unit) if a.method in ("1", "2") else np.frombuffer(b"".join(corpus.mixed_unit(s, a.unit) for s in range(a.units)), dtype=np.uint8).copy() if a.method == "5" else corpus.text_corpus(a.units, a.unit)).cuda()
cap = int(zq.lib.zq_compress_bound(a.unit)) * a.units
out = torch.empty(cap, dtype=torch.uint8, device="cuda")
offs = np.arange(a.units, dtype=np.uint64) * a.unit
lens = np.full(a.units, a.unit, dtype=np.uint32)
ctx = zq.Context(0)
for _ in range(a.steps):
    ctx.compress_blocks_device(arena.data_ptr(), offs, lens, out.data_ptr(), cap, method=a.method, filename="", comment="")
    print(ctx.last_timings())
ctx.close()
