"""Quick device-resident timing of the -m2 step for tuning: python tools/quick_bench.py --units U [--steps K]
Prints the per-stage CUDA-event ms averaged over K steps (after 2 warm-ups) and checks parity on 3 units."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import zpaqfranz_b200 as zq  # noqa: E402
from zpaqfranz_b200 import corpus  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--units", type=int, default=10000)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--method", default="2")
ap.add_argument("--unit", type=int, default=65536)
ap.add_argument("--check", type=int, default=1)
a = ap.parse_args()
h = corpus.text_corpus(a.units, a.unit)
arena = torch.from_numpy(h).cuda()
cap = int(zq.lib.zq_compress_bound(a.unit)) * a.units
out = torch.empty(cap, dtype=torch.uint8, device="cuda")
offs = np.arange(a.units, dtype=np.uint64) * a.unit
lens = np.full(a.units, a.unit, dtype=np.uint32)
ctx = zq.Context(0)
acc = None
for s in range(a.steps + 2):
    oo, ol = ctx.compress_blocks_device(arena.data_ptr(), offs, lens, out.data_ptr(), cap, method=a.method, filename="", comment="")
    t = ctx.last_timings(ex=True)
    if s >= 2:
        acc = t if acc is None else {k: acc[k] + t[k] for k in t}
print({k: round(v / a.steps, 2) for k, v in acc.items()}, "MB/s=%.0f" % (a.units * a.unit / 1e6 / (acc["total"] / a.steps / 1e3)),
      "env", {k: v for k, v in os.environ.items() if k.startswith("ZQ_")})
if a.check:
    import oracle_bindings as ob
    ref = ob.load_ref()
    ho = out.cpu().numpy()
    for i in (0, a.units // 2, a.units - 1):
        want = ref.compress_block(h[i * a.unit:(i + 1) * a.unit].tobytes(), a.method, "", "")
        assert ho[int(oo[i]): int(oo[i]) + int(ol[i])].tobytes() == want, "unit %d differs" % i
    print("parity ok")
ctx.close()
