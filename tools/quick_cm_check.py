"""Fast hardware sanity of the CM path: a few units per method through zq_compress_blocks, compared with the reference.
  python tools/quick_cm_check.py            default engine
  ZQ_CM_JIT=1 python tools/quick_cm_check.py    contexts from the translated HCOMP (NVRTC)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import zpaqfranz_b200 as zq  # noqa: E402
from zpaqfranz_b200 import corpus  # noqa: E402
import oracle_bindings as ob  # noqa: E402

ref = ob.load_ref()
units = [corpus.text_unit(1, 30000), corpus.mixed_unit(2, 12000), b"abc" * 500]
arena = np.frombuffer(b"".join(units), dtype=np.uint8)
lens = [len(u) for u in units]
offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
with zq.Context(0) as ctx:
    for method in ("3", "36,200,1", "4", "5"):
        t = time.time()
        out, ooff, olen = ctx.compress_blocks(arena, offs, lens, method=method, filename="", comment="")
        ok = all(out[int(ooff[i]): int(ooff[i]) + int(olen[i])].tobytes() == ref.compress_block(u, method, "", "")
                 for i, u in enumerate(units))
        print("method %-9s %s  %.2fs  launches %d  ZQ_CM_JIT=%s" % (method, "bit-exact" if ok else "MISMATCH", time.time() - t,
                                                                    ctx.launch_count(), os.environ.get("ZQ_CM_JIT", "0")), flush=True)
