"""Secondary measurements for BASELINE.json configs[2..4] (scaled to fit a few GPU-minutes):
  C3  enwik-shaped synthetic corpus, -m3, compress on GPU -> decompress on GPU -> compare   (64 KiB units)
  C4  dedup fragmenter + fragment SHA-1 + per-file BLAKE3 over a synthetic filesystem image
  C5  mixed-entropy 64 KiB fragments, -m5
Prints one JSON object; the parity of every leg is checked against the reference (oracle/_ref) on samples."""
import argparse
import json
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

ap = argparse.ArgumentParser()
ap.add_argument("--c3-mb", type=int, default=100)
ap.add_argument("--c4-mb", type=int, default=2000)
ap.add_argument("--c5-units", type=int, default=1500)
ap.add_argument("--skip", default="")
a = ap.parse_args()
ref = ob.load_ref()
ctx = zq.Context(0)
res = {}
UNIT = 65536


def wall(fn):
    t = time.perf_counter()
    r = fn()
    return r, time.perf_counter() - t


if "c3" not in a.skip:
    n = a.c3_mb * 1000000 // UNIT
    arena = corpus.text_corpus(n)
    offs = np.arange(n, dtype=np.uint64) * UNIT
    lens = np.full(n, UNIT, dtype=np.uint32)
    for method in ("3", "36,200,1"):
        (comp, coff, clen), tc = wall(lambda: ctx.compress_blocks(arena, offs, lens, method=method, filename="", comment="jDC\x01"))
        tms = ctx.last_timings()
        (out, ooff, olen), td = wall(lambda: ctx.decompress_blocks(comp, coff, clen))
        ok = out[: n * UNIT].tobytes() == arena.tobytes()
        i = n // 2
        exact = ref is None or comp[int(coff[i]): int(coff[i]) + int(clen[i])].tobytes() == ref.compress_block(arena[i * UNIT:(i + 1) * UNIT].tobytes(), method, "", "jDC\x01")
        res["c3_" + method] = {"units": n, "bytes": n * UNIT, "ratio": round(float(clen.sum()) / (n * UNIT), 4),
                               "compress_MBps_e2e": round(n * UNIT / 1e6 / tc, 1), "decompress_MBps_e2e": round(n * UNIT / 1e6 / td, 1),
                               "compress_stage_ms": {k: round(v, 1) for k, v in tms.items()}, "round_trip_ok": bool(ok), "bit_exact_sample": bool(exact)}

if "c4" not in a.skip:
    rng = np.random.Generator(np.random.PCG64(4))
    target = a.c4_mb * 1000000
    sizes = []
    tot = 0
    while tot < target:
        s = int(min(rng.lognormal(np.log(16384), 2.0), 256 << 20, target - tot + 1))
        s = max(s, 1)
        sizes.append(s)
        tot += s
    sizes = np.array(sizes, dtype=np.uint64)
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64)
    arena = np.empty(int(sizes.sum()), dtype=np.uint8)
    kinds = rng.integers(0, 10, len(sizes))
    t0 = time.perf_counter()
    for f, (o, s, kd) in enumerate(zip(offs, sizes, kinds)):
        o, s = int(o), int(s)
        if kd < 4:
            arena[o:o + s] = corpus.text_bytes(f, s)
        elif kd < 7:
            arena[o:o + s] = np.random.Generator(np.random.PCG64(f)).integers(0, 256, s, dtype=np.uint8)
        elif kd < 9 and f > 10:
            src = int(rng.integers(0, f))
            so, ss = int(offs[src]), int(sizes[src])
            k = min(s, ss)
            arena[o:o + k] = arena[so:so + k]
            arena[o + k:o + s] = 0
        else:
            arena[o:o + s] = 0
    gen_s = time.perf_counter() - t0
    (fr, tfrag) = wall(lambda: ctx.fragment(arena, offs, sizes, fragment=6, want_sha1=True))
    fl, fh, fs, first = fr
    (b3, tb3) = wall(lambda: ctx.blake3(arena, offs, sizes))
    # parity on a sample of files
    ok = True
    orc = ob.load_oracle()
    for f in list(range(0, len(sizes), max(1, len(sizes) // 12)))[:12]:
        d = arena[int(offs[f]): int(offs[f]) + int(sizes[f])].tobytes()
        ol, oh = orc.fragment(d, 6)
        x, y = int(first[f]), int(first[f + 1])
        ok = ok and (y - x == len(ol)) and bool((fl[x:y] == ol).all()) and (ref is None or b3[f].tobytes() == ref.blake3(d))
        if y > x:
            ok = ok and fs[x].tobytes() == orc.sha1(d[: int(fl[x])])
    uniq = len({bytes(r) for r in fs})
    res["c4"] = {"files": len(sizes), "bytes": int(sizes.sum()), "fragments": int(len(fl)), "unique_fragments": uniq,
                 "fragment_sha1_GBps_e2e": round(float(sizes.sum()) / 1e9 / tfrag, 2), "blake3_GBps_e2e": round(float(sizes.sum()) / 1e9 / tb3, 2),
                 "parity_sample_ok": bool(ok), "gen_s": round(gen_s, 1)}

if "c5" not in a.skip:
    n = a.c5_units
    units = [corpus.mixed_unit(s, UNIT) for s in range(n)]
    arena = np.frombuffer(b"".join(units), dtype=np.uint8)
    offs = np.arange(n, dtype=np.uint64) * UNIT
    lens = np.full(n, UNIT, dtype=np.uint32)
    (r5, t5) = wall(lambda: ctx.compress_blocks(arena, offs, lens, method="5", filename="", comment=""))
    comp, coff, clen = r5
    tms = ctx.last_timings()
    exact = True
    if ref is not None:
        for i in (0, 1, 2, 3):
            exact = exact and comp[int(coff[i]): int(coff[i]) + int(clen[i])].tobytes() == ref.compress_block(units[i], "5", "", "")
    res["c5"] = {"units": n, "bytes": n * UNIT, "ratio": round(float(clen.sum()) / (n * UNIT), 4), "MBps_e2e": round(n * UNIT / 1e6 / t5, 2),
                 "stage_ms": {k: round(v, 1) for k, v in tms.items()}, "bit_exact_sample": bool(exact)}
print(json.dumps(res))
ctx.close()
