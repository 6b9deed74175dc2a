"""Attribute ncu per-SASS-instruction counts to CUDA source lines.
usage: python tools/ncu_lines.py <report.ncu-rep> <kernel-mangled-substring> [top]
Needs the library the report was taken from (zpaqfranz_b200/libzqb200.so, built with -lineinfo)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, ksub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "zpaqfranz_b200", "libzqb200.so")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.startswith("zq_api.")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# instructions of the kernel, in order, with their innermost source line
lines, cur, inside = [], None, False
for ln in dis:
    if ln.startswith(".text."):
        inside = ksub in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append((cur, ln.strip()))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if "Address" in r][0]
hdr = rows[hi]
iI, iS, iSrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
sass = [r for r in rows[hi + 1:] if len(r) > iI]
if len(sass) != len(lines):
    print("warning: %d sass rows vs %d disassembled instructions" % (len(sass), len(lines)))
agg = collections.defaultdict(lambda: [0, 0])
tot = 0
for (loc, _), r in zip(lines, sass):
    c = int(r[iI]) if r[iI].isdigit() else 0
    s = int(r[iS]) if r[iS].isdigit() else 0
    agg[loc][0] += c
    agg[loc][1] += s
    tot += c
srcs = {}
print("total warp instructions: %d" % tot)
for loc, (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if loc:
        if loc[0] not in srcs:
            p = os.path.join(ROOT, "zpaqfranz_b200", "csrc", loc[0])
            srcs[loc[0]] = open(p).read().splitlines() if os.path.exists(p) else []
        if 0 < loc[1] <= len(srcs[loc[0]]):
            text = srcs[loc[0]][loc[1] - 1].strip()[:100]
    print("%5.1f%% inst %5d smp  %s:%s  %s" % (100.0 * c / max(tot, 1), s, loc[0] if loc else "?", loc[1] if loc else "?", text))
