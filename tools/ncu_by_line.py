#!/usr/bin/env python
"""Aggregate an ncu source-page (SASS) CSV by CUDA source line using nvdisasm line info.

    ncu -i X.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all rrt_b200/librrtb200.so ; nvdisasm --print-line-info -c rrtb_render.sm_100a.cubin > render.sass
    python tools/ncu_by_line.py src.csv render.sass <mangled kernel name fragment> [top N]
"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50

# instruction offset -> (file, line, inline chain) from nvdisasm
lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l)
cur = ("?", 0)
offs = {}
for l in lines[start + 1:]:
    if l.startswith("//---") or (l.startswith(".text.") and kern not in l):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)), "inlined" in m.group(3))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        offs[int(m.group(1), 16)] = (cur, m.group(2))

rows = list(csv.reader(open(src_csv)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ci = {n: hdr.index(n) for n in ("Address", "Source", "# Samples", "Instructions Executed", "Thread Instructions Executed")}
base = None
agg = defaultdict(lambda: [0.0, 0.0, 0.0])
tot = [0.0, 0.0, 0.0]
for r in rows[h + 1:]:
    if len(r) < len(hdr):
        continue
    addr = int(r[ci["Address"]], 16)
    if base is None:
        base = addr
    key = offs.get(addr - base, (("?", 0, False), ""))[0]
    v = [float(r[ci["Instructions Executed"]] or 0), float(r[ci["Thread Instructions Executed"]] or 0), float(r[ci["# Samples"]] or 0)]
    for k in range(3):
        agg[key[:2]][k] += v[k]
        tot[k] += v[k]
print("total warp-inst %.3e  thread-inst %.3e  avg lanes %.2f  samples %d" % (tot[0], tot[1], tot[1] / tot[0], tot[2]))
srcs = {}
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in srcs:
        try:
            srcs[f] = open("/root/repo/rrt_b200/csrc/" + f).read().split("\n")
        except Exception:
            srcs[f] = []
    text = srcs[f][ln - 1].strip()[:90] if 0 < ln <= len(srcs[f]) else ""
    print("%5.2f%% inst %5.2f%% smp  lanes %5.1f  %s:%d  %s" % (100 * v[0] / tot[0], 100 * v[2] / max(tot[2], 1), v[1] / max(v[0], 1), f, ln, text))
