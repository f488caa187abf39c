#!/usr/bin/env python
"""Aggregate an ncu report's per-SASS-instruction counters by CUDA source line.

    python scripts/ncu_lines.py <report.ncu-rep> <lib.so> <mangled-kernel-substring> [top]

ncu's CSV source page carries no line column, so the SASS listing of the same cubin
(nvdisasm -g) is zipped with it in instruction order."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, lib, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp,
                      stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True,
                     text=True).stdout.splitlines()
# locate the function
start = None
for i, l in enumerate(dis):
    if l.startswith(".text.") and kname in l:
        start = i
        break
assert start is not None, "kernel not found in cubin"
lines = []  # (line_no, sass)
cur = None
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith("\t.section"):
        break
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)), m.group(1))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        lines.append((cur, m.group(2).strip()))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# first kernel instance only
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hdr_idx[0]]
end = hdr_idx[1] - 1 if len(hdr_idx) > 1 else len(rows)
body = [r for r in rows[hdr_idx[0] + 1:end] if len(r) == len(h)]
ci, si = h.index("Instructions Executed"), h.index("# Samples")
print(f"sass instrs: ncu={len(body)} nvdisasm={len(lines)}")
n = min(len(body), len(lines))
agg = {}
srcs = {}
def src_line(key):
    if not key: return "?"
    b, ln, path = key
    if path not in srcs:
        try: srcs[path] = open(path).read().splitlines()
        except OSError: srcs[path] = []
    t = srcs[path]
    return t[ln - 1].strip()[:90] if ln <= len(t) else "?"
tot_i = tot_s = 0
for k in range(n):
    ln = lines[k][0]
    ie, sa = int(body[k][ci] or 0), int(body[k][si] or 0)
    a = agg.setdefault(ln, [0, 0, 0])
    a[0] += ie; a[1] += sa; a[2] += 1
    tot_i += ie; tot_s += sa
print(f"total warp-instructions {tot_i}, samples {tot_s}")
for ln, (ie, sa, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    ln_key = ln
    text = src_line(ln)
    print(f"{(ln[0][:14] if ln else '?'):14s}:{(ln[1] if ln else 0):5d} inst={100 * ie / tot_i:5.1f}% samp={100 * sa / max(1, tot_s):5.1f}% sass={cnt:4d}  {text}")
