#!/usr/bin/env python
"""Static size of the innermost loop that contains the Q stores (STG...128) of a kernel.
usage: sass_loop.py lib.so kernel-substring"""
import os, re, subprocess, sys, tempfile
lib, kname = sys.argv[1:3]
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kname in l)
ins = []   # (idx, text)
labels = {}
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith("\t.section"):
        break
    m = re.match(r"^(\.L_x_\d+):", l)
    if m:
        labels[m.group(1)] = len(ins)
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append(m.group(2).strip())
stores = [i for i, t in enumerate(ins) if re.search(r"STG\.E(\.\w+)*\.128", t)]
loops = []
for i, t in enumerate(ins):
    m = re.search(r"BRA\S*\s+.*`\((\.L_x_\d+)\)", t)
    if m and m.group(1) in labels and labels[m.group(1)] <= i:
        loops.append((labels[m.group(1)], i))
best = None
for a, b in loops:
    if stores and all(a <= s <= b for s in stores[:4]):
        if best is None or (b - a) < (best[1] - best[0]):
            best = (a, b)
print("total", len(ins), "stores at", stores[:8])
if best:
    a, b = best
    body = ins[a:b + 1]
    print("loop", a, b, "len", len(body))
    from collections import Counter
    c = Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for t in body)
    print(c.most_common(25))
