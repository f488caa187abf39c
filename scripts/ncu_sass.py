#!/usr/bin/env python
"""Dump SASS of the first kernel instance in an ncu report with executed counts and source lines.
usage: ncu_sass.py report.ncu-rep lib.so kernel-substring > out.txt"""
import csv, io, os, re, subprocess, sys, tempfile
rep, lib, kname = sys.argv[1:4]
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kname in l)
lines = []
cur = (None, None)
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith("\t.section"):
        break
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        lines.append((cur, m.group(2).strip()))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hdr_idx[0]]
end = hdr_idx[1] - 1 if len(hdr_idx) > 1 else len(rows)
body = [r for r in rows[hdr_idx[0] + 1:end] if len(r) == len(h)]
ci, si = h.index("Instructions Executed"), h.index("# Samples")
for k in range(min(len(body), len(lines))):
    (f, ln), sass = lines[k]
    print(f"{k:5d} {int(body[k][ci] or 0):10d} {int(body[k][si] or 0):5d}  {f}:{ln}  {sass}")
