#!/usr/bin/env python
"""Hot-spot view of an ncu report's SASS page (first kernel instance): opcode histogram by
executed warp-instructions, and contiguous regions ranked by executed instructions.
    python scripts/ncu_hot.py <report.ncu-rep> [min_count_for_region]"""
import csv, io, subprocess, sys, collections
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hi[0]]
end = hi[1] - 1 if len(hi) > 1 else len(rows)
body = [r for r in rows[hi[0] + 1:end] if len(r) == len(h)]
ie, ss, sm = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
tot = sum(int(r[ie]) for r in body)
tots = sum(int(r[sm]) for r in body)
print(f"{len(body)} SASS instructions, {tot} warp-instructions executed, {tots} samples")
ops = collections.Counter(); smp = collections.Counter()
for r in body:
    t = r[ss].split()
    op = t[1] if t[0].startswith("@") else t[0]
    op = op.split(".")[0]
    ops[op] += int(r[ie]); smp[op] += int(r[sm])
for op, n in ops.most_common(28):
    print(f"  {op:10s} {100*n/tot:5.1f}% inst  {100*smp[op]/max(1,tots):5.1f}% samples")
# execution-count bands: group consecutive instructions with similar exec count
print("regions (consecutive SASS with the same execution count):")
i = 0
regs = []
while i < len(body):
    c = int(body[i][ie]); j = i
    while j < len(body) and int(body[j][ie]) == c: j += 1
    regs.append((i, j, c, sum(int(body[k][sm]) for k in range(i, j))))
    i = j
for (i, j, c, s) in sorted(regs, key=lambda t: -(t[1]-t[0])*t[2])[:14]:
    print(f"  sass[{i:5d}:{j:5d}] n={j-i:4d} exec/inst={c:9d} share={100*(j-i)*c/tot:5.1f}% samples={100*s/max(1,tots):5.1f}%  first: {body[i][ss].strip()[:60]}")
