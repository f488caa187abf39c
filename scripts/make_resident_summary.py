#!/usr/bin/env python
"""profiles/<round>_k_resident_<tag>.md from one gpurun visit (scripts/gpu_final.sh): launch list of
`bench.py --workload sweep`, `ncu --set full` metrics of k_resident, SASS evidence of the cluster
barrier / DSMEM / dp4a / redux instructions.   python scripts/make_resident_summary.py <tag> [round]"""
import collections, csv, io, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]; rnd = sys.argv[2] if len(sys.argv) > 2 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
rep = os.path.join(G, f"prof_res_{tag}.ncu-rep")
L = [f"# ncu evidence for the lattice-resident cluster kernel, round {rnd[1:]}, capture `{tag}`", "",
     "Command profiled: `python bench.py --workload sweep --steps 1 --warmup 1 --inner 300` (BASELINE config 3: "
     "60 replicas of L=200 as two batched handles of 30, one thread-block cluster of 8 CTAs per replica) on one "
     "B200; the program was first run to exit 0 without ncu.", ""]
lcsv = os.path.join(G, f"launches_res_{tag}.csv")
if os.path.exists(lcsv):
    shutil.copy(lcsv, os.path.join(P, f"{rnd}_launches_res_{tag}.csv"))
    rows = list(csv.reader(open(lcsv)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    d = collections.defaultdict(list)
    for r in rows[hi + 1:]:
        if len(r) == len(h):
            d[r[ki]].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    L += ["## Launch list (`--metrics gpu__time_duration.sum --clock-control none`)", "",
          "| kernel | launches | mean us | share of GPU time |", "|---|---|---|---|"]
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        L.append(f"| `{k.split('(')[0][:60]}` | {len(v)} | {sum(v)/len(v)/1e3:.1f} | {100*sum(v)/tot:.1f}% |")
    L += ["", "One `k_resident` launch = one whole chunk (300 iterations) of 30 replicas; nothing else runs inside the "
          "timed region.", ""]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, r0 = rows[0], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__cluster_max_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_active.avg"]
L += ["## `ncu --set full --clock-control none --import-source on -k regex:k_resident`, one launch", "",
      "| metric | value |", "|---|---|"]
units = rows[1]
for w in want:
    if w in h:
        i = h.index(w)
        L.append(f"| {w} | {r0[i][:90]} {units[i]} |")
L += ["", "DRAM traffic of the whole 300-iteration launch is the initial load and final write-back of the 30 "
      "lattices plus one statistics row per iteration: the state never leaves the chip in between.", ""]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]; ie, ss = h.index("Instructions Executed"), h.index("Source")
ops = collections.Counter(); n_sass = 0
for r in rows[hi + 1:]:
    if len(r) != len(h):
        continue
    n_sass += 1
    t = r[ss].split(); op = t[1] if t[0].startswith("@") else t[0]
    ops[op] += int(r[ie])
tot = sum(ops.values())
L += ["## SASS of the captured kernel", "", f"{n_sass} SASS instructions, {tot} warp-instructions executed.", "",
      "| what | mnemonic | executed warp-instructions |", "|---|---|---|"]
for what, m in (("cluster barrier arrive / wait", ["UCGABAR_ARV", "UCGABAR_WAIT"]),
                ("DSMEM stores (ghost rows, maxima, partial rows: generic stores into the peer's window)", ["ST.E", "ST.E.64"]),
                ("byte-parallel reputation sums", ["IDP.4A.S8.S8"]),
                ("warp reductions of the integer statistics / block max", ["REDUX.SUM", "CREDUX.MAX"]),
                ("shared atomics (block accumulators)", ["ATOMS.ADD", "ATOMS.MAX"]),
                ("block barriers", ["BAR.SYNC.DEFER_BLOCKING"]),
                ("global loads / stores (threshold word, statistics row, chunk load / write-back)",
                 ["LDG.E", "LDG.E.CONSTANT", "LDG.E.128", "STG.E.64", "STG.E.128"])):
    L.append(f"| {what} | " + ", ".join(f"`{x}`" for x in m) + " | " + ", ".join(str(ops.get(x, 0)) for x in m) + " |")
L += ["", "| opcode | share of executed instructions |", "|---|---|"]
grp = collections.Counter()
for k, v in ops.items():
    grp[k.split(".")[0]] += v
for k, v in grp.most_common(14):
    L.append(f"| {k} | {100*v/tot:.1f}% |")
open(os.path.join(P, f"{rnd}_k_resident_{tag}.md"), "w").write("\n".join(L) + "\n")
print("wrote", os.path.join(P, f"{rnd}_k_resident_{tag}.md"))
