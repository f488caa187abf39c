"""Times (and, under ncu, exposes) the reference-precision iteration: one fp64 lattice, C4 physics.
usage: python scripts/prof_fp64.py [L] [n_iter] [precision]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

L = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
prec = sys.argv[3] if len(sys.argv) > 3 else "fp64"
extra = json.loads(sys.argv[4]) if len(sys.argv) > 4 else {}
us, desc = bench.measure_iteration(dict(bench.C4, L=L, **extra), prec, n_warm=2, n_iter=n)
bps = bench.BYTES_PER_SITE["fp64" if prec == "fp64" else "fp32"]
print(json.dumps({"L": L, "precision": prec, "us_per_iteration": us,
                  "GBps": bps * L * L / (us * 1e-6) / 1e9, "path": desc}))
