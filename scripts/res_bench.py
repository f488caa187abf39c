#!/usr/bin/env python
"""Device time per iteration of small lattices: resident cluster kernel vs per-iteration kernels."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import spgg_b200
from helpers import C1, C2, full_params

def run(p, nrep, n, reps=3):
    L = p["L"]
    eng = spgg_b200.Engine([p] * nrep, seeds=list(range(nrep)), precision="fp32")
    for r in range(nrep):
        eng.init_random(100 + r, r)
    eng.step(50); eng.sync()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); eng.step(n); eng.sync(); best = min(best, time.perf_counter() - t0)
    it = eng.status().iteration
    eng.close()
    return best / n * 1e6, nrep * L * L * n / best

only = sys.argv[1:] 
out = {}
cases = [("c1_L100", dict(C1, L=100), 1), ("c2_L200_act_m2", dict(C2, L=200), 1), ("c1_L200", dict(C1, L=200), 1),
         ("c1_L100x10", dict(C1, L=100), 10), ("c1_L200x18", dict(C1, L=200), 18), ("c1_L200x60", dict(C1, L=200), 60),
         ("c1_L240", dict(C1, L=240), 1), ("c1_L64", dict(C1, L=64), 1),
         # grid mode (cooperative launch over every SM)
         ("c1_L400", dict(C1, L=400), 1), ("c1_L512", dict(C1, L=512), 1), ("c1_L1000", dict(C1, L=1000), 1),
         ("c2_L1000_act_m2", dict(C2, L=1000), 1), ("c1_L1024", dict(C1, L=1024), 1)]
for name, p, nrep in cases:
    if only and name not in only: continue
    p = full_params(p)
    for mode in ("resident", "periter"):
        if mode == "periter": os.environ["SPGG_NO_RESIDENT"] = "1"
        else: os.environ.pop("SPGG_NO_RESIDENT", None)
        n = 2000 if mode == "resident" else 500
        us, rate = run(p, nrep, n)
        out[f"{name}:{mode}"] = {"us_per_iteration": round(us, 3), "site_updates_per_s": rate}
        print(name, mode, round(us, 3), "us/iter", f"{rate:.3e}", flush=True)
print(json.dumps(out))
