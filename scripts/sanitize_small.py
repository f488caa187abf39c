#!/usr/bin/env python
"""Small resident / fast / general runs for compute-sanitizer (memcheck, racecheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import spgg_b200
from helpers import C1, C2, full_params
which = sys.argv[1] if len(sys.argv) > 1 else "resident"
cases = {"resident": [dict(C1, L=36, use_second_order=True), dict(C2, L=10), dict(C2, L=102, use_second_order=False),
                      dict(C1, L=64), dict(C1, L=40, use_second_order=True)],
         "fast": [dict(C1, L=256), dict(C2, L=128)],
         "general": [dict(C1, L=50), dict(C2, L=70)]}[which]
if which != "resident":
    os.environ["SPGG_NO_RESIDENT"] = "1"
for p in cases:
    p = full_params(p)
    for nrep in (1, 5):
        eng = spgg_b200.Engine([p] * nrep, seeds=list(range(nrep)), precision="fp32")
        for r in range(nrep):
            eng.init_random(10 + r, r)
        eng.step(6); eng.step(3)
        S, R, Q = eng.get_state(nrep - 1)
        rows = eng.stats(nrep - 1)
        print(which, "L", p["L"], "replicas", nrep, "coop", float((S == 0).mean()), "rows", rows.shape, flush=True)
        eng.close()
print("done")
