#!/usr/bin/env python
"""Per-phase cycle breakdown of the resident kernel (needs build/lib_restrace.so, -DSPGG_RES_TRACE)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["SPGG_B200_LIB"] = os.path.join(ROOT, "build", "lib_restrace.so")
os.environ["SPGG_RES_TRACE_PRINT"] = "1"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import spgg_b200
from helpers import C1, C2, full_params
n = 2000
CASES = (("L64", dict(C1, L=64)), ("L100", dict(C1, L=100)), ("L200", dict(C1, L=200)), ("L200_act_m2", dict(C2, L=200)),
         ("L512_grid", dict(C1, L=512)), ("L1000_grid", dict(C1, L=1000)))
only = sys.argv[1:]
for name, p in CASES:
    if only and name not in only:
        continue
    for cs8 in (("",) if "grid" in name else ("", "1")):
        if cs8: os.environ["SPGG_RES_CS8"] = "1"
        else: os.environ.pop("SPGG_RES_CS8", None)
        eng = spgg_b200.Engine(full_params(p), seeds=1, precision="fp32")
        eng.init_random(7)
        eng.step(n); eng.sync()
        print(f"== {name} cs8={bool(cs8)}: cycles per iteration and phase (phase1, phase2+blockmax, exchange+barrierA, phase3, "
              "reductions+pushes, barrierB, fold) for each CTA:", file=sys.stderr, flush=True)
        eng.step(1); eng.sync()     # prints the trace of the previous (n-iteration) launch
        eng.close()
