#!/usr/bin/env python
"""cProfile of runner.run_experiments for the reference's figure_2_3_4 sweep (10 tuples, L=100, 100001 iterations)."""
import cProfile, pstats, os, sys, tempfile, time, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spgg_b200
from spgg_b200 import runner
eng = spgg_b200.Engine(dict(L=16), seeds=0); eng.close()      # CUDA context up front
tmp = tempfile.mkdtemp()
combos = [(r, k, False, 0.8, 1.0, 1.0, "reputation") for r in (3.0, 3.6, 4.0, 4.5, 5.0) for k in (0.0, 1.0)]
pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable()
res = runner.run_experiments(combos, num_processes=1, use_progress_bar=False, base_dir=os.path.join(tmp, "sweep"), seed=3)
pr.disable(); dt = time.perf_counter() - t0
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18)
print("sweep", round(dt, 3), "s")
for l in s.getvalue().splitlines()[4:32]:
    print(l.replace(ROOT, "")[:170])
