#!/usr/bin/env python
"""cProfile of SPGG.run for the C4 configuration (L=4096, 10^4 iterations) - host side."""
import cProfile, pstats, os, sys, tempfile, time, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import spgg_b200
RUN = dict(c=1, cost=1, gamma=0.9, epsilon=0.5, epsilon_decay=0.99, epsilon_min=0.01, lambda_epsilon=0.01,
           delta_R_C=1, delta_R_D=1, R_min=-10, R_max=10, alpha=0.8)
tmp = tempfile.mkdtemp()
L = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
t0 = time.perf_counter()
m = spgg_b200.SPGG(**RUN, r=3.0, influence_factor=1.0, use_second_order=False, reward_weight_payoff=0.95,
                   rep_gain_C=1.0, L=L, iterations=T, seed=4)
m.folder = tmp
print("ctor", round(time.perf_counter() - t0, 3), "s")
pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable(); m.run(os.path.join(tmp, "c4.h5")); pr.disable(); dt = time.perf_counter() - t0
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(25)
print("run", round(dt, 3), "s")
import re
print("\n".join(re.sub(r"/\S*/", "", l)[:150] for l in s.getvalue().splitlines()[4:45]))
