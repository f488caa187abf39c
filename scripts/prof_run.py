#!/usr/bin/env python
"""cProfile of SPGG.run (host side) for C1/C2-shaped runs."""
import cProfile, pstats, os, sys, tempfile, time, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import spgg_b200
RUN = dict(c=1, cost=1, gamma=0.9, epsilon=0.5, epsilon_decay=0.99, epsilon_min=0.01, lambda_epsilon=0.01,
           delta_R_C=1, delta_R_D=1, R_min=-10, R_max=10, alpha=0.8)
tmp = tempfile.mkdtemp()
def c2():
    m = spgg_b200.SPGG(**RUN, r=4.0, influence_factor=1.0, use_second_order=True, reward_weight_payoff=1.0,
                       rep_gain_C=1.0, state_representation="action", L=200, iterations=10000, seed=2)
    m.folder = tmp
    return m.run(os.path.join(tmp, "c2.h5"))
def c1():
    m = spgg_b200.SPGG(**RUN, r=3.0, influence_factor=1.0, use_second_order=False, reward_weight_payoff=0.95,
                       rep_gain_C=1.0, L=100, iterations=100001, seed=1)
    m.folder = tmp
    return m.run(os.path.join(tmp, "c1.h5"))
for name, fn in (("c2", c2), ("c2_again", c2), ("c1", c1)):
    pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable(); fn(); pr.disable(); dt = time.perf_counter() - t0
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22)
    print("=====", name, round(dt, 3), "s"); print("\n".join(s.getvalue().splitlines()[4:40]))
