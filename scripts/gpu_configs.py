#!/usr/bin/env python
"""Wall time of the BASELINE.json configurations through the public API (SPGG.run /
run_experiments), HDF5 output included.  Prints one JSON object."""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import spgg_b200
from spgg_b200 import runner

out = {}
tmp = tempfile.mkdtemp()
RUN = dict(c=1, cost=1, gamma=0.9, epsilon=0.5, epsilon_decay=0.99, epsilon_min=0.01, lambda_epsilon=0.01,
           delta_R_C=1, delta_R_D=1, R_min=-10, R_max=10, alpha=0.8)
# C1: default_config.yaml run: reputation state, M=1, r=3, kappa=1, L=100, 100001 iterations
m = spgg_b200.SPGG(**RUN, r=3.0, influence_factor=1.0, use_second_order=False, reward_weight_payoff=0.95,
                   rep_gain_C=1.0, L=100, iterations=100001, seed=1)
m.folder = tmp
t0 = time.perf_counter(); ret = m.run(os.path.join(tmp, "c1.h5")); dt = time.perf_counter() - t0
out["C1 L=100 reputation M=1 r=3 100001 iterations"] = {"seconds": dt, "site_updates_per_s": 100 * 100 * 100001 / dt,
                                                         "final_coop": ret[0], "launches": m.kernel_launches}
# C2: action state, M=2, r=4, kappa=1, L=200, 10^4 steps
m = spgg_b200.SPGG(**RUN, r=4.0, influence_factor=1.0, use_second_order=True, reward_weight_payoff=1.0,
                   rep_gain_C=1.0, state_representation="action", L=200, iterations=10000, seed=2)
m.folder = tmp
t0 = time.perf_counter(); ret = m.run(os.path.join(tmp, "c2.h5")); dt = time.perf_counter() - t0
out["C2 L=200 action M=2 r=4 10^4 iterations"] = {"seconds": dt, "site_updates_per_s": 200 * 200 * 10000 / dt,
                                                   "final_coop": ret[0]}
# C3: the reference's figure_2_3_4 sweep as the runner runs it (L=100, 100001 iterations, 10 tuples), batched
combos = [(r, k, False, 0.8, 1.0, 1.0, "reputation") for r in (3.0, 3.6, 4.0, 4.5, 5.0) for k in (0.0, 1.0)]
t0 = time.perf_counter()
res = runner.run_experiments(combos, num_processes=1, use_progress_bar=False, base_dir=os.path.join(tmp, "sweep"), seed=3)
dt = time.perf_counter() - t0
out["C3 sweep 10 tuples x (L=100, 100001 iterations), one GPU"] = {
    "seconds": dt, "site_updates_per_s": 10 * 100 * 100 * 100001 / dt, "final_coop": [round(c, 3) for _p, (c, _r) in res]}
# C4 at its full length: L=4096, 10^4 iterations through the class (HDF5 of 16.7M-site lattices included)
m = spgg_b200.SPGG(**RUN, r=3.0, influence_factor=1.0, use_second_order=False, reward_weight_payoff=0.95,
                   rep_gain_C=1.0, L=4096, iterations=10000, seed=4)
m.folder = tmp
t0 = time.perf_counter(); ret = m.run(os.path.join(tmp, "c4.h5")); dt = time.perf_counter() - t0
out["C4 L=4096 reputation M=1 10^4 iterations (SPGG.run incl. snapshots + HDF5)"] = {
    "seconds": dt, "site_updates_per_s": 4096 * 4096 * 10000 / dt, "final_coop": ret[0]}
print(json.dumps(out, indent=1))
