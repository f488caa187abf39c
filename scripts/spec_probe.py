"""Debug helper: one configuration of tests/test_gpu_speculation.py in-process, with progress prints."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests")
import spgg_b200
from helpers import C1, C2, full_params
cfg, L = sys.argv[1], int(sys.argv[2])
p = full_params(dict(C1 if cfg == "c1" else C2, L=L))
n_rep = 3 if cfg == "c1b" else 1
if cfg == "c1b":
    p = [full_params(dict(C1, L=L, r=r)) for r in (3.0, 4.0, 5.0)]
rs = np.random.RandomState(5)
eng = spgg_b200.Engine(p, seeds=list(range(40, 40 + n_rep)), precision="fp32")
print(eng.describe(), flush=True)
for r in range(n_rep):
    eng.set_state(rs.randint(0, 2, (L, L)), np.zeros((L, L)), rs.uniform(-0.01, 0.01, (L, L, 2, 2)), replica=r)
for n in (60, 1, 90, 149):
    t = time.time()
    eng.step(n)
    eng.sync()
    st = eng.status()
    print(f"chunk {n}: {time.time()-t:.3f}s it={st.iteration} spec={st.speculative_launches} fail={st.speculation_failures} launches={st.kernel_launches}", flush=True)
for r in range(n_rep):
    S, R, Q = eng.get_state(r)
    print("replica", r, "coop", (S == 0).mean(), "digest", eng.digest(r), eng.status(r).stopped_at, flush=True)
