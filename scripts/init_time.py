import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
t0 = time.perf_counter()
import spgg_b200
from helpers import C1, full_params
t1 = time.perf_counter()
L = int(sys.argv[1]) if len(sys.argv) > 1 else 100
eng = spgg_b200.Engine(full_params(dict(C1, L=L)), seeds=1, precision="fp32")
t2 = time.perf_counter()
eng.init_random(3); eng.step(10); eng.sync()
t3 = time.perf_counter()
eng2 = spgg_b200.Engine(full_params(dict(C1, L=L)), seeds=1, precision="fp32")
t4 = time.perf_counter()
print(f"L={L} NO_RESIDENT={os.environ.get('SPGG_NO_RESIDENT')} MODULE_LOADING={os.environ.get('CUDA_MODULE_LOADING')}: import {t1-t0:.2f} s, first create {t2-t1:.2f} s, first steps {t3-t2:.2f} s, second create {t4-t3:.3f} s")
