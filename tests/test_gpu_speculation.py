"""The TMA fast path runs ONE launch per iteration: it assumes the lattice-global maximum |reward
difference| (spgg.py:488) of the previous iteration, computes the true one as a by-product of the
update and compares; a wrong guess is re-run (include/spgg.h, spgg_step).  Results must not depend
on any of this: bit-identical (S, R, Q and every statistic row) to the exact two-launch iteration
(SPGG_NO_SPEC=1), with guesses that fail on their own and with guesses spoiled on purpose."""
import subprocess
import sys
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_CHILD = r"""
import sys, json, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import spgg_b200
from helpers import C1, C2, full_params
cfg, L, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
p = full_params(dict(C1 if cfg == "c1" else C2, L=L))
n_rep = 3 if cfg == "c1b" else 1
if cfg == "c1b":
    p = [full_params(dict(C1, L=L, r=r)) for r in (3.0, 4.0, 5.0)]
rs = np.random.RandomState(5)
eng = spgg_b200.Engine(p, seeds=list(range(40, 40 + n_rep)), precision="fp32")
assert eng.describe().startswith("fast"), eng.describe()
for r in range(n_rep):
    eng.set_state(rs.randint(0, 2, (L, L)), np.zeros((L, L)), rs.uniform(-0.01, 0.01, (L, L, 2, 2)), replica=r)
rows = []
for n in (60, 1, 90, 149):
    eng.step(n)
    rows.append(np.stack([eng.stats(r)[1:] for r in range(n_rep)]))
st = eng.status()
S, R, Q = zip(*[eng.get_state(r) for r in range(n_rep)])
np.savez(out, S=np.stack(S), R=np.stack(R), Q=np.stack(Q), rows=np.concatenate(rows, axis=1),
         spec=np.array([st.speculative_launches, st.speculation_failures]), it=np.array(st.iteration))
"""


def _run(tmp_path, tag, cfg, L, env):
    out = str(tmp_path / f"{tag}.npz")
    e = dict(os.environ, SPGG_NO_RESIDENT="1", **env)
    r = subprocess.run([sys.executable, "-c", _CHILD.format(root=ROOT), cfg, str(L), out], env=e,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return np.load(out)


@pytest.mark.parametrize("cfg,L", [("c1", 256), ("c2", 256), ("c1", 1024), ("c1b", 128)])
def test_speculative_iteration_equals_exact_pair(tmp_path, cfg, L):
    exact = _run(tmp_path, "exact", cfg, L, {"SPGG_NO_SPEC": "1"})
    spec = _run(tmp_path, "spec", cfg, L, {})
    poisoned = _run(tmp_path, "poison", cfg, L, {"SPGG_SPEC_TEST_POISON": "37"})
    assert exact["spec"][0] == 0
    assert spec["spec"][0] > 0                              # the one-launch path really ran
    assert poisoned["spec"][1] >= 3                         # ... and so did the re-run after a wrong guess
    for other in (spec, poisoned):
        assert int(other["it"]) == int(exact["it"]) == 300
        for k in ("S", "R", "Q", "rows"):
            assert np.array_equal(exact[k], other[k]), k
