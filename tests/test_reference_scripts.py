"""north_star: "scripts/run_experiments.py and plot_figures.py work unchanged".

The UNMODIFIED reference tree (``/root/reference`` in the build container, the copy
``oracle/fetch_ref.py`` ships as ``oracle/_ref`` on the GPU box) is executed against the drop-in:
``spgg_b200.dropin.install`` rebinds ``src.model.SPGG`` in the interpreter, nothing in the
reference is edited.  matplotlib is absent from the image, so the plotting calls go to the
permissive stand-in of ``oracle/ref_harness.py``; everything up to them - argument parsing,
config loading, the fork-based runner, folder layout, HDF5 writing, ``plotting.load_data`` reading
the files back by dataset name - is the reference's own code.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fetch_ref  # noqa: E402

REF = fetch_ref.reference_root()
needs_ref = pytest.mark.skipif(REF is None, reason="no copy of the reference (oracle/_ref) on this box")

# datasets plot_figures.py / plot_state_comparison.py read through load_data (plotting.py)
PLOTTED = ["coop_rate_history", "neighbor_influence_percent", "rep_avg_history_final",
           "switch_C_to_D", "switch_D_to_C", "avg_q_s0_c_history", "cooperators_q_s1_d_history"]

_CHILD = r"""
import os, sys, json, types
sys.path.insert(0, {root!r})
import spgg_b200
from spgg_b200 import dropin
from oracle import ref_harness
# matplotlib stand-in only: h5py must be the drop-in's own shim, not the recording stub
sys.modules.setdefault("h5py", __import__("spgg_b200.h5lite", fromlist=["x"]))
ref_harness.install_stubs()
dropin.install({ref!r})
{body}
"""


def _run_child(body, cwd, timeout=900):
    code = _CHILD.format(root=ROOT, ref=REF, body=body)
    return subprocess.run([sys.executable, "-c", code], cwd=cwd, capture_output=True, text=True,
                          timeout=timeout)


@needs_ref
def test_reference_load_data_reads_h5lite_files(tmp_path):
    """plotting.load_data (plotting.py:36-63: ``h5py.File(path, 'r')``, ``name in f``, ``f[name][:]``)
    unchanged, on a file written by the drop-in's HDF5 writer; no GPU needed."""
    from spgg_b200 import h5lite
    path = str(tmp_path / "experiment_data.h5")
    rs = np.random.RandomState(0)
    want = {"coop_rate_history": rs.rand(1000), "Sn_final": rs.randint(0, 2, (50, 50)),
            "switch_C_to_D": rs.randint(0, 99, 1000), "empty": np.zeros(0)}
    want.update({f"extra_{k}": rs.rand(3 + k) for k in range(40)})     # more than one B-tree node's worth
    with h5lite.File(path, "w") as f:
        for k, v in want.items():
            f.create_dataset(k, data=v)
    np.save(tmp_path / "want.npy", np.array(json.dumps({k: v.tolist() for k, v in want.items()})))
    body = f"""
import numpy as np
from src.visualization.plotting import load_data
want = json.loads(str(np.load({str(tmp_path / 'want.npy')!r})))
for k, v in want.items():
    got = load_data({path!r}, k)
    assert got is not None, k
    assert np.array_equal(np.asarray(got), np.asarray(v)), k
assert load_data({path!r}, "not_there") is None
assert load_data({str(tmp_path / 'missing.h5')!r}, "x") is None
print("load_data ok")
"""
    r = _run_child(body, str(tmp_path))
    assert r.returncode == 0 and "load_data ok" in r.stdout, r.stdout + r.stderr


@needs_ref
@pytest.mark.gpu
def test_reference_runner_runs_one_experiment_unchanged(tmp_path):
    """``src.experiments.runner.run_one_experiment`` (runner.py:48-114) as shipped: folder layout,
    SPGG(...) with its hard-coded arguments (L=100, 100 001 iterations), ``.folder``, ``.run(h5)``,
    the 3-tuple, ``rep_avg_history``; then the reference's load_data reads the series back."""
    body = f"""
import numpy as np
from src.experiments.runner import run_one_experiment, get_folder_name
from src.visualization.plotting import load_data
params = (3.0, 1.0, False, 0.8, 0.95, 1.0, 'reputation', 'qlearning')
got_params, (coop, rep_mean) = run_one_experiment(params)
assert got_params == params and 0.0 <= coop <= 1.0 and -10.0 <= rep_mean <= 10.0
folder = get_folder_name(*params)
for sub in ("configurations", "reputations", "plots", os.path.join("plots", "snapshots"), "data"):
    assert os.path.isdir(os.path.join(folder, sub)), sub
h5 = os.path.join(folder, "data", "experiment_data.h5")
out = {{}}
for name in {PLOTTED!r}:
    d = load_data(h5, name)
    assert d is not None and len(d) >= 100000, (name, None if d is None else len(d))
    out[name] = float(np.asarray(d, dtype=float)[-2000:].mean())
fc = load_data(h5, "coop_rate_history")
assert abs(fc[-1] - coop) < 0.05
sn = load_data(h5, "Sn_final")
assert sn.shape == (100, 100) and abs((sn == 0).mean() - coop) < 1e-12
print("RESULT " + json.dumps(dict(coop=coop, rep_mean=rep_mean, tail=out)))
"""
    r = _run_child(body, str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][0][7:])
    # C1 physics: the reference's own long-run cooperation level (tests/golden/band_c1_long.npz:
    # 0.709 +- 0.009 over iterations 8000..10^4, still rising slowly) - a loose sanity band
    assert 0.55 <= res["tail"]["coop_rate_history"] <= 0.95, res


@needs_ref
@pytest.mark.gpu
def test_reference_cli_sweep_and_plot_scripts_unchanged(tmp_path):
    """``scripts/run_experiments.py --experiment-type figure_2_3_4`` (10 tuples, its own
    multiprocessing.Pool with 2 forked workers: CUDA is initialised in the children) and then
    ``scripts/plot_figures.py --figures 2 4`` over the files it wrote (matplotlib stand-in)."""
    body = f"""
import runpy
sys.argv = ["run_experiments.py", "--experiment-type", "figure_2_3_4", "--num-processes", "2", "--no-progress"]
runpy.run_path(os.path.join({REF!r}, "scripts", "run_experiments.py"), run_name="__main__")
folders = sorted(d for d in os.listdir(".") if d.startswith("results_r"))
print("FOLDERS " + json.dumps(folders))
from src.visualization import plotting
calls = []
real = plotting.load_data
def counting(path, name):
    out = real(path, name)
    calls.append((os.path.exists(path), out is not None))
    return out
plotting.load_data = counting
import src.visualization as vis
sys.argv = ["plot_figures.py", "--data-dir", ".", "--output-dir", "paper_figures", "--figures", "2", "4"]
runpy.run_path(os.path.join({REF!r}, "scripts", "plot_figures.py"), run_name="__main__")
print("LOADS " + json.dumps(dict(total=len(calls), found=sum(1 for e, o in calls if e and o))))
"""
    r = _run_child(body, str(tmp_path), timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    folders = json.loads([l for l in r.stdout.splitlines() if l.startswith("FOLDERS ")][0][8:])
    assert len(folders) == 10, folders                       # default_config.yaml: figure_2_3_4 has 10 tuples
    for f in folders:
        assert os.path.getsize(tmp_path / f / "data" / "experiment_data.h5") > 10 ** 6
    loads = json.loads([l for l in r.stdout.splitlines() if l.startswith("LOADS ")][0][6:])
    assert loads["total"] > 0 and loads["found"] == loads["total"], loads
