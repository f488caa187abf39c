"""TEST INFRASTRUCTURE ONLY - never imported by the product path.

Runs the *unmodified* reference (``/root/reference``) inside this container so
that (a) the oracle restatements in this directory can be pinned against it and
(b) golden input/output vectors can be generated for ``tests/golden/``.

The reference cannot be imported as shipped because ``h5py`` and ``matplotlib``
are not installed (reference ``src/model/spgg.py:6-10``); two in-memory stub
modules are registered in ``sys.modules`` first.  ``/root/reference`` does not
exist on the GPU box, so nothing that runs there may import this file; only the
golden generator (``oracle/make_golden.py``) and the "reference available"
CPU tests do.

What is recorded for a run
--------------------------
* the ctor draws (Q table, initial strategies)  - ``spgg.py:121,162``
* for every step the two draw arrays consumed by ``QLearning.select_action``
  (``algorithms.py:105,108``): ``u = rand(L,L)`` then ``b = randint(0,2,(L,L))``
* every dataset the run writes through ``h5py.File.create_dataset``
* final ``q_table``, ``R``, ``_Sn``
"""
from __future__ import annotations

import contextlib
import os
import sys
import tempfile
import types

import numpy as np

def _find_reference() -> str:
    # the live tree in the build container, else the copy oracle/fetch_ref.py shipped (GPU box)
    live = os.environ.get("SPGG_REFERENCE_ROOT", "/root/reference")
    shipped = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
    for root in (live, shipped):
        if os.path.isfile(os.path.join(root, "src", "model", "spgg.py")):
            return root
    return live


REFERENCE_ROOT = _find_reference()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "model", "spgg.py"))


class _Anything:
    """Permissive stand-in for matplotlib objects: callable, attribute-returning,
    unpackable into two (``fig, ax = plt.subplots()``)."""

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()

    def __iter__(self):
        return iter((_Anything(), _Anything()))


class _RecordingFile:
    """h5py.File stand-in: keeps datasets in a dict shared per filename."""

    store: dict = {}

    def __init__(self, name, mode="r"):
        self.name = name
        if "w" in mode:
            _RecordingFile.store[name] = {}
        self.d = _RecordingFile.store.setdefault(name, {})

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def create_dataset(self, key, data=None, **kw):
        self.d[key] = np.array(data)
        return self.d[key]

    def __contains__(self, key):
        return key in self.d

    def __getitem__(self, key):
        return self.d[key]

    def keys(self):
        return self.d.keys()


def install_stubs() -> None:
    if "h5py" not in sys.modules:
        h5 = types.ModuleType("h5py")
        h5.File = _RecordingFile
        sys.modules["h5py"] = h5
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        col = types.ModuleType("matplotlib.colors")
        anything = _Anything()
        for mod in (mpl, plt, col):
            mod.__getattr__ = lambda name, _a=anything: _a  # PEP 562 module getattr
        mpl.pyplot = plt
        mpl.colors = col
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
        sys.modules["matplotlib.colors"] = col


def import_reference():
    """Return the reference's ``src.model`` package (stubs installed first)."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import src.model as ref_model  # noqa: E402

    return ref_model


@contextlib.contextmanager
def pinned_seed(seed: int):
    """The ctor calls ``np.random.seed()`` with no argument (``spgg.py:98``);
    make that call seed a fixed value instead."""
    real = np.random.seed
    np.random.seed = lambda *a, **k: real(seed)
    try:
        yield
    finally:
        np.random.seed = real


@contextlib.contextmanager
def recorded_draws(log: dict):
    """Record every ``np.random.rand`` / ``np.random.randint`` array."""
    real_rand, real_randint = np.random.rand, np.random.randint
    log.setdefault("rand", [])
    log.setdefault("randint", [])

    def rand(*shape):
        out = real_rand(*shape)
        log["rand"].append(np.array(out))
        return out

    def randint(*a, **k):
        out = real_randint(*a, **k)
        log["randint"].append(np.array(out))
        return out

    np.random.rand, np.random.randint = rand, randint
    try:
        yield log
    finally:
        np.random.rand, np.random.randint = real_rand, real_randint


def neutralise_cluster_tail():
    """``spgg.py:631-633`` is O(#clusters * L^2); replace ``label`` for large L.
    Only the ``cluster_sizes`` dataset changes."""
    import src.model.spgg as ref_spgg

    ref_spgg.label = lambda a: (np.zeros(a.shape, dtype=int), 0)


def run_reference(seed: int, cluster_tail: bool = True, **params):
    """Construct + run the reference with a pinned seed.

    Returns a dict with the ctor state, the recorded per-step draws, every
    dataset written, and the final state.
    """
    ref_model = import_reference()
    if not cluster_tail:
        neutralise_cluster_tail()
    log: dict = {}
    with pinned_seed(seed):
        model = ref_model.SPGG(**params)
    q0 = model.q_table.copy()
    dq_tables0 = None
    if getattr(model.algorithm, "q_table_1", None) is not None:   # DoubleQLearning (algorithms.py:245-260)
        dq_tables0 = (model.algorithm.q_table_1.copy(), model.algorithm.q_table_2.copy())
    s0 = model._Sn.copy()
    r0 = model.R.copy()
    tmp = tempfile.mkdtemp(prefix="spgg_ref_")
    model.folder = tmp
    fname = os.path.join(tmp, "run.h5")
    with recorded_draws(log):
        ret = model.run(fname)
    datasets = dict(_RecordingFile.store.pop(fname))
    n_steps = len(datasets["epsilon_history_final"])
    algo = type(model.algorithm).__name__
    out = {
        "params": dict(params),
        "algorithm": algo,
        "q0": q0,
        "s0": s0,
        "r0": r0,
        "rand": log["rand"],
        "randint": log["randint"],
        "datasets": datasets,
        "q_final": model.q_table.copy(),
        "r_final": model.R.copy(),
        "s_final": model._Sn.copy(),
        "ret": ret,
        "n_steps": n_steps,
    }
    if dq_tables0 is not None:
        out["q1_0"], out["q2_0"] = dq_tables0
        out["q1_final"] = model.algorithm.q_table_1.copy()
        out["q2_final"] = model.algorithm.q_table_2.copy()
    if algo == "QLearning":
        # one rand + one randint per completed step (algorithms.py:105,108)
        assert len(log["rand"]) == n_steps and len(log["randint"]) == n_steps
        L = params["L"]
        out["u"] = (np.stack(log["rand"]) if n_steps else np.zeros((0, L, L)))
        out["b"] = (np.stack(log["randint"]).astype(np.uint8) if n_steps
                    else np.zeros((0, L, L), np.uint8))
    return out


def time_reference(n_warm: int, n_timed: int, seed: int = 0, **params):
    """Wall time per iteration of the UNMODIFIED reference loop (``spgg.py:368-592``) on this
    host: one ``SPGG.run`` of ``n_warm + n_timed`` iterations; the boundaries between iterations
    are observed through the one public hook the loop calls exactly once per iteration,
    ``algorithm.decay_epsilon`` (``spgg.py:548-550``), so each interval covers one whole loop
    body.  ``label`` is neutralised for L > 1000 (the O(#clusters L^2)
    tail ``spgg.py:631-633`` is not part of the step).  Returns (seconds per timed iteration,
    list of the timed intervals)."""
    import time

    ref_model = import_reference()
    if int(params.get("L", 50)) > 1000:
        neutralise_cluster_tail()
    params = dict(params, iterations=int(n_warm + n_timed))
    with pinned_seed(seed):
        model = ref_model.SPGG(**params)
    tmp = tempfile.mkdtemp(prefix="spgg_ref_")
    model.folder = tmp
    stamps = []
    real_decay = model.algorithm.decay_epsilon

    def decay_and_stamp(*a, **k):
        out = real_decay(*a, **k)
        stamps.append(time.perf_counter())
        return out

    model.algorithm.decay_epsilon = decay_and_stamp
    t_start = time.perf_counter()
    fname = os.path.join(tmp, "run.h5")
    model.run(fname)
    _RecordingFile.store.pop(fname, None)
    stamps = [t_start] + stamps
    iv = [b - a for a, b in zip(stamps[:-1], stamps[1:])]
    timed = iv[n_warm:n_warm + n_timed]
    if len(timed) < n_timed:      # the lattice became uniform (spgg.py:405): fewer iterations ran
        timed = iv[-max(1, len(iv) - n_warm):]
    return sum(timed) / len(timed), timed
