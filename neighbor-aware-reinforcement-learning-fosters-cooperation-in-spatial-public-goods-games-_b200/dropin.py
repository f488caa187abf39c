"""Run the reference's own scripts against the B200 step without editing them.

The reference's only boundary for the hot path is the class ``src.model.SPGG``
(``src/experiments/runner.py:8`` imports it, ``runner.py:88-105`` constructs and runs it;
``src/visualization/plotting.py:36-63`` later reads the HDF5 datasets by name).  ``install()``
performs, in the running interpreter, the substitution INTEGRATION.md describes as a one-line
edit of ``src/model/__init__.py``:

    import spgg_b200.dropin as dropin
    dropin.install("/path/to/reference")          # before the reference's modules are imported
    runpy.run_path("/path/to/reference/scripts/run_experiments.py", run_name="__main__")

or from a shell:  ``python -m spgg_b200.dropin /path/to/reference scripts/run_experiments.py
--experiment-type figure_2_3_4 --num-processes 2``

What it does: puts the reference root on ``sys.path``; registers the flat-file HDF5 subset
(``h5lite``) under the name ``h5py`` when the real h5py is not importable (the reference does
``import h5py`` at module scope, ``spgg.py:6``, ``plotting.py:6``); imports ``src.model`` and
rebinds ``SPGG`` and the algorithm classes there to this package's.  Nothing is written into the
reference tree.
"""
from __future__ import annotations

import importlib
import os
import runpy
import sys


def install_h5py_shim() -> bool:
    """``import h5py`` -> the built-in flat-file reader/writer, unless h5py is installed.
    Returns True when the shim was registered."""
    try:
        import h5py  # noqa: F401
        return False
    except ImportError:
        from . import h5lite
        sys.modules["h5py"] = h5lite
        return True


def install(reference_root: str):
    """Bind ``src.model.SPGG`` (and the RL algorithm classes) of the reference found under
    ``reference_root`` to the B200 implementations.  Returns the reference's ``src.model``."""
    root = os.path.abspath(reference_root)
    if not os.path.isfile(os.path.join(root, "src", "model", "__init__.py")):
        raise FileNotFoundError(f"no reference tree under {root} (src/model/__init__.py missing)")
    if root not in sys.path:
        sys.path.insert(0, root)
    install_h5py_shim()
    from . import (SPGG, RLAlgorithm, QLearning, SARSA, ExpectedSARSA, DoubleQLearning,
                   create_algorithm)
    model = importlib.import_module("src.model")
    for name, obj in (("SPGG", SPGG), ("RLAlgorithm", RLAlgorithm), ("QLearning", QLearning),
                      ("SARSA", SARSA), ("ExpectedSARSA", ExpectedSARSA),
                      ("DoubleQLearning", DoubleQLearning), ("create_algorithm", create_algorithm)):
        setattr(model, name, obj)
    # modules that already did `from ..model import SPGG`
    runner = sys.modules.get("src.experiments.runner")
    if runner is not None:
        runner.SPGG = SPGG
    return model


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) < 2:
        raise SystemExit("usage: python -m spgg_b200.dropin <reference root> <script relative to it> [args...]")
    root, script = os.path.abspath(argv[0]), argv[1]
    install(root)
    path = script if os.path.isabs(script) else os.path.join(root, script)
    sys.argv = [path] + argv[2:]
    runpy.run_path(path, run_name="__main__")


if __name__ == "__main__":
    main()
