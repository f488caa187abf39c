"""TEST / BASELINE INFRASTRUCTURE ONLY - never imported by the product package.

Copies the UNMODIFIED reference tree (``/root/reference``: ``src/``, ``scripts/``,
``config/``) into ``oracle/_ref/`` so that it travels to the GPU box with the ``gpurun``
snapshot (``oracle/_ref/`` is git-ignored: the reference's sources never enter this
repository's history, and nothing here edits them).  On the GPU box the copy serves

* ``bench.py --impl reference`` and the ``cpu_baseline`` leg: the reference's own NumPy loop
  timed on the box's host cores, and
* ``tests/test_reference_scripts.py``: the reference's ``scripts/run_experiments.py`` /
  ``src/experiments/runner.py`` / ``src/visualization/plotting.py`` executed unchanged against
  the drop-in class.

``__graft_entry__.build()`` calls ``fetch()`` whenever ``/root/reference`` exists.

    python -m oracle.fetch_ref [--force]
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE = os.environ.get("SPGG_REFERENCE_ROOT", "/root/reference")
DEST = os.path.join(HERE, "_ref")
PARTS = ("src", "scripts", "config", "requirements.txt", "LICENSE")


def _is_reference(root: str) -> bool:
    return os.path.isfile(os.path.join(root, "src", "model", "spgg.py"))


def fetch(force: bool = False) -> str | None:
    """Copy the reference next to the oracle; returns the destination, or None when the
    source tree is not present (GPU box: the prebuilt copy, if any, is used as it is)."""
    if not _is_reference(SOURCE):
        return DEST if _is_reference(DEST) else None
    if _is_reference(DEST) and not force:
        src_m = os.path.getmtime(os.path.join(SOURCE, "src", "model", "spgg.py"))
        if os.path.getmtime(os.path.join(DEST, "src", "model", "spgg.py")) >= src_m:
            return DEST
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc", ".git")
    for part in PARTS:
        s = os.path.join(SOURCE, part)
        if os.path.isdir(s):
            shutil.copytree(s, os.path.join(DEST, part), ignore=ignore)
        elif os.path.isfile(s):
            shutil.copy2(s, os.path.join(DEST, part))
    with open(os.path.join(DEST, "PROVENANCE.txt"), "w") as f:
        f.write(f"verbatim copy of {SOURCE} made by oracle/fetch_ref.py; git-ignored, not part of this repository\n")
    return DEST


def reference_root() -> str | None:
    """Where an importable copy of the reference lives: the live tree in the build container,
    else the copy that travelled with the snapshot, else None."""
    for root in (SOURCE, DEST):
        if _is_reference(root):
            return root
    return None


if __name__ == "__main__":
    print(fetch(force="--force" in sys.argv))
