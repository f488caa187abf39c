"""Importable alias of the package directory
``neighbor-aware-reinforcement-learning-fosters-cooperation-in-spatial-public-goods-games-_b200``
(its name is not a valid Python identifier)."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(
    os.path.dirname(os.path.abspath(__file__)),
    "neighbor-aware-reinforcement-learning-fosters-cooperation-in-spatial-public-goods-games-_b200")
_spec = importlib.util.spec_from_file_location(
    "spgg_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["spgg_b200"] = _mod
_spec.loader.exec_module(_mod)
