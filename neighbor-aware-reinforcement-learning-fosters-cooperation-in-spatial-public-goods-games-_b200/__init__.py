"""B200-native SPGG lattice step behind the reference's Python surface.

``SPGG`` mirrors ``src/model/spgg.py::SPGG`` of the reference; ``Engine`` is the thin
wrapper over the C ABI (``include/spgg.h``) that ``SPGG.run`` drives.
"""
from ._lib import LIB_PATH, NSTAT, load  # noqa: F401
from .engine import Engine, params_struct  # noqa: F401

try:  # the class API needs only numpy + the engine
    from .spgg import SPGG  # noqa: F401
    from .algorithms import (RLAlgorithm, QLearning, SARSA, ExpectedSARSA,  # noqa: F401
                             DoubleQLearning, create_algorithm)
except ImportError:  # pragma: no cover - during bring-up
    pass
