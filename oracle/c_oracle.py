"""TEST INFRASTRUCTURE ONLY - ctypes front-end of ``oracle/spgg_oracle.c``.

Used by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs as the *checker*; never imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libspgg_oracle.so")
NSTAT = 40

# column indices of a stat row (mirror of spgg_oracle.h / include/spgg.h)
ST = dict(NC_OLD=0, N_CD=1, N_DC=2, NC_NEW=3, SUM_P=4, SUM_P_C=5, SUM_P_D=6, SUM_WP_P=7,
          SUM_REW_C=8, SUM_REW_D=9, SUM_RATIO=10, GROUP0=11, SUM_R=17, SUM_Q=18,
          SUM_Q_C=22, SUM_Q_D=26, SUM_NI=30, N_BEST_POS=31, N_BEST_2ND=32, GMAX=33)


class OracleParams(C.Structure):
    _fields_ = [("L", C.c_int32), ("M", C.c_int32), ("state_mode", C.c_int32),
                ("reserved", C.c_int32),
                ("r", C.c_double), ("c", C.c_double), ("cost", C.c_double),
                ("alpha", C.c_double), ("gamma", C.c_double), ("kappa", C.c_double),
                ("lambda_eps", C.c_double), ("rep_gain_C", C.c_double),
                ("delta_R_D", C.c_double), ("R_min", C.c_double), ("R_max", C.c_double),
                ("wP", C.c_double)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "spgg_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libspgg_oracle.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        dp = C.POINTER(C.c_double)
        for name, qt in (("oracle_step_f64", C.c_double), ("oracle_step_f32", C.c_float)):
            fn = getattr(_lib, name)
            fn.restype = C.c_int
            fn.argtypes = [C.POINTER(OracleParams), C.c_void_p, C.c_void_p, C.c_void_p,
                           C.c_double, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32,
                           C.c_uint32, dp]
        _lib.oracle_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oracle_thr24.restype = C.c_uint32
        _lib.oracle_thr24.argtypes = [C.c_double]
        _lib.oracle_reward_table.argtypes = [C.POINTER(OracleParams), C.c_void_p]
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def params_from_dict(p: dict) -> OracleParams:
    """``p`` uses the reference ctor's argument names (spgg.py:50-56)."""
    return OracleParams(
        L=int(p["L"]), M=2 if p.get("use_second_order", True) else 1,
        state_mode=1 if p.get("state_representation", "reputation") == "action" else 0,
        reserved=0, r=float(p["r"]), c=float(p.get("c", 1)), cost=float(p.get("cost", 0.5)),
        alpha=float(p.get("alpha", 0.1)), gamma=float(p.get("gamma", 0.9)),
        kappa=float(p.get("influence_factor", 1.0)),
        lambda_eps=float(p.get("lambda_epsilon", 0.01)),
        rep_gain_C=float(p.get("rep_gain_C", 0.5)), delta_R_D=float(p.get("delta_R_D", 1)),
        R_min=float(p.get("R_min", -10)), R_max=float(p.get("R_max", 10)),
        wP=float(p.get("reward_weight_payoff", 1.0)))


def philox(ctr, key):
    c = np.asarray(ctr, np.uint32)
    k = np.asarray(key, np.uint32)
    out = np.zeros(4, np.uint32)
    lib().oracle_philox4x32_10(c.ctypes.data, k.ctypes.data, out.ctypes.data)
    return out


def thr24(eps: float) -> int:
    return int(lib().oracle_thr24(float(eps)))


def reward_table(p: dict) -> np.ndarray:
    tab = np.zeros(128, np.float32)
    op = params_from_dict(p)
    lib().oracle_reward_table(C.byref(op), tab.ctypes.data)
    return tab


class Sim:
    """Stateful driver: holds S (uint8), R, Q and steps them with either
    arithmetic.  ``precision`` = 'fp64' (reference order) or 'fp32'
    (throughput-mode arithmetic)."""

    def __init__(self, p: dict, S0, R0, Q0, precision="fp64", seed=0):
        self.p = dict(p)
        self.op = params_from_dict(p)
        self.L = int(p["L"])
        self.precision = precision
        ft = np.float64 if precision == "fp64" else np.float32
        self.S = np.ascontiguousarray(np.asarray(S0) != 0, dtype=np.uint8)
        self.R = np.ascontiguousarray(R0, dtype=ft)
        self.Q = np.ascontiguousarray(np.asarray(Q0).reshape(self.L, self.L, 2, 2), dtype=ft)
        self.eps = float(p.get("epsilon", 0.5))
        self.decay = float(p.get("epsilon_decay", 0.995))
        self.eps_min = float(p.get("epsilon_min", 0.01))
        self.seed = int(seed)
        self.t = 0
        self.fn = lib().oracle_step_f64 if precision == "fp64" else lib().oracle_step_f32

    def step(self, u=None, b=None):
        """One iteration; returns the stat row (NSTAT doubles)."""
        self.t += 1
        st = np.zeros(NSTAT)
        if u is not None:
            u = np.ascontiguousarray(u, np.float64)
            b = np.ascontiguousarray(b, np.uint8)
            up, bp = u.ctypes.data, b.ctypes.data
        else:
            up = bp = None
        rc = self.fn(C.byref(self.op), self.S.ctypes.data, self.R.ctypes.data,
                     self.Q.ctypes.data, self.eps, up, bp, self.seed, self.t,
                     thr24(self.eps), st.ctypes.data_as(C.POINTER(C.c_double)))
        assert rc == 0
        self.eps = max(self.eps * self.decay, self.eps_min)  # algorithms.py:42
        return st

    def run(self, n, draws=None):
        """``draws(t, L) -> (u, b)`` or None for Philox.  Stops like spgg.py:405
        when the lattice is uniform.  Returns the stat rows of completed steps."""
        rows = []
        for _ in range(n):
            nC = int((self.S == 0).sum())
            if nC == 0 or nC == self.L * self.L:
                break
            if draws is not None:
                u, b = draws(self.t + 1, self.L)
                rows.append(self.step(u, b))
            else:
                rows.append(self.step())
        return np.array(rows).reshape(-1, NSTAT)
