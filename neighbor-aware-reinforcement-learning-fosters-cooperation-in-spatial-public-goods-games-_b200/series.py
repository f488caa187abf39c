"""Statistic rows (sums, ``include/spgg.h``) -> the series the reference appends per
iteration and writes to HDF5 (``src/model/spgg.py:381-394, 419-426, 512-592, 595-618``)."""
from __future__ import annotations

import numpy as np

from . import _lib as L_

Q_NAMES = ("q_s0_c", "q_s0_d", "q_s1_c", "q_s1_d")  # spgg.py:356-359


def _safe_div(num, den, empty):
    num = np.asarray(num, dtype=np.float64)
    den = np.asarray(den, dtype=np.float64)
    out = np.full(num.shape, empty, dtype=np.float64)
    np.divide(num, den, out=out, where=den > 0)
    return out


def epsilon_after(eps0: float, decay: float, eps_min: float, n: int) -> np.ndarray:
    """epsilon recorded after each of n iterations (algorithms.py:40-42, spgg.py:548-550):
    e <- max(e * decay, eps_min), iterated.  For a non-increasing sequence (0 <= decay <= 1,
    eps0 >= eps_min) this is the running left-to-right product clamped from below - the same
    sequence of roundings, without a Python loop per iteration."""
    n = int(n)
    if n <= 0:
        return np.empty(0)
    e0, d, m = float(eps0), float(decay), float(eps_min)
    if 0.0 <= d <= 1.0 and e0 >= m >= 0.0:
        out = np.empty(n + 1)
        out[0] = e0
        out[1:] = d
        np.multiply.accumulate(out, out=out)       # ((e0*d)*d)*d ... sequentially
        below = np.nonzero(out[1:] < m)[0]
        if below.size == 0:
            return out[1:].copy()
        k = int(below[0])                          # first clamped entry; once at eps_min it stays there
        if m * d <= m:
            res = out[1:].copy()
            res[k:] = m
            return res
    out = np.empty(n)
    e = e0
    for t in range(n):
        e = max(e * d, m)
        out[t] = e
    return out


def uniform_payoff(all_coop: bool, r, c, cost):
    """Normalised payoff of every site when the lattice is uniform, evaluated in the
    reference's operation order (spgg.py:256-257, 373-377)."""
    n = 5 if all_coop else 0
    share = r * c * n / 5
    C, D = (1, 0) if all_coop else (0, 1)
    term = (share - cost) * C + share * D
    tot = term
    for _ in range(4):
        tot = tot + term
    lo = r - 5
    return (tot - lo) / (4 * r - lo)


def assemble(rows_iter: np.ndarray, sum_r_before: np.ndarray, n_sites: int, params: dict,
             eps0: float, stopped: bool = False, stop_sum_r: float = 0.0,
             stop_all_coop: bool = False) -> dict:
    """``rows_iter``: (T, NSTAT) rows of the T completed iterations; ``sum_r_before``:
    (T,) sum of R before each of them.  If ``stopped`` the loop broke at iteration T+1
    on a uniform lattice (spgg.py:405) after recording its pre-action entries."""
    T = rows_iter.shape[0]
    N = float(n_sites)
    r = rows_iter
    wP = params.get("reward_weight_payoff", 1.0)
    wR = 1 - wP                                                    # spgg.py:108
    nC = r[:, L_.ST_NC_OLD]
    nD = N - nC
    nCn = r[:, L_.ST_NC_NEW]
    out = {}
    coop = nC / N
    it = np.stack([coop, nD / N, r[:, L_.ST_SUM_P], r[:, L_.ST_SUM_P] / N,
                   _safe_div(r[:, L_.ST_SUM_P_C], nC, 0.0),
                   _safe_div(r[:, L_.ST_SUM_P_D], nD, 0.0)], axis=1) if T else np.zeros((0, 6))
    rep_avg = np.asarray(sum_r_before, dtype=np.float64) / N
    if stopped:
        Pv = uniform_payoff(stop_all_coop, params.get("r", 2), params.get("c", 1),
                            params.get("cost", 0.5))
        c_last = 1.0 if stop_all_coop else 0.0
        coop = np.append(coop, c_last)
        it = np.vstack([it, [c_last, 1.0 - c_last, Pv * N, Pv, Pv if stop_all_coop else 0,
                             0 if stop_all_coop else Pv]])
        rep_avg = np.append(rep_avg, stop_sum_r / N)
    out["it_records_final"] = it
    out["epsilon_history_final"] = epsilon_after(eps0, params.get("epsilon_decay", 0.995),
                                                 params.get("epsilon_min", 0.01), T)
    out["rep_avg_history_final"] = rep_avg
    out["coop_rate_history"] = coop
    out["switch_C_to_D"] = np.rint(r[:, L_.ST_N_CD]).astype(np.int64)
    out["switch_D_to_C"] = np.rint(r[:, L_.ST_N_DC]).astype(np.int64)
    out["neighbor_influence_percent"] = r[:, L_.ST_SUM_NI] / N
    out["payoff_component_history"] = r[:, L_.ST_SUM_WP_P] / N
    out["rep_component_history"] = (wR * 0.5) * nCn / N
    out["best_neighbor_second_order_percent"] = _safe_div(
        r[:, L_.ST_N_BEST_2ND], r[:, L_.ST_N_BEST_POS], 0.0) * 100
    out["reputation_reward_ratio"] = _safe_div(r[:, L_.ST_SUM_RATIO], nCn, np.nan)
    out["avg_reward_C_history"] = _safe_div(r[:, L_.ST_SUM_REW_C], nCn, 0.0)
    out["avg_reward_D_history"] = _safe_div(r[:, L_.ST_SUM_REW_D], N - nCn, 0.0)
    for k in range(6):
        out[f"group_comp_d{k}_history"] = r[:, L_.ST_GROUP0 + k] / N * 100
    for z, nm in enumerate(Q_NAMES):
        out[f"cooperators_{nm}_history"] = _safe_div(r[:, L_.ST_SUM_Q_C + z], nC, np.nan)
        out[f"defectors_{nm}_history"] = _safe_div(r[:, L_.ST_SUM_Q_D + z], nD, np.nan)
    for z, nm in enumerate(Q_NAMES):
        out[f"avg_{nm}_history"] = r[:, L_.ST_SUM_Q + z] / N
    return out
