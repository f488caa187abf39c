"""GPU-aware experiment runner behind the reference's runner API
(``src/experiments/runner.py``: ``get_folder_name`` :11, ``run_one_experiment`` :48,
``run_experiments`` :117).

The reference forks one OS process per parameter tuple (``multiprocessing.Pool``,
``runner.py:142-154``); here the tuples that share a lattice geometry become batched replicas
of one device handle (``spgg.run_models``) and, under ``torchrun``, the batches are dealt to
the ranks (one process per GPU, no communication between replicas).  Folder layout, file
names, the hard-coded model constants of ``runner.py:88-101`` and the returned
``(params, (final_coop_ratio, final_rep_mean))`` tuples are the reference's.
"""
from __future__ import annotations

import os
from typing import List, Tuple

from . import sweep
from .spgg import SPGG, run_models

# runner.py:88-101: every CLI experiment of the reference uses these
RUNNER_MODEL = dict(c=1, cost=1, iterations=100001, L=100, num_of_strategies=2, K=0.1,
                    population_type=0, gamma=0.9, epsilon=0.5, epsilon_decay=0.99,
                    epsilon_min=0.01, lambda_epsilon=0.01, delta_R_C=1, delta_R_D=1,
                    R_min=-10, R_max=10)


def get_folder_name(r, kappa, use_second_order, alpha, reward_weight_payoff, rep_gain_C,
                    state_representation='reputation', algorithm='qlearning') -> str:
    """Same string as runner.py:40-45."""
    suffix = ("_action" if state_representation == 'action' else "") + \
             (f"_{algorithm}" if algorithm != 'qlearning' else "")
    return (f"results_r{r}_inf{kappa}_order{use_second_order}_alpha{alpha}_"
            f"rw{reward_weight_payoff:.2f}_rgC{rep_gain_C:.2f}{suffix}")


def _unpack(params: Tuple):
    """6-, 7- or 8-tuples, like runner.py:62-72."""
    p = tuple(params)
    if len(p) == 8:
        return p
    if len(p) == 7:
        return p + ('qlearning',)
    if len(p) == 6:
        return p + ('reputation', 'qlearning')
    raise ValueError(f"parameter tuple of length {len(p)}; expected 6, 7 or 8 entries")


def _make_model(params: Tuple, overrides: dict, base_dir: str):
    r, kappa, second, alpha, wP, rgC, state, algo = _unpack(params)
    folder = os.path.join(base_dir, get_folder_name(r, kappa, second, alpha, wP, rgC, state, algo))
    for sub in ("", "configurations", "reputations", "plots", os.path.join("plots", "snapshots"), "data"):
        os.makedirs(os.path.join(folder, sub), exist_ok=True)           # runner.py:79-85
    kw = dict(RUNNER_MODEL, r=r, alpha=alpha, influence_factor=kappa, use_second_order=second,
              reward_weight_payoff=wP, rep_gain_C=rgC, state_representation=state, algorithm=algo)
    kw.update(overrides)
    m = SPGG(**kw)
    m.folder = folder
    return m, os.path.join(folder, "data", "experiment_data.h5")


def _finish(params, model, record, verbose=True):
    final_coop_ratio, _final_def_ratio, _ = record
    final_rep_mean = model.rep_avg_history[-1] if model.rep_avg_history else 0   # runner.py:108
    r, kappa, second, alpha, wP, rgC, state, algo = _unpack(params)
    if verbose:
        print(f"Done: r={r}, κ={kappa}, M={2 if second else 1}, α={alpha}, w_P={wP}, "
              f"ΔR_C={rgC}, state={'action' if state == 'action' else 'rep'}, algo={algo}")
    return params, (final_coop_ratio, final_rep_mean)


def run_one_experiment(params: Tuple, base_dir: str = ".", **overrides):
    """runner.py:48-114 for one tuple.  ``overrides`` (e.g. ``iterations=``, ``L=``, ``seed=``,
    ``device=``) replace the hard-coded constants of runner.py:88-101."""
    model, fname = _make_model(params, overrides, base_dir)
    return _finish(params, model, model.run(fname))


def run_experiments(param_combinations: List[Tuple], num_processes: int = None,
                    use_progress_bar: bool = True, base_dir: str = ".", max_batch: int = 32,
                    **overrides) -> List[Tuple]:
    """runner.py:117-156.  ``num_processes`` is accepted for signature compatibility: the
    parallelism is replicas per launch x GPUs (ranks of the current ``torch.distributed`` job),
    not host processes.  Returns ``[(params, (final_coop_ratio, final_rep_mean)), ...]`` in
    input order (the reference's ``imap_unordered`` order is arbitrary)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    combos = [tuple(p) for p in param_combinations]
    keyed = []
    for p in combos:
        r, kappa, second, alpha, wP, rgC, state, algo = _unpack(p)
        keyed.append(dict(L=overrides.get("L", RUNNER_MODEL["L"]), use_second_order=second,
                          state_representation=state, algorithm=algo))
    bar = None
    if use_progress_bar and rank == 0:
        try:
            from tqdm import tqdm
            bar = tqdm(total=len(combos), desc="Running simulations")
        except ImportError:
            print("tqdm not available, running without progress bar")
    mine = {}
    for owner, idx in sweep.plan(keyed, world, max_batch):
        if owner != rank:
            continue
        ov = dict(overrides)
        ov.setdefault("device", sweep.default_device())
        built = [_make_model(combos[i], ov, base_dir) for i in idx]
        records = run_models([m for m, _f in built], [f for _m, f in built])
        for i, (m, _f), rec in zip(idx, built, records):
            mine[i] = _finish(combos[i], m, rec, verbose=not use_progress_bar)
        if bar is not None:
            bar.update(len(idx))
    if bar is not None:
        bar.close()
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        mine = {}
        for part in parts:
            mine.update(part)
    return [mine[i] for i in range(len(combos))]
