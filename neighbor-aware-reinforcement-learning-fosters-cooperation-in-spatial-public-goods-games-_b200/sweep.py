"""Parameter sweeps as batched independent replicas (BASELINE config 3).

The reference runs one OS process per parameter tuple (``src/experiments/runner.py:117-156``:
``multiprocessing.Pool.imap_unordered(run_one_experiment, combos)``), each constructing an
``SPGG`` and looping in NumPy.  Here the tuples that share the lattice geometry
(L, M = use_second_order, state_representation) are batched into ONE device handle - the
fused kernels take ``gridDim = replicas x CTAs`` - and the batches are dealt round-robin to
the ranks (one process per GPU).  Replicas never communicate; results are gathered on the host.

Parameter dicts use the reference ctor's argument names (``spgg.py:50-56``).
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

from . import _lib as L_
from . import series


def default_device() -> int:
    """The GPU of this process: under torchrun (one process per GPU) LOCAL_RANK names it - and it is
    made torch's current device, which is 0 on every rank until somebody says otherwise; in a plain
    process, torch's current device."""
    import os
    import torch
    n = torch.cuda.device_count()
    if "LOCAL_RANK" in os.environ and n > 0:
        d = int(os.environ["LOCAL_RANK"]) % n
        torch.cuda.set_device(d)
        return d
    return torch.cuda.current_device()


def group_key(p: dict):
    """Replicas can share a launch iff these agree (spgg_create's batch-invariant fields)."""
    return (int(p.get("L", 50)), bool(p.get("use_second_order", True)),
            str(p.get("state_representation", "reputation")), str(p.get("algorithm", "qlearning")).lower())


def plan(param_list: Sequence[dict], world: int = 1, max_batch: int = 64):
    """Deterministic assignment: [(rank, [replica indices])...].  Replicas are grouped by
    ``group_key`` (stable order), cut into batches of at most ``max_batch`` and the batches
    dealt round-robin over the ranks."""
    groups: dict = {}
    for i, p in enumerate(param_list):
        groups.setdefault(group_key(p), []).append(i)
    batches = []
    for key in sorted(groups, key=lambda k: (k[0], k[1], k[2], k[3])):
        idx = groups[key]
        # spread a group over the ranks before filling batches, so all GPUs work on it
        n_b = max(world if len(idx) >= world else 1, -(-len(idx) // max_batch))
        n_b = min(n_b, len(idx))
        size = -(-len(idx) // n_b)
        batches += [idx[k:k + size] for k in range(0, len(idx), size)]
    return [(b % max(1, world), batch) for b, batch in enumerate(batches)]


def run_batch(params: Sequence[dict], seeds: Sequence[int], iterations: int, precision: str = "fp32",
              device: int = 0, chunk: int = 2048, init=None):
    """Run one batch of replicas to ``iterations`` (each stops early like spgg.py:405).
    ``init``: optional list of (S0, R0, Q0) per replica; default = the ctor's distributions
    drawn from ``np.random.RandomState(seed)`` (uniform Q, then randint S - spgg.py:121,162).
    Returns one dict per replica: series (reference dataset names), final S and R, iterations."""
    from .engine import Engine
    params = [dict(p) for p in params]
    n = len(params)
    L = int(params[0].get("L", 50))
    eng = Engine(params, seeds=list(seeds), precision=precision, device=device)
    try:
        for r in range(n):
            if init is not None:
                S0, R0, Q0 = init[r]
            else:
                rs = np.random.RandomState(int(seeds[r]) & 0x7FFFFFFF)
                Q0 = rs.uniform(-0.01, 0.01, (L, L, 2, 2))
                S0 = rs.randint(0, 2, (L, L))
                R0 = np.zeros((L, L))
            eng.set_state(S0, R0, Q0, replica=r)
        rows = [[] for _ in range(n)]
        before = [[] for _ in range(n)]
        last_sum_r = [0.0] * n
        done = [0] * n
        t = 0
        while t < iterations:
            k = min(chunk, iterations - t)
            eng.step(k)
            for r in range(n):
                st = eng.status(r)
                kk = int(st.iteration) - done[r]
                if kk > 0 or done[r] == 0:
                    rr = eng.stats(r)
                    rows[r].append(rr[1:kk + 1])
                    before[r].append(rr[:kk, L_.ST_SUM_R])
                    last_sum_r[r] = rr[kk, L_.ST_SUM_R]
                    done[r] += kk
            t += k
            if all(eng.status(r).stopped_at >= 0 for r in range(n)):
                break
        out = []
        for r in range(n):
            st = eng.status(r)
            stopped = st.stopped_at >= 0 and st.stopped_at <= done[r]
            S, R, _ = eng.get_state(r, want_q=False)
            ri = np.vstack(rows[r]) if rows[r] else np.zeros((0, L_.NSTAT))
            sb = np.concatenate(before[r]) if before[r] else np.zeros(0)
            ser = series.assemble(ri, sb, L * L, params[r], float(params[r].get("epsilon", 0.5)),
                                  stopped=stopped, stop_sum_r=last_sum_r[r],
                                  stop_all_coop=bool((S == 0).all()))
            out.append({"params": params[r], "seed": int(seeds[r]), "iterations": done[r],
                        "stopped": bool(stopped), "series": ser, "S": S, "R": R,
                        "final_coop": float((S == 0).mean())})
        return out
    finally:
        eng.close()


def run_sweep(param_list: Sequence[dict], seeds: Sequence[int] | None = None, iterations: int | None = None,
              precision: str = "fp32", max_batch: int = 64, chunk: int = 2048):
    """All replicas of ``param_list`` over the GPUs of this job.  Under ``torchrun`` every rank
    calls this with the same arguments (one process per GPU; no data-path collective, results
    are exchanged as Python objects); single process: everything on the current device.
    Returns the per-replica result dicts in input order (on every rank)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    device = default_device()
    n = len(param_list)
    if seeds is None:
        seeds = list(range(n))
    mine = {}
    for r, idx in plan(param_list, world, max_batch):
        if r != rank:
            continue
        its = iterations if iterations is not None else int(param_list[idx[0]].get("iterations", 1000))
        res = run_batch([param_list[i] for i in idx], [seeds[i] for i in idx], its, precision, device, chunk)
        for i, o in zip(idx, res):
            mine[i] = o
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        mine = {}
        for p in parts:
            mine.update(p)
    return [mine[i] for i in range(n)]
