"""Row-strip decomposition of one large lattice over the GPUs of a node
(BASELINE config 5: L=32768 at 2/4/8 B200), one process per GPU.

The reference has no distributed mode (its only parallelism is one process per parameter
tuple, ``src/experiments/runner.py:117-156``); what the strips must reproduce is the
single-lattice loop body ``src/model/spgg.py:368-592``.  Only two things cross a strip
boundary in that loop:

* the stencils read neighbours' R (radius M, ``spgg.py:292-307``), strategy bits (radius 2,
  ``spgg.py:373-377``) and reward codes (radius M, ``spgg.py:478-486``)  ->  one halo
  exchange per iteration: ``GH`` boundary rows of the three planes, packed into one
  contiguous buffer per direction (``spgg_halo_pack`` / ``spgg_halo_unpack``);
* ``global_max = np.max(np.abs(diffs))`` (``spgg.py:488``) is lattice-global  ->  one scalar
  all-reduce(MAX) per iteration between the light k_gmax kernel and the fused k_step.

Q is strictly site-local and never moves.  The Philox counters are keyed on the *global*
row and column, so an N-strip run is bit-identical (S, R, Q, integer statistics) to the
single-GPU run of the same seed (tests/test_gpu_strips.py).

``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests) is the
transport; the kernels come from the C ABI.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time
from typing import Sequence

import numpy as np

from . import _lib as L_

# stat columns that are not plain sums over sites
_MAX_COLS = (L_.ST_GMAX,)


# ------------------------------------------------------------------ partition (pure host logic)
def strip_rows(L: int, world: int, rank: int, align: int = 16) -> tuple[int, int]:
    """(row0, rows) of strip ``rank``: contiguous blocks of rows, multiples of ``align``
    (the fast kernel's tile height) while the lattice allows it; remainders go to the first strips."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    unit = align if (L % align == 0 and L // align >= world) else 1
    blocks = L // unit
    if blocks < world:
        raise ValueError(f"L={L} cannot be split into {world} strips")
    base, extra = divmod(blocks, world)
    nb = base + (1 if rank < extra else 0)
    b0 = rank * base + min(rank, extra)
    rows, row0 = nb * unit, b0 * unit
    if rows < 4:
        raise ValueError(f"strips of {rows} rows are thinner than the two ghost-row halos")
    return row0, rows


def neighbours(world: int, rank: int) -> tuple[int, int]:
    """(up, down): ranks owning the rows just above (row0-1) and just below; periodic."""
    return (rank - 1) % world, (rank + 1) % world


def exchange_halos(dist, to_up, to_down, from_up, from_down, rank: int, world: int, group=None):
    """One halo exchange: my top rows go to ``up`` (they become its bottom ghosts), my
    bottom rows to ``down``.  Sends are posted [up, down], receives [down, up]: with two
    ranks both neighbours are the same peer and messages between a pair match in order."""
    up, down = neighbours(world, rank)
    if world == 1:
        from_down.copy_(to_up)       # my own top rows are the ghosts below my last row
        from_up.copy_(to_down)
        return
    ops = [dist.P2POp(dist.isend, to_up, up, group), dist.P2POp(dist.isend, to_down, down, group),
           dist.P2POp(dist.irecv, from_down, down, group), dist.P2POp(dist.irecv, from_up, up, group)]
    for w in dist.batch_isend_irecv(ops):
        w.wait()


def reduce_stat_rows(dist, rows, world: int, device=None, group=None):
    """Per-strip statistic rows (sums over the strip's sites) -> whole-lattice rows on every
    rank.  Counts are exact integers in doubles, so their sum is exact; ST_GMAX is a max."""
    import torch
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    if world == 1:
        return rows
    t = torch.from_numpy(rows.copy())
    if device is not None and dist.get_backend(group) != "gloo":
        t = t.to(device)
    mx = t[:, list(_MAX_COLS)].clone()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    t[:, list(_MAX_COLS)] = mx
    return t.cpu().numpy()


class _DevArray:
    """Zero-copy view of device memory owned by the C library as a torch tensor."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


# ------------------------------------------------------------------ the strip engine
class StripEngine:
    """One strip of an ``L x L`` lattice on this process's GPU.  ``params`` uses the
    reference ctor's argument names (spgg.py:50-56)."""

    def __init__(self, params: dict, seed: int = 0, precision: str = "fp32", device: int | None = None,
                 group=None):
        import torch
        import torch.distributed as dist
        from .engine import Engine
        self.torch, self.dist, self.group = torch, dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.L = int(params.get("L", 50))
        self.row0, self.rows = strip_rows(self.L, self.world, self.rank)
        self.device = int(torch.cuda.current_device() if device is None else device)
        self.precision = precision
        # every strip uses the same Philox key: counters carry the global row
        self.eng = Engine(params, seeds=seed, precision=precision, device=self.device,
                          rows=self.rows, row0=self.row0)
        self.lib, self.h = self.eng.lib, self.eng._h
        nb = int(self.lib.spgg_halo_bytes(self.h))
        dev = torch.device("cuda", self.device)
        self.buf = {k: torch.empty(nb, dtype=torch.uint8, device=dev)
                    for k in ("to_up", "to_down", "from_up", "from_down")}
        self._gtype = "<f8" if precision == "fp64" else "<f4"
        # gloo (CPU tests, or several ranks sharing one GPU) moves the halos through the host
        self.host_staged = dist.is_initialized() and dist.get_backend(group) == "gloo"
        if self.host_staged:
            self.hbuf = {k: torch.empty(nb, dtype=torch.uint8).pin_memory() for k in self.buf}
        self.iteration = 0
        self._last_n = 0

    # -- state
    def set_state_global(self, S, R, Q):
        """Every rank passes the whole lattice; each uploads its own rows."""
        a, b = self.row0, self.row0 + self.rows
        self.eng.set_state(np.asarray(S)[a:b], np.asarray(R)[a:b], np.asarray(Q)[a:b])
        self.iteration = 0

    def init_random(self, seed: int):
        self.eng.init_random(seed)
        self.iteration = 0

    def get_state_local(self, want_q=True):
        return self.eng.get_state(want_q=want_q)

    def gather_state(self, want_q=True):
        """Whole lattice on every rank (test / small-lattice helper)."""
        S, R, Q = self.eng.get_state(want_q=want_q)
        if self.world == 1:
            return S, R, Q
        parts: list = [None] * self.world
        self.dist.all_gather_object(parts, (S, R, Q), group=self.group)
        cat = lambda i: np.concatenate([p[i] for p in parts], axis=0)
        return cat(0), cat(1), (cat(2) if want_q else None)

    # -- stepping
    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def _exchange(self, st):
        b = self.buf
        L_.check(self.lib.spgg_halo_pack(self.h, b["to_up"].data_ptr(), b["to_down"].data_ptr(), st))
        if self.host_staged:
            hb = self.hbuf
            hb["to_up"].copy_(b["to_up"]); hb["to_down"].copy_(b["to_down"])
            self.torch.cuda.current_stream().synchronize()
            exchange_halos(self.dist, hb["to_up"], hb["to_down"], hb["from_up"], hb["from_down"],
                           self.rank, self.world, self.group)
            b["from_up"].copy_(hb["from_up"]); b["from_down"].copy_(hb["from_down"])
        else:
            exchange_halos(self.dist, b["to_up"], b["to_down"], b["from_up"], b["from_down"],
                           self.rank, self.world, self.group)
        L_.check(self.lib.spgg_halo_unpack(self.h, b["from_up"].data_ptr(), b["from_down"].data_ptr(), st))

    def step(self, n: int):
        """n iterations of spgg.py:368-592 over the whole lattice (all ranks call this)."""
        lib, h = self.lib, self.h
        st = self._stream()
        L_.check(lib.spgg_begin_steps(h, int(n), st))
        # device table of the per-iteration maxima of this call: entry s belongs to iteration s
        gtab = self.torch.as_tensor(_DevArray(lib.spgg_gmax_device_ptr(h), n + 1, self._gtype),
                                    device=self.torch.device("cuda", self.device))
        self._exchange(st)                                   # ghosts of the starting state
        L_.check(lib.spgg_phase_kernel(h, 0, 1, st))         # action of the first iteration
        for s in range(1, n + 1):
            self._exchange(st)                               # codes / R / strategies just written
            L_.check(lib.spgg_phase_gmax(h, st))             # strip-local max |reward difference|
            if self.world > 1:                               # spgg.py:488 is lattice-global
                g = gtab[s:s + 1]
                if self.host_staged:
                    gh = g.cpu()
                    self.dist.all_reduce(gh, op=self.dist.ReduceOp.MAX, group=self.group)
                    g.copy_(gh)
                else:
                    self.dist.all_reduce(g, op=self.dist.ReduceOp.MAX, group=self.group)
            L_.check(lib.spgg_phase_kernel(h, 1, 1 if s < n else 0, st))
        L_.check(lib.spgg_end_steps(h, st))
        self.iteration += int(n)
        self._last_n = int(n)

    def sync(self):
        self.eng.sync()

    def stats_local(self):
        self.eng._last_n = self._last_n
        return self.eng.stats()

    def stats(self):
        """Whole-lattice statistic rows of the last ``step`` call (row 0 = starting state)."""
        return reduce_stat_rows(self.dist, self.stats_local(), self.world,
                                device=self.torch.device("cuda", self.device), group=self.group)

    def kernel_launches(self) -> int:
        return int(self.eng.status().kernel_launches)

    def close(self):
        self.eng.close()


# ------------------------------------------------------------------ bench leg for N > 1
def bench_main(args, rank: int, local_rank: int, world: int):
    """``bench.py --gpus N`` under torchrun: BASELINE config 5, one L x L lattice (default
    L=32768) split into N row strips; barrier + synchronize on both sides of exactly K timed
    steps, device-side timing, max over ranks; rank 0 prints the JSON line."""
    import torch
    import torch.distributed as dist
    from bench import C4, BYTES_PER_SITE_FP32, ClockSampler, measured_peak_gbs

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    L = args.L or 32768
    K, W, inner = args.steps, args.warmup, args.inner
    p = dict(C4, L=L)
    se = StripEngine(p, seed=2024, precision="fp32", device=local_rank)
    se.init_random(2024)
    for _ in range(W):
        se.step(inner)
    se.sync()
    l0 = se.kernel_launches()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        se.step(inner)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    se.sync()
    launches = se.kernel_launches() - l0
    n_sites = L * L
    value = n_sites * inner * K / (ms * 1e-3)

    # end to end: host state in (pinned), K steps with the whole-lattice stat rows read back
    # each step, final strategies/reputations out; every rank moves its own strip
    S_h = torch.empty((se.rows, L), dtype=torch.uint8).pin_memory()
    R_h = torch.zeros((se.rows, L), dtype=torch.float64).pin_memory()
    Q_h = torch.empty((se.rows, L, 2, 2), dtype=torch.float64).pin_memory()
    rs = np.random.RandomState(100 + rank)
    S_h.numpy()[...] = rs.randint(0, 2, (se.rows, L))
    Q_h.numpy()[...] = rs.uniform(-0.01, 0.01, (se.rows, L, 2, 2))
    S_o, R_o = torch.empty_like(S_h).pin_memory(), torch.empty_like(R_h).pin_memory()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    L_.check(se.lib.spgg_set_state(se.h, 0, S_h.data_ptr(), R_h.data_ptr(), Q_h.data_ptr()))
    d2h = 0
    for _ in range(K):
        se.step(inner)
        d2h += se.stats().nbytes
    L_.check(se.lib.spgg_get_state(se.h, 0, S_o.data_ptr(), R_o.data_ptr(), None))
    torch.cuda.synchronize()
    dist.barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    t_e2e = float(t_e2e.item())
    h2d_bytes = world * (S_h.numel() + 8 * R_h.numel() + 8 * Q_h.numel())
    d2h_bytes = world * (S_o.numel() + 8 * R_o.numel() + d2h)

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        achieved = BYTES_PER_SITE_FP32 * value / 1e9 / world
        line = {
            "metric": "site-updates/s", "value": value, "unit": "site-updates/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C5: one L={L} lattice in {world} row strips of {se.rows} rows, "
                                   "reputation state, M=1, r=3, kappa=1, w_P=0.95, Q-learning; "
                                   f"{inner} iterations per bench step, per iteration one halo "
                                   "exchange (NCCL send/recv) + one all-reduce(MAX) of the "
                                   "global reward-difference maximum; statistics on, Philox draws",
                       "L": L, "iterations_per_step": inner, "strip_rows": se.rows,
                       "precision": "fp32 Q (float4) + int8 R + bit S",
                       "l2": "per-GPU state far larger than the 126 MB L2 (no flush needed)",
                       "scaling_note": "strong scaling over N>=2 at fixed L; N=1 runs config 4 (L=4096)"},
            "clocks": clocks,
            "e2e": {"value": n_sites * inner * K / t_e2e, "unit": "site-updates/s",
                    "h2d_bytes_per_step": h2d_bytes / K, "d2h_bytes_per_step": d2h_bytes / K,
                    "seconds": t_e2e},
            "gpu_launches": int(launches) * world,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None,
                         "kernel": "whole iteration (k_step + k_gmax + halo exchange + all-reduce), per GPU",
                         "algorithmic_bytes_per_site": BYTES_PER_SITE_FP32, "peak_source": peak_src},
            "cpu_baseline": None,
        }
        print(json.dumps(line), flush=True)
    se.close()
    dist.barrier()
    dist.destroy_process_group()
