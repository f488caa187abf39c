"""``SPGG``: the reference's model class (``src/model/spgg.py:39-637``) with the loop body
of ``run`` replaced by the fused CUDA step behind the C ABI.

Same constructor signature, attributes, ``run(filename) -> (coop, def, mean P)`` and
HDF5 dataset names/dtypes/shapes, so ``src/experiments/runner.py`` and the plotting
scripts of the reference work unchanged when ``src.model.SPGG`` is this class
(INTEGRATION.md).  Extra keyword arguments (stored in ``params`` like any other, as the
reference does with ``**params``):

``seed``       int: pins the NumPy stream of the ctor draws (same draws as the reference
               with its ``np.random.seed()`` call pinned) and keys the device Philox stream.
``precision``  ``"fp32"`` (default: throughput instantiation) or ``"fp64"`` (reference
               operation order, no FMA).
``draws``      ``"philox"`` (default, on device) or ``"numpy"``: replay the reference's
               own draw stream (``rand(L,L)`` then ``randint(0,2,(L,L))`` per step); with
               ``precision="fp64"`` and a ``seed`` this reproduces the reference bit for bit.
``device``     CUDA device index (default 0).
``chunk``      iterations per device call (default 4096).
"""
from __future__ import annotations

import os

import numpy as np

from . import _lib as L_
from . import h5lite, series
from .algorithms import RLAlgorithm, create_algorithm
from .engine import Engine

_REF_SNAPSHOTS = (1, 10, 100, 1000, 5000, 10000, 20000, 30000, 40000)   # spgg.py:153


class SPGG:
    def __init__(self, r=2, c=1, cost=0.5, K=0.1, L=50, iterations=1000,
                 num_of_strategies=2, population_type=0, S_in_one=None,
                 alpha=0.1, gamma=0.9, epsilon=0.5, epsilon_decay=0.995,
                 epsilon_min=0.01, influence_factor=1.0, use_second_order=True,
                 lambda_epsilon=0.01, delta_R_C=1, delta_R_D=1, R_min=-10,
                 R_max=10, reward_weight_payoff=1.0, rep_gain_C=0.5,
                 state_representation='reputation', algorithm='qlearning', **params):
        seed = params.get("seed")
        if seed is None:
            np.random.seed()                      # spgg.py:98
            rng = np.random                       # global stream, like the reference
            self._philox_seed = int(np.random.randint(0, 2 ** 31 - 1)) | (
                int(np.random.randint(0, 2 ** 31 - 1)) << 31)
        else:
            rng = np.random.RandomState(int(seed))
            self._philox_seed = int(seed)
        self._rng = rng

        all_params = dict(locals(), **params)     # spgg.py:101-105
        for k in ("self", "params", "rng", "seed"):
            all_params.pop(k, None)
        all_params.update(params)
        self.params = all_params
        for key, val in self.params.items():
            setattr(self, key, val)

        self.reward_weight_rep = 1 - self.reward_weight_payoff        # spgg.py:108
        if state_representation not in ("reputation", "action"):
            # the reference raises this from get_state() inside run() (spgg.py:309)
            self._bad_state = state_representation
        if isinstance(algorithm, str):                                 # spgg.py:111-118
            self.algorithm = create_algorithm(algorithm, alpha, gamma, epsilon, epsilon_decay,
                                              epsilon_min, **params)
        elif isinstance(algorithm, RLAlgorithm):
            self.algorithm = algorithm
        else:
            raise ValueError(f"algorithm must be str or RLAlgorithm, got {type(algorithm)}")

        self.q_table = rng.uniform(low=-0.01, high=0.01, size=(L, L, 2, 2))   # spgg.py:121
        self.R = np.zeros((L, L))                                              # spgg.py:129
        self.cache = {}
        self._Sn = S_in_one
        self.create_population()
        self.track_positions = [(L // 2, L // 2), (L // 4, L // 4), (3 * L // 4, 3 * L // 4)]
        self.q_history = {pos: {'q_c': [], 'q_d': []} for pos in self.track_positions}
        self.it_records = []
        self.epsilon_history = []
        self.rep_avg_history = []
        self.influence_counts = []
        self.best_neighbor_type_history = []
        self.normlize_max = 4 * r
        self.normlize_min = r - 5
        int(np.floor(np.log10(self.iterations)))  # spgg.py:152 (raises for iterations <= 0 like the reference)
        self.snapshot_iters = set(_REF_SNAPSHOTS)
        self.folder = None
        self.P = None
        self.kernel_launches = 0

    # ------------------------------------------------------------------ reference helpers
    def create_population(self):
        """spgg.py:158-164."""
        L = self.L
        if self._Sn is None:
            self._Sn = self._rng.randint(0, 2, size=(L, L))
        self._Sn = np.asarray(self._Sn)
        self._S = [(self._Sn == j).astype(int) for j in range(self.num_of_strategies)]
        return self._S

    def update_reputation(self, actions):
        """spgg.py:319-323 (host convenience; the run loop does this on the device)."""
        delta_R = np.where(actions == 0, self.rep_gain_C, -self.delta_R_D)
        self.R = np.clip(self.R + delta_R, self.R_min, self.R_max)

    # ------------------------------------------------------------------ engine plumbing
    def _engine_params(self):
        return dict(L=self.L, r=self.r, c=self.c, cost=self.cost, alpha=self.algorithm.alpha,
                    gamma=self.algorithm.gamma, epsilon=self.algorithm.epsilon,
                    epsilon_decay=self.algorithm.epsilon_decay,
                    epsilon_min=self.algorithm.epsilon_min,
                    influence_factor=self.influence_factor,
                    use_second_order=self.use_second_order, lambda_epsilon=self.lambda_epsilon,
                    delta_R_D=self.delta_R_D, R_min=self.R_min, R_max=self.R_max,
                    reward_weight_payoff=self.reward_weight_payoff, rep_gain_C=self.rep_gain_C,
                    state_representation=self.state_representation,
                    algorithm=getattr(self.algorithm, "name", "qlearning"))

    def run(self, filename):
        """spgg.py:325-637."""
        if getattr(self, "_bad_state", None) is not None:
            raise ValueError(f"Unknown state_representation: {self._bad_state}. "
                             f"Must be 'reputation' or 'action'")
        if getattr(self.algorithm, "kernel_tag", None) is None:
            raise ValueError(
                f"algorithm '{getattr(self.algorithm, 'name', type(self.algorithm).__name__)}' is not "
                "built into the fused CUDA step (only Q-learning is); there is no CPU fallback")
        L = self.L
        N = L * L
        precision = self.params.get("precision", "fp32")
        draws = self.params.get("draws", "philox")
        chunk_max = int(self.params.get("chunk", 4096))
        if draws == "numpy":
            chunk_max = max(1, min(chunk_max, (64 << 20) // (9 * N) or 1))
        pdict = self._engine_params()
        eps0 = float(self.algorithm.epsilon)

        snapshots_dir = os.path.join(self.folder, 'plots', 'snapshots') if self.folder else 'snapshots'
        os.makedirs(snapshots_dir, exist_ok=True)                      # spgg.py:365-366

        eng = Engine(pdict, seeds=self._philox_seed, precision=precision,
                     device=int(self.params.get("device", 0)))
        try:
            eng.set_state(self._Sn, self.R, self.q_table)
            rows_it, sum_r_before = [], []
            done, stopped = 0, False
            snaps = {}
            T = int(self.iterations)
            cut_points = sorted(i - 1 for i in self.snapshot_iters if 1 <= i <= T)
            with h5lite.open_file(filename, "w") as data_file:
                while done < T and not stopped:
                    if done in cut_points or (done == 0 and 1 in self.snapshot_iters):
                        self._snapshot(eng, done + 1, data_file, snaps)
                    nxt = min([c for c in cut_points if c > done] + [T])
                    n = min(nxt - done, chunk_max)
                    if draws == "numpy":
                        u = np.empty((n, L, L))
                        b = np.empty((n, L, L), np.uint8)
                        for t in range(n):                              # algorithms.py:105,108
                            u[t] = self._rng.rand(L, L)
                            b[t] = self._rng.randint(0, 2, size=(L, L))
                        eng.set_replay(u, b)
                    eng.step(n)
                    st = eng.status()
                    rows = eng.stats()
                    k = int(st.iteration) - done                        # iterations really completed
                    rows_it.append(rows[1:k + 1])
                    sum_r_before.append(rows[:k, L_.ST_SUM_R])
                    stop_sum_r = rows[k, L_.ST_SUM_R]
                    done += k
                    if st.stopped_at >= 0 and st.stopped_at <= done:
                        stopped = True
                        if (done + 1) in self.snapshot_iters and (done + 1) not in snaps:
                            self._snapshot(eng, done + 1, data_file, snaps)   # spgg.py:397 precedes :405
                S, R, Q = eng.get_state()
                rows_it = np.vstack(rows_it) if rows_it else np.zeros((0, L_.NSTAT))
                sum_r_before = np.concatenate(sum_r_before) if sum_r_before else np.zeros(0)
                ser = series.assemble(rows_it, sum_r_before, N, self.params, eps0, stopped=stopped,
                                      stop_sum_r=stop_sum_r if stopped else 0.0,
                                      stop_all_coop=bool((S == 0).all()))
                self._write_final(data_file, ser, S, R)
            self.kernel_launches = int(eng.status().kernel_launches)
        finally:
            eng.close()

        # post-run attributes the reference leaves behind
        self.q_table, self.R, self._Sn = Q, R, S.astype(np.int64)
        self._S = [(self._Sn == j).astype(int) for j in range(self.num_of_strategies)]
        for _ in range(done):
            self.algorithm.decay_epsilon()
        self.epsilon = self.algorithm.epsilon
        self.avg_q_history = {k: list(ser[f"avg_{k}_history"]) for k in series.Q_NAMES}
        self.q_history_by_strategy = {
            g: {k: list(ser[f"{g}_{k}_history"]) for k in series.Q_NAMES}
            for g in ("cooperators", "defectors")}
        self.group_composition_history = [list(ser[f"group_comp_d{k}_history"]) for k in range(6)]
        it = ser["it_records_final"]
        mean_P = float(it[-1, 3]) if len(it) else float("nan")
        nC = int((S == 0).sum())
        return (nC / N, (N - nC) / N, mean_P)

    def _snapshot(self, eng, i, data_file, snaps):
        """State before iteration i acts (spgg.py:397-402)."""
        S, R, _ = eng.get_state(want_q=False)
        data_file.create_dataset(f"R_snapshot_{i}", data=R)
        rep_hist, rep_bins = np.histogram(R, bins=20, range=(self.R_min, self.R_max))
        data_file.create_dataset(f"rep_hist_{i}", data=rep_hist)
        data_file.create_dataset(f"rep_bins_{i}", data=rep_bins)
        data_file.create_dataset(f"Sn_snapshot_{i}", data=S.astype(np.int64))
        snaps[i] = True

    def _write_final(self, data_file, ser, S, R):
        """Dataset names, order and dtypes of spgg.py:595-633."""
        for key in ("it_records_final", "epsilon_history_final", "rep_avg_history_final",
                    "coop_rate_history", "switch_C_to_D", "switch_D_to_C",
                    "neighbor_influence_percent", "payoff_component_history",
                    "rep_component_history", "best_neighbor_second_order_percent",
                    "reputation_reward_ratio", "avg_reward_C_history", "avg_reward_D_history"):
            data_file.create_dataset(key, data=ser[key])
        for k in range(6):
            data_file.create_dataset(f"group_comp_d{k}_history", data=ser[f"group_comp_d{k}_history"])
        for grp in ("cooperators", "defectors"):
            for nm in series.Q_NAMES:
                data_file.create_dataset(f"{grp}_{nm}_history", data=ser[f"{grp}_{nm}_history"])
        for nm in series.Q_NAMES:
            data_file.create_dataset(f"avg_{nm}_history", data=ser[f"avg_{nm}_history"])
        for pos in self.track_positions:                               # always empty, spgg.py:345,620-622
            data_file.create_dataset(f"q_c_pos_{pos[0]}_{pos[1]}_final", data=np.zeros(0))
            data_file.create_dataset(f"q_d_pos_{pos[0]}_{pos[1]}_final", data=np.zeros(0))
        data_file.create_dataset("Sn_final", data=S.astype(np.int64))
        data_file.create_dataset("R_final", data=R)
        rep_hist_final, rep_bins_final = np.histogram(R, bins=20, range=(self.R_min, self.R_max))
        data_file.create_dataset("rep_hist_final", data=rep_hist_final)
        data_file.create_dataset("rep_bins_final", data=rep_bins_final)
        # 4-connected, non-periodic clusters of cooperators (spgg.py:631-633), linear time
        from scipy.ndimage import label
        clusters, n_clusters = label(S == 0)
        sizes = np.bincount(clusters.ravel(), minlength=n_clusters + 1)[1:]
        data_file.create_dataset("cluster_sizes", data=sizes.astype(np.int64))
