"""TEST INFRASTRUCTURE ONLY - generates ``tests/golden/*.npz`` from the
UNMODIFIED reference executed in the build container (``/root/reference``,
see ``oracle/ref_harness.py``).  The fixtures travel to the GPU box; the
reference does not.

    python -m oracle.make_golden            # replay fixtures + seed band
    python -m oracle.make_golden --no-band  # replay fixtures only

Fixtures
--------
``replay_<name>.npz``  one recorded run of the reference: ctor state (q0, s0),
    per-step draws (u, b), final state (q_final, r_final, s_final) and the
    series the reference wrote to HDF5 (prefixed ``ds_``).
``band_<name>.npz``    cooperation-rate curves ``coop_rate_history`` of
    ``n_seeds`` reference runs (L=200, 1000 steps) - the seed-to-seed band the
    native-Philox kernel is judged against (BASELINE.json north_star).
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                          "tests", "golden")

# runner.py:88-101 hard-codes these for every experiment of the reference CLI
RUNNER_FIXED = dict(c=1, cost=1, num_of_strategies=2, K=0.1, population_type=0, gamma=0.9,
                    epsilon=0.5, epsilon_decay=0.99, epsilon_min=0.01, lambda_epsilon=0.01,
                    delta_R_C=1, delta_R_D=1, R_min=-10, R_max=10, alpha=0.8)

REPLAY_CASES = {
    # BASELINE config 1 physics (default_config.yaml): reputation, M=1, r=3, kappa=1, wP=.95
    "c1_rep_m1": dict(RUNNER_FIXED, L=16, iterations=40, r=3.0, influence_factor=1.0,
                      use_second_order=False, reward_weight_payoff=0.95, rep_gain_C=1.0,
                      state_representation="reputation"),
    # BASELINE config 2 physics: action state, M=2, r=4, kappa=1
    "c2_act_m2": dict(RUNNER_FIXED, L=16, iterations=40, r=4.0, influence_factor=1.0,
                      use_second_order=True, reward_weight_payoff=1.0, rep_gain_C=1.0,
                      state_representation="action"),
    "rep_m2_r36": dict(RUNNER_FIXED, L=12, iterations=40, r=3.6, influence_factor=0.5,
                       use_second_order=True, reward_weight_payoff=0.95, rep_gain_C=0.5,
                       state_representation="reputation"),
    "act_m1_k0": dict(RUNNER_FIXED, L=12, iterations=40, r=4, influence_factor=0.0,
                      use_second_order=False, reward_weight_payoff=1.0, rep_gain_C=1.0,
                      state_representation="action"),
    # ctor defaults of the reference (spgg.py:50-56): cost=.5, alpha=.1, decay=.995, rgC=.5
    "ctor_defaults": dict(L=10, iterations=30),
    # odd lattice size, non-integer reputation steps
    "odd_L_fracR": dict(RUNNER_FIXED, L=13, iterations=30, r=2.5, influence_factor=2.0,
                        use_second_order=True, reward_weight_payoff=0.83, rep_gain_C=0.25,
                        delta_R_D=0.75, state_representation="reputation"),
}

BAND_CASES = {
    "c1": dict(RUNNER_FIXED, L=200, iterations=1000, r=3.0, influence_factor=1.0,
               use_second_order=False, reward_weight_payoff=0.95, rep_gain_C=1.0,
               state_representation="reputation"),
    "c2": dict(RUNNER_FIXED, L=200, iterations=1000, r=4.0, influence_factor=1.0,
               use_second_order=True, reward_weight_payoff=1.0, rep_gain_C=1.0,
               state_representation="action"),
}

# long-run band ("long-run" means more than 10^3 iterations): the default_config.yaml
# lattice (L=100) for 10^4 iterations, curves stored at every 10th iteration
LONG_BAND_CASES = {
    "c1_long": dict(RUNNER_FIXED, L=100, iterations=10000, r=3.0, influence_factor=1.0,
                    use_second_order=False, reward_weight_payoff=0.95, rep_gain_C=1.0,
                    state_representation="reputation"),
}
LONG_STRIDE = 10

KEEP_DATASETS = (
    "coop_rate_history", "it_records_final", "rep_avg_history_final", "epsilon_history_final",
    "switch_C_to_D", "switch_D_to_C", "neighbor_influence_percent",
    "payoff_component_history", "rep_component_history", "best_neighbor_second_order_percent",
    "reputation_reward_ratio", "avg_reward_C_history", "avg_reward_D_history",
    "group_comp_d0_history", "group_comp_d3_history", "group_comp_d5_history",
    "avg_q_s0_c_history", "avg_q_s1_d_history", "cooperators_q_s0_d_history",
    "defectors_q_s1_c_history", "Sn_final", "R_final", "rep_hist_final", "rep_bins_final",
    "cluster_sizes", "R_snapshot_10", "Sn_snapshot_10", "rep_hist_10",
)


TD_CASES = {
    # other TD rules of algorithms.py (reference CI matrix, run-experiments.yml:17)
    "sarsa_rep_m1": dict(RUNNER_FIXED, L=12, iterations=30, r=3.0, influence_factor=1.0,
                         use_second_order=False, reward_weight_payoff=0.95, rep_gain_C=1.0,
                         state_representation="reputation", algorithm="sarsa"),
    "sarsa_act_m2": dict(RUNNER_FIXED, L=12, iterations=30, r=4.0, influence_factor=0.5,
                         use_second_order=True, reward_weight_payoff=1.0, rep_gain_C=1.0,
                         state_representation="action", algorithm="sarsa"),
    "expsarsa_rep_m2": dict(RUNNER_FIXED, L=12, iterations=30, r=3.6, influence_factor=1.0,
                            use_second_order=True, reward_weight_payoff=0.95, rep_gain_C=0.5,
                            state_representation="reputation", algorithm="expected_sarsa"),
    "expsarsa_act_m1": dict(RUNNER_FIXED, L=14, iterations=30, r=4.0, influence_factor=1.0,
                            use_second_order=False, reward_weight_payoff=1.0, rep_gain_C=1.0,
                            state_representation="action", algorithm="expected_sarsa"),
    "doubleq_rep_m1": dict(RUNNER_FIXED, L=12, iterations=30, r=3.0, influence_factor=1.0,
                           use_second_order=False, reward_weight_payoff=0.95, rep_gain_C=1.0,
                           state_representation="reputation", algorithm="double_qlearning"),
    "doubleq_act_m2": dict(RUNNER_FIXED, L=12, iterations=30, r=4.0, influence_factor=1.0,
                           use_second_order=True, reward_weight_payoff=1.0, rep_gain_C=1.0,
                           state_representation="action", algorithm="double_qlearning"),
}


def make_replay(name, params, seed):
    out = ref_harness.run_reference(seed, **params)
    extra = {}
    if "u" not in out:
        # rules that draw more than one pair per iteration: keep the stream as (n_steps, pairs, L, L).
        # SARSA: 3 (rand, randint) pairs (spgg.py:410,433,452).  Double-Q: rand, randint, rand
        # (algorithms.py:285,288,303) -> 2 "pairs", the second randint slot is zero padding.
        n, L = out["n_steps"], params["L"]
        pairs = len(out["rand"]) // max(n, 1)
        assert pairs * n == len(out["rand"])
        out["u"] = np.stack(out["rand"]).reshape(n, pairs, L, L)
        if len(out["randint"]) == len(out["rand"]):
            out["b"] = np.stack(out["randint"]).astype(np.uint8).reshape(n, pairs, L, L)
        else:
            assert pairs == 2 and len(out["randint"]) == n
            out["b"] = np.zeros((n, 2, L, L), np.uint8)
            out["b"][:, 0] = np.stack(out["randint"])
        for k in ("q1_0", "q2_0", "q1_final", "q2_final"):
            if k in out:
                extra[k] = out[k]
    blob = dict(params_json=np.array(json.dumps(params)), seed=np.array(seed),
                q0=out["q0"], s0=out["s0"].astype(np.uint8), u=out["u"], b=out["b"],
                q_final=out["q_final"], r_final=out["r_final"],
                s_final=out["s_final"].astype(np.uint8),
                ret=np.array(out["ret"], dtype=np.float64), **extra)
    for k in KEEP_DATASETS:
        if k in out["datasets"]:
            blob["ds_" + k] = out["datasets"][k]
    blob["dataset_names"] = np.array(sorted(out["datasets"].keys()))
    blob["dataset_shapes"] = np.array(
        json.dumps({k: [str(v.dtype), list(v.shape)] for k, v in out["datasets"].items()}))
    path = os.path.join(GOLDEN_DIR, f"replay_{name}.npz")
    np.savez_compressed(path, **blob)
    return path


def _band_worker(args):
    name, params, seed = args
    out = ref_harness.run_reference(seed, cluster_tail=False, **params)
    return name, seed, out["datasets"]["coop_rate_history"]


def make_bands(n_seeds, procs, cases=None, stride=1):
    cases = BAND_CASES if cases is None else cases
    jobs = [(name, p, 1000 + s) for name, p in cases.items() for s in range(n_seeds)]
    with mp.Pool(procs) as pool:
        res = pool.map(_band_worker, jobs)
    for name, p in cases.items():
        curves = np.stack([c for (n, s, c) in res if n == name])[:, ::stride]
        seeds = np.array([s for (n, s, c) in res if n == name])
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"band_{name}.npz"),
                            params_json=np.array(json.dumps(p)), seeds=seeds, stride=np.array(stride),
                            coop_rate_history=curves.astype(np.float32))
        if stride != 1:
            print(name, "f_c at the last stored iteration: mean", curves[:, -1].mean(), "sd", curves[:, -1].std())
            continue
        print(name, "f_c(t=10,100,300,1000) mean",
              curves[:, [9, 99, 299, 999]].mean(0), "sd", curves[:, [9, 99, 299, 999]].std(0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-band", action="store_true")
    ap.add_argument("--no-replay", action="store_true")
    ap.add_argument("--td", action="store_true", help="also (re)write the SARSA / Expected-SARSA fixtures")
    ap.add_argument("--long-band", action="store_true",
                    help="only (re)write the 10^4-iteration band of the default_config.yaml lattice")
    ap.add_argument("--seeds", type=int, default=8)
    ap.add_argument("--procs", type=int, default=max(1, (os.cpu_count() or 2) - 1))
    a = ap.parse_args()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    if a.long_band:
        make_bands(a.seeds, a.procs, LONG_BAND_CASES, LONG_STRIDE)
        return
    if not a.no_replay:
        for i, (name, p) in enumerate(REPLAY_CASES.items()):
            print("wrote", make_replay(name, p, 7 + i))
    if a.td:
        for i, (name, p) in enumerate(TD_CASES.items()):
            print("wrote", make_replay(name, p, 40 + i))
    if not a.no_band:
        make_bands(a.seeds, a.procs)


if __name__ == "__main__":
    main()
