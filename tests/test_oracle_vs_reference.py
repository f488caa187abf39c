"""Container-only pin: the unmodified reference (``/root/reference``, executed through
``oracle/ref_harness.py`` with h5py/matplotlib stubs) against the oracle restatements and
against the committed golden fixtures.  Skipped where the reference is absent (GPU box)."""
import json
import os

import numpy as np
import pytest

from helpers import RUNNER_FIXED, full_params, load_golden

from oracle import ref_harness

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_harness.reference_available(),
                                 reason="/root/reference not present")]


@pytest.mark.parametrize("state,second,r,kappa,wP", [
    ("reputation", False, 3.0, 1.0, 0.95),
    ("action", True, 4.0, 1.0, 1.0),
    ("reputation", True, 3.6, 0.5, 1.0),
])
def test_reference_run_equals_oracles(state, second, r, kappa, wP):
    from oracle import c_oracle, spgg_numpy
    L, n, seed = 14, 30, 321
    params = dict(RUNNER_FIXED, L=L, iterations=n, r=r, influence_factor=kappa,
                  use_second_order=second, reward_weight_payoff=wP, rep_gain_C=1.0,
                  state_representation=state)
    ref = ref_harness.run_reference(seed, **params)
    assert ref["n_steps"] == n
    draws = lambda t, L_: (ref["u"][t - 1], ref["b"][t - 1])
    p = full_params(params)
    out = spgg_numpy.simulate(p, ref["s0"], ref["r0"], ref["q0"], draws)
    assert np.array_equal(out["Sn_final"], ref["s_final"])
    assert np.array_equal(out["R_final"], ref["r_final"])
    assert np.array_equal(out["q_final"], ref["q_final"])
    for key in ("coop_rate_history", "switch_C_to_D", "neighbor_influence_percent",
                "best_neighbor_second_order_percent", "group_comp_d2_history",
                "cooperators_q_s1_c_history", "avg_q_s0_d_history", "reputation_reward_ratio"):
        assert np.array_equal(out[key], ref["datasets"][key], equal_nan=True), key
    sim = c_oracle.Sim(p, ref["s0"], ref["r0"], ref["q0"], "fp64")
    sim.run(n, draws)
    assert np.array_equal(sim.S, ref["s_final"])
    assert np.array_equal(sim.R, ref["r_final"])
    assert np.array_equal(sim.Q, ref["q_final"])


def test_reconstructed_legacy_stream_is_what_the_reference_consumes():
    from oracle import spgg_numpy
    L, n, seed = 10, 6, 99
    params = dict(RUNNER_FIXED, L=L, iterations=n, r=3.0, influence_factor=1.0,
                  use_second_order=False, reward_weight_payoff=0.95, rep_gain_C=1.0)
    ref = ref_harness.run_reference(seed, **params)
    Q0, S0, draws = spgg_numpy.legacy_draws(seed, L)
    assert np.array_equal(Q0, ref["q0"]) and np.array_equal(S0, ref["s0"])
    for t in range(1, n + 1):
        u, b = draws(t, L)
        assert np.array_equal(u, ref["u"][t - 1]) and np.array_equal(b, ref["b"][t - 1])


def test_committed_golden_fixture_is_reproducible(golden_dir):
    """Re-run the reference for one fixture and compare with what is committed."""
    z, params = load_golden(golden_dir, "c1_rep_m1")
    ref = ref_harness.run_reference(int(z["seed"]), **params)
    assert np.array_equal(ref["q_final"], z["q_final"])
    assert np.array_equal(ref["s_final"], z["s_final"])
    assert np.array_equal(ref["u"], z["u"])
    shapes = json.loads(str(z["dataset_shapes"]))
    assert sorted(shapes) == sorted(ref["datasets"])


def test_dropin_class_has_the_reference_signature_and_exports():
    import inspect
    import spgg_b200
    ref_model = ref_harness.import_reference()
    ref_sig = inspect.signature(ref_model.SPGG.__init__)
    our_sig = inspect.signature(spgg_b200.SPGG.__init__)
    assert [(p.name, p.default, p.kind) for p in ref_sig.parameters.values()] == \
           [(p.name, p.default, p.kind) for p in our_sig.parameters.values()]
    for name in ("SPGG", "RLAlgorithm", "QLearning", "SARSA", "ExpectedSARSA", "DoubleQLearning",
                 "create_algorithm"):
        assert hasattr(ref_model, name) and hasattr(spgg_b200, name)
    # attribute surface after construction
    with ref_harness.pinned_seed(5):
        a = ref_model.SPGG(L=10, iterations=3, r=3.0)
    b = spgg_b200.SPGG(L=10, iterations=3, r=3.0, seed=5)
    assert np.array_equal(a.q_table, b.q_table) and np.array_equal(a._Sn, b._Sn)
    for attr in ("R", "params", "folder", "snapshot_iters", "track_positions", "normlize_max",
                 "normlize_min", "reward_weight_rep", "algorithm", "cache", "it_records",
                 "epsilon_history", "rep_avg_history", "q_history"):
        assert hasattr(a, attr) and hasattr(b, attr), attr
    for k, v in a.params.items():
        if k in ("S_in_one",):
            continue
        assert k in b.params, k


def test_runner_folder_name_equals_reference():
    import importlib
    from spgg_b200 import runner
    ref_harness.import_reference()
    ref_runner = importlib.import_module("src.experiments.runner")
    for args in [(3.0, 1.0, False, 0.8, 0.95, 1.0), (3.6, 0.0, True, 0.1, 1.0, 0.5, "action"),
                 (5, 2, True, 0.8, 0.9, 0.25, "reputation", "double_qlearning")]:
        assert runner.get_folder_name(*args) == ref_runner.get_folder_name(*args)


def test_double_q_ctor_draw_order_equals_reference():
    """spgg.py:121-127: uniform q_table (discarded), then table 1, table 2, then randint S."""
    import spgg_b200
    ref_model = ref_harness.import_reference()
    with ref_harness.pinned_seed(9):
        a = ref_model.SPGG(L=10, iterations=3, algorithm="double_qlearning")
    b = spgg_b200.SPGG(L=10, iterations=3, algorithm="double_qlearning", seed=9)
    assert np.array_equal(a.algorithm.q_table_1, b.algorithm.q_table_1)
    assert np.array_equal(a.algorithm.q_table_2, b.algorithm.q_table_2)
    assert np.array_equal(a.q_table, b.q_table) and np.array_equal(a._Sn, b._Sn)
