"""Host-side mirror of the reference's class API (src/model/spgg.py:50-164,
src/model/algorithms.py): constructor signature, attributes, error behaviour.  No GPU."""
import inspect

import numpy as np
import pytest

import spgg_b200
from spgg_b200 import SPGG, series

# spgg.py:50-56, verbatim order and defaults
REF_SIGNATURE = [
    ("r", 2), ("c", 1), ("cost", 0.5), ("K", 0.1), ("L", 50), ("iterations", 1000),
    ("num_of_strategies", 2), ("population_type", 0), ("S_in_one", None), ("alpha", 0.1),
    ("gamma", 0.9), ("epsilon", 0.5), ("epsilon_decay", 0.995), ("epsilon_min", 0.01),
    ("influence_factor", 1.0), ("use_second_order", True), ("lambda_epsilon", 0.01),
    ("delta_R_C", 1), ("delta_R_D", 1), ("R_min", -10), ("R_max", 10),
    ("reward_weight_payoff", 1.0), ("rep_gain_C", 0.5), ("state_representation", "reputation"),
    ("algorithm", "qlearning")]


def test_constructor_signature_is_the_reference_one():
    sig = inspect.signature(SPGG.__init__)
    params = [p for p in sig.parameters.values() if p.name != "self"]
    assert params[-1].kind is inspect.Parameter.VAR_KEYWORD and params[-1].name == "params"
    got = [(p.name, p.default) for p in params[:-1]]
    assert got == REF_SIGNATURE


def test_attributes_and_ctor_draw_order():
    """Every ctor argument becomes an attribute and a ``params`` entry (spgg.py:101-105);
    with a pinned seed the ctor draws are the reference's: uniform Q then randint S
    (spgg.py:121,162)."""
    m = SPGG(r=3.0, L=12, iterations=7, alpha=0.8, reward_weight_payoff=0.95, seed=42, foo="bar")
    assert m.r == 3.0 and m.L == 12 and m.alpha == 0.8 and m.foo == "bar"
    assert m.params["reward_weight_payoff"] == 0.95 and m.params["foo"] == "bar"
    assert m.reward_weight_rep == 1 - 0.95
    rs = np.random.RandomState(42)
    q0 = rs.uniform(-0.01, 0.01, (12, 12, 2, 2))
    s0 = rs.randint(0, 2, (12, 12))
    assert np.array_equal(m.q_table, q0) and np.array_equal(m._Sn, s0)
    assert np.array_equal(m.R, np.zeros((12, 12)))
    assert m._S[0].sum() + m._S[1].sum() == 144
    assert m.normlize_max == 12.0 and m.normlize_min == -2.0
    assert m.snapshot_iters == {1, 10, 100, 1000, 5000, 10000, 20000, 30000, 40000}
    assert m.folder is None
    assert m.track_positions == [(6, 6), (3, 3), (9, 9)]
    assert m.algorithm.alpha == 0.8 and m.algorithm.epsilon == 0.5


def test_S_in_one_injection():
    S = np.eye(8, dtype=int)
    m = SPGG(L=8, S_in_one=S, seed=0)
    assert np.array_equal(m._Sn, S)
    assert np.array_equal(m._S[1], S) and np.array_equal(m._S[0], 1 - S)


def test_errors_match_the_reference():
    with pytest.raises(ValueError, match="Unknown algorithm"):
        SPGG(L=8, algorithm="nope")                       # algorithms.py:382
    with pytest.raises(ValueError, match="must be str or RLAlgorithm"):
        SPGG(L=8, algorithm=3)                            # spgg.py:118
    m = SPGG(L=8, state_representation="bogus", seed=0)   # reference raises from run(), spgg.py:309
    with pytest.raises(ValueError, match="Unknown state_representation"):
        m.run("/dev/null")
    with pytest.raises((ValueError, OverflowError)):
        SPGG(L=8, iterations=0)                           # log10(0), spgg.py:152


def test_algorithm_interface():
    from spgg_b200 import (RLAlgorithm, QLearning, SARSA, ExpectedSARSA, DoubleQLearning,
                           create_algorithm)
    a = create_algorithm("QLearning", 0.1, 0.9, 0.5, 0.99, 0.01)
    assert isinstance(a, QLearning) and isinstance(a, RLAlgorithm)
    for _ in range(3):
        a.decay_epsilon()
    assert a.epsilon == max(max(max(0.5 * 0.99, 0.01) * 0.99, 0.01) * 0.99, 0.01)
    for name, cls in (("sarsa", SARSA), ("expected_sarsa", ExpectedSARSA),
                      ("double_qlearning", DoubleQLearning), ("q-learning", QLearning)):
        assert isinstance(create_algorithm(name, 0.1, 0.9, 0.5, 0.99, 0.01), cls)
    inst = QLearning(0.3, 0.9, 0.4, 0.99, 0.01)
    m = SPGG(L=8, algorithm=inst, seed=0)
    assert m.algorithm is inst
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        inst.select_action(np.zeros((8, 8, 2, 2)), np.zeros((8, 8), int), 8)


def test_epsilon_series_is_an_iterated_product():
    e = series.epsilon_after(0.5, 0.99, 0.01, 500)
    x = 0.5
    for t in range(500):
        x = max(x * 0.99, 0.01)
        assert e[t] == x
    assert e[-1] == 0.01


def test_uniform_payoff_matches_numpy_oracle():
    from oracle import spgg_numpy
    for allc in (True, False):
        S = np.zeros((6, 6), np.int64) if allc else np.ones((6, 6), np.int64)
        P = spgg_numpy.normalised_payoff(S, 3.6, 1, 1)
        assert series.uniform_payoff(allc, 3.6, 1, 1) == P[0, 0]


def test_assemble_early_exit_lengths():
    """T completed iterations then a break on a uniform lattice: T_c = T + 1 (spgg.py:405)."""
    T, N = 4, 64
    rows = np.zeros((T, spgg_b200.NSTAT))
    rows[:, 0] = [30, 40, 50, 60]
    ser = series.assemble(rows, np.zeros(T), N, dict(r=3.0, c=1, cost=1), 0.5, stopped=True,
                          stop_sum_r=64.0, stop_all_coop=True)
    assert ser["coop_rate_history"].shape == (T + 1,) and ser["coop_rate_history"][-1] == 1.0
    assert ser["it_records_final"].shape == (T + 1, 6)
    assert ser["rep_avg_history_final"][-1] == 1.0
    assert ser["epsilon_history_final"].shape == (T,)
    assert ser["switch_C_to_D"].dtype == np.int64


def test_runner_folder_names_and_tuple_formats():
    """src/experiments/runner.py:11-45, 62-72."""
    from spgg_b200 import runner
    assert runner.get_folder_name(3.0, 1.0, False, 0.8, 0.95, 1.0) == \
        "results_r3.0_inf1.0_orderFalse_alpha0.8_rw0.95_rgC1.00"
    assert runner.get_folder_name(4.0, 0.5, True, 0.8, 1.0, 0.5, "action", "sarsa") == \
        "results_r4.0_inf0.5_orderTrue_alpha0.8_rw1.00_rgC0.50_action_sarsa"
    assert runner._unpack((3.0, 1.0, False, 0.8, 0.95, 1.0)) == \
        (3.0, 1.0, False, 0.8, 0.95, 1.0, "reputation", "qlearning")
    assert runner._unpack((3.0, 1.0, False, 0.8, 0.95, 1.0, "action"))[-2:] == ("action", "qlearning")
    with pytest.raises(ValueError):
        runner._unpack((1, 2, 3))
    assert runner.RUNNER_MODEL["L"] == 100 and runner.RUNNER_MODEL["iterations"] == 100001


def test_epsilon_series_matches_the_iterated_rule():
    """series.epsilon_after vectorises e <- max(e*decay, eps_min) (algorithms.py:40-42); it must
    reproduce the iterated roundings exactly, including degenerate parameters."""
    import numpy as np
    from spgg_b200 import series

    def ref(e, d, m, n):
        out = []
        for _ in range(n):
            e = max(e * d, m)
            out.append(e)
        return np.array(out)

    cases = [(0.5, 0.99, 0.01, 2000), (0.5, 0.995, 0.01, 30000), (0.5, 1.0, 0.01, 50), (0.0, 0.99, 0.0, 30),
             (0.5, 0.99, 0.5, 10), (0.3, 0.9, 0.0, 500), (0.5, 1.5, 0.01, 20), (0.5, 0.0, 0.01, 5),
             (0.005, 0.99, 0.01, 8), (0.5, 0.99, 0.01, 0)]
    for e, d, m, n in cases:
        assert np.array_equal(series.epsilon_after(e, d, m, n), ref(e, d, m, n)), (e, d, m, n)


def test_host_worker_runs_jobs_in_order_and_reraises():
    """The snapshot post-processing thread of SPGG.run: batches run in submission order, one at a
    time, and an exception in a job resurfaces on the caller's thread."""
    import pytest
    from spgg_b200.spgg import _HostWorker
    out = []
    w = _HostWorker()
    w.run([lambda: out.append(1), lambda: out.append(2)])
    w.run([lambda: out.append(3)])          # joins the first batch before starting the second
    w.join()
    assert out == [1, 2, 3]
    w.run([])                               # nothing to do: no thread
    w.join()

    def boom():
        raise ValueError("disk full")
    w.run([boom, lambda: out.append(4)])
    with pytest.raises(ValueError, match="disk full"):
        w.join()
    assert out == [1, 2, 3]                 # jobs after the failing one are not run
    w.join()                                # the error is reported once
