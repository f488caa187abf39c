"""The drop-in class on the GPU: ``SPGG(**params).run(h5)`` against the golden runs of the
executed reference (same call pattern as src/experiments/runner.py:88-105)."""
import json
import os

import numpy as np
import pytest

from helpers import C1, C2, RUNNER_FIXED, load_golden

pytestmark = pytest.mark.gpu

GOLDEN = ["c1_rep_m1", "c2_act_m2", "rep_m2_r36", "act_m1_k0", "ctor_defaults", "odd_L_fracR"]


def _read(path):
    from spgg_b200 import h5lite
    with h5lite.open_file(path, "r") as f:
        return {k: np.array(f[k]) for k in f.keys()}


@pytest.mark.parametrize("name", GOLDEN)
def test_run_replays_the_reference_bit_for_bit(tmp_path, golden_dir, name):
    """precision='fp64', draws='numpy', seed=<the golden's seed>: the class consumes the same
    NumPy stream as the reference (ctor: uniform Q then randint S; per step rand then
    randint) and must reproduce its files: every dataset name/dtype/shape, S/R/Q bit for bit."""
    import spgg_b200
    z, p = load_golden(golden_dir, name)
    m = spgg_b200.SPGG(**p, seed=int(z["seed"]), precision="fp64", draws="numpy")
    assert np.array_equal(m.q_table, z["q0"]) and np.array_equal(m._Sn, z["s0"])
    m.folder = str(tmp_path)
    path = str(tmp_path / "run.h5")
    ret = m.run(path)
    assert os.path.isdir(tmp_path / "plots" / "snapshots")          # spgg.py:365-366
    assert np.array_equal(m._Sn, z["s_final"]) and m._Sn.dtype == np.int64
    assert np.array_equal(m.R, z["r_final"])
    assert np.array_equal(m.q_table, z["q_final"])
    np.testing.assert_allclose(ret, z["ret"], rtol=1e-12)
    got = _read(path)
    shapes = json.loads(str(z["dataset_shapes"]))
    assert sorted(got) == sorted(shapes)                             # the full dataset contract
    for k, (dt, shp) in shapes.items():
        assert str(got[k].dtype) == dt and list(got[k].shape) == shp, k
    for k in [f[3:] for f in z.files if f.startswith("ds_")]:
        want = z["ds_" + k]
        if want.dtype.kind == "i":
            assert np.array_equal(got[k], want), k
        else:
            np.testing.assert_allclose(got[k], want, rtol=1e-9, atol=1e-12, equal_nan=True, err_msg=k)
    assert m.kernel_launches > 0
    assert m.epsilon == pytest.approx(float(got["epsilon_history_final"][-1]))


def test_runner_call_pattern_default_mode(tmp_path):
    """runner.py:88-105 verbatim (shorter run): throughput mode, Philox draws, snapshots at the
    reference's iterations, series lengths, value ranges."""
    import spgg_b200
    spgg = spgg_b200.SPGG(
        r=3.0, c=1, cost=1, iterations=1200, L=100, num_of_strategies=2, K=0.1, population_type=0,
        alpha=0.8, gamma=0.9, epsilon=0.5, epsilon_decay=0.99, epsilon_min=0.01,
        influence_factor=1.0, use_second_order=False, lambda_epsilon=0.01,
        delta_R_C=1, delta_R_D=1, R_min=-10, R_max=10, reward_weight_payoff=0.95, rep_gain_C=1.0,
        state_representation="reputation", algorithm="qlearning")
    spgg.folder = str(tmp_path)
    record = spgg.run(str(tmp_path / "experiment_data.h5"))
    final_coop_ratio, final_def_ratio, _ = record
    assert 0 <= final_coop_ratio <= 1 and abs(final_coop_ratio + final_def_ratio - 1) < 1e-12
    final_rep_mean = spgg.rep_avg_history[-1] if spgg.rep_avg_history else 0   # runner.py:108
    assert final_rep_mean == 0
    d = _read(str(tmp_path / "experiment_data.h5"))
    T = 1200
    for k in ("coop_rate_history", "neighbor_influence_percent", "switch_C_to_D", "avg_q_s1_d_history"):
        assert d[k].shape == (T,)
    for i in (1, 10, 100, 1000):
        assert d[f"R_snapshot_{i}"].shape == (100, 100) and d[f"Sn_snapshot_{i}"].dtype == np.int64
        assert d[f"rep_hist_{i}"].sum() == 100 * 100
    assert "R_snapshot_5000" not in d
    assert np.array_equal(d["R_snapshot_1"], np.zeros((100, 100)))
    assert d["Sn_final"].dtype == np.int64 and set(np.unique(d["Sn_final"])) <= {0, 1}
    assert d["coop_rate_history"][-1] == pytest.approx((spgg._Sn == 0).mean(), abs=0.05)
    assert d["cluster_sizes"].sum() == (d["Sn_final"] == 0).sum()
    assert d["epsilon_history_final"][-1] == 0.01
    assert np.all((d["group_comp_d0_history"] >= 0) & (d["group_comp_d0_history"] <= 100))


@pytest.mark.parametrize("cfg", ["c1", "c2"])
def test_philox_stream_stays_in_the_reference_seed_band(golden_dir, cfg):
    """north_star: with the native Philox stream the long-run cooperation-rate curve must fall
    within the reference's seed-to-seed band (8 reference seeds, L=200, 1000 iterations;
    band = mean +- max(4 sd, 0.02) at t = 100, 300, 1000)."""
    import spgg_b200
    z = np.load(os.path.join(golden_dir, f"band_{cfg}.npz"))
    p = json.loads(str(z["params_json"]))
    curves = z["coop_rate_history"].astype(np.float64)
    mean, sd = curves.mean(0), curves.std(0)
    T = curves.shape[1]
    for seed in (1, 2, 3):
        eng = spgg_b200.Engine(p, seeds=seed, precision="fp32")
        rs = np.random.RandomState(seed)
        L = p["L"]
        eng.set_state(rs.randint(0, 2, (L, L)), np.zeros((L, L)), rs.uniform(-0.01, 0.01, (L, L, 2, 2)))
        eng.step(T)
        rows = eng.stats()[1:]
        fc = rows[:, 0] / (L * L)
        for t in (99, 299, 999):
            tol = max(4 * sd[t], 0.02)
            assert abs(fc[t] - mean[t]) <= tol, (cfg, seed, t, fc[t], mean[t], sd[t])
        eng.close()


def test_philox_long_run_stays_in_the_reference_seed_band(golden_dir):
    """north_star "long-run": the default_config.yaml lattice (L=100, C1 physics) for 10^4
    iterations, 8 Philox seeds against the band of 8 runs of the executed reference
    (tests/golden/band_c1_long.npz, every 10th entry of coop_rate_history).  Checked: the curve at
    t = 10^3, 3*10^3, 9990 within mean +- max(4 sd, 0.02), and the time average over the last 2000
    iterations (every 10th) within mean +- max(4 sd, 0.01) of the reference's."""
    import spgg_b200
    z = np.load(os.path.join(golden_dir, "band_c1_long.npz"))
    p = json.loads(str(z["params_json"]))
    stride = int(z["stride"])
    curves = z["coop_rate_history"].astype(np.float64)       # (seeds, T/stride)
    mean, sd = curves.mean(0), curves.std(0)
    T, L = p["iterations"], p["L"]
    tail_ref = curves[:, -2000 // stride:].mean(1)
    tails = []
    for seed in range(11, 19):
        eng = spgg_b200.Engine(p, seeds=seed, precision="fp32")
        rs = np.random.RandomState(seed)
        eng.set_state(rs.randint(0, 2, (L, L)), np.zeros((L, L)), rs.uniform(-0.01, 0.01, (L, L, 2, 2)))
        eng.step(T)
        rows = eng.stats()
        # coop_rate_history[t] = cooperators of S_t (spgg.py:383): column NC_OLD of iteration t+1
        fc = rows[1:, 0] / (L * L)
        assert len(fc) == T == curves.shape[1] * stride
        for t in (1000, 3000, 9990):
            tol = max(4 * sd[t // stride], 0.02)
            assert abs(fc[t] - mean[t // stride]) <= tol, (seed, t, fc[t], mean[t // stride], sd[t // stride])
        tails.append(fc[-2000::stride].mean())
        eng.close()
    tol = max(4 * tail_ref.std(), 0.01)
    assert abs(np.mean(tails) - tail_ref.mean()) <= tol, (np.mean(tails), tail_ref.mean(), tail_ref.std())


def test_foreign_algorithm_subclass_fails_loudly(tmp_path):
    """The reference accepts any RLAlgorithm instance (spgg.py:115-116); a rule that is not one of
    its four cannot run inside the fused kernel and must say so (no CPU fallback)."""
    import spgg_b200

    class MyRule(spgg_b200.RLAlgorithm):
        name = "my_rule"

    m = spgg_b200.SPGG(L=16, iterations=3, algorithm=MyRule(0.1, 0.9, 0.5, 0.99, 0.01), seed=1)
    with pytest.raises(ValueError, match="no CPU fallback"):
        m.run(str(tmp_path / "x.h5"))


@pytest.mark.parametrize("name", ["doubleq_rep_m1", "doubleq_act_m2"])
def test_double_q_replays_the_reference(tmp_path, golden_dir, name):
    """algorithms.py:237-341 + spgg.py:464-468,498-505: both tables, the combined table, S and R
    bit for bit with the reference's own stream (rand, randint, rand per iteration)."""
    import spgg_b200
    z, p = load_golden(golden_dir, name)
    m = spgg_b200.SPGG(**p, seed=int(z["seed"]), precision="fp64", draws="numpy")
    assert np.array_equal(m.algorithm.q_table_1, z["q1_0"]) and np.array_equal(m.algorithm.q_table_2, z["q2_0"])
    assert np.array_equal(m.q_table, z["q0"]) and np.array_equal(m._Sn, z["s0"])
    m.folder = str(tmp_path)
    ret = m.run(str(tmp_path / "run.h5"))
    assert np.array_equal(m._Sn, z["s_final"]) and np.array_equal(m.R, z["r_final"])
    assert np.array_equal(m.algorithm.q_table_1, z["q1_final"])
    assert np.array_equal(m.algorithm.q_table_2, z["q2_final"])
    assert np.array_equal(m.q_table, z["q_final"])
    np.testing.assert_allclose(ret, z["ret"], rtol=1e-12)
    got = _read(str(tmp_path / "run.h5"))
    for k in [f[3:] for f in z.files if f.startswith("ds_")]:
        want = z["ds_" + k]
        if want.dtype.kind == "i":
            assert np.array_equal(got[k], want), k
        else:
            np.testing.assert_allclose(got[k], want, rtol=1e-9, atol=1e-12, equal_nan=True, err_msg=k)


TD_GOLDEN = ["sarsa_rep_m1", "sarsa_act_m2", "expsarsa_rep_m2", "expsarsa_act_m1"]


@pytest.mark.parametrize("name", TD_GOLDEN)
def test_sarsa_and_expected_sarsa_replay_the_reference(tmp_path, golden_dir, name):
    """algorithms.py:136-234 as applied by spgg.py:431-473: fp64 instantiation + the reference's
    own draw stream (SARSA: three pairs per iteration) == the reference run, bit for bit."""
    import spgg_b200
    z, p = load_golden(golden_dir, name)
    m = spgg_b200.SPGG(**p, seed=int(z["seed"]), precision="fp64", draws="numpy")
    m.folder = str(tmp_path)
    ret = m.run(str(tmp_path / "run.h5"))
    assert np.array_equal(m._Sn, z["s_final"])
    assert np.array_equal(m.R, z["r_final"])
    assert np.array_equal(m.q_table, z["q_final"])
    np.testing.assert_allclose(ret, z["ret"], rtol=1e-12)
    got = _read(str(tmp_path / "run.h5"))
    for k in [f[3:] for f in z.files if f.startswith("ds_")]:
        want = z["ds_" + k]
        if want.dtype.kind == "i":
            assert np.array_equal(got[k], want), k
        else:
            np.testing.assert_allclose(got[k], want, rtol=1e-9, atol=1e-12, equal_nan=True, err_msg=k)


@pytest.mark.parametrize("algo", ["sarsa", "expected_sarsa"])
def test_td_rules_engine_vs_numpy_oracle(algo):
    """Direct engine check at a larger size against the NumPy restatement (pinned to the
    reference by the fixtures above): L=48, 60 iterations, both neighbour orders."""
    import spgg_b200
    from oracle import spgg_numpy
    from helpers import full_params
    for second, state in ((False, "reputation"), (True, "action")):
        L, n = 48, 60
        p = full_params(dict(C1, L=L, use_second_order=second, state_representation=state, algorithm=algo,
                             iterations=n))
        rs = np.random.RandomState(3)
        Q0 = rs.uniform(-0.01, 0.01, (L, L, 2, 2))
        S0 = rs.randint(0, 2, (L, L))
        pairs = 3 if algo == "sarsa" else 1
        u = rs.rand(n, pairs, L, L)
        b = rs.randint(0, 2, (n, pairs, L, L)).astype(np.uint8)
        eng = spgg_b200.Engine(p, precision="fp64")
        eng.set_state(S0, np.zeros((L, L)), Q0)
        eng.set_replay(u if pairs > 1 else u[:, 0], b if pairs > 1 else b[:, 0])
        eng.step(n)
        S, R, Q = eng.get_state()
        draws = (lambda t, L_: tuple(x for k in range(3) for x in (u[t - 1, k], b[t - 1, k]))) if pairs == 3 \
            else (lambda t, L_: (u[t - 1, 0], b[t - 1, 0]))
        ref = spgg_numpy.simulate(p, S0, np.zeros((L, L)), Q0, draws)
        assert np.array_equal(S, ref["Sn_final"]) and np.array_equal(R, ref["R_final"])
        assert np.array_equal(Q, ref["q_final"])
        rows = eng.stats()[1:]
        np.testing.assert_allclose(rows[:, 30] / (L * L), ref["neighbor_influence_percent"], rtol=1e-9)
        eng.close()


@pytest.mark.parametrize("algo", ["sarsa", "expected_sarsa", "double_qlearning"])
def test_td_rules_throughput_mode_runs_and_is_reproducible(tmp_path, algo):
    """fp32 + Philox (streams 0/1/2 for SARSA) through the class: same seed -> same files."""
    import spgg_b200
    outs = []
    for k in range(2):
        m = spgg_b200.SPGG(L=96, iterations=150, r=3.0, cost=1, alpha=0.8, epsilon_decay=0.99,
                           use_second_order=False, reward_weight_payoff=0.95, rep_gain_C=1.0,
                           algorithm=algo, seed=11)
        m.folder = str(tmp_path / str(k))
        m.run(str(tmp_path / f"{k}.h5"))
        outs.append(_read(str(tmp_path / f"{k}.h5")))
    for key in ("Sn_final", "R_final", "coop_rate_history", "switch_C_to_D"):
        assert np.array_equal(outs[0][key], outs[1][key]), key
    fc = outs[0]["coop_rate_history"]
    assert len(fc) == 150 and 0.0 < fc[-1] < 1.0


def test_gpu_runner_batches_equal_single_experiments(tmp_path):
    """run_experiments (batched replicas) == run_one_experiment per tuple; folder layout and
    file names of runner.py:74-85,104."""
    from spgg_b200 import runner
    combos = [(3.0, 1.0, False, 0.8, 0.95, 1.0, "reputation", "qlearning"),
              (4.0, 0.5, False, 0.8, 0.95, 1.0, "reputation"),
              (4.0, 1.0, True, 0.8, 1.0, 1.0, "action", "qlearning"),
              (3.6, 0.0, False, 0.8, 1.0, 0.5)]
    kw = dict(L=64, iterations=120, seed=5)
    res = runner.run_experiments(combos, num_processes=2, use_progress_bar=False,
                                 base_dir=str(tmp_path / "batch"), **kw)
    assert [p for p, _ in res] == combos
    for (p, (coop, rep_mean)) in res:
        single_p, (coop1, rep1) = runner.run_one_experiment(p, base_dir=str(tmp_path / "single"), **kw)
        assert coop == coop1 and rep_mean == rep1 == 0
        folder = runner.get_folder_name(*runner._unpack(p))
        a = _read(str(tmp_path / "batch" / folder / "data" / "experiment_data.h5"))
        b = _read(str(tmp_path / "single" / folder / "data" / "experiment_data.h5"))
        assert sorted(a) == sorted(b)
        for k in ("Sn_final", "R_final", "coop_rate_history", "switch_C_to_D", "Sn_snapshot_100",
                  "R_snapshot_10", "rep_hist_final", "cluster_sizes"):
            assert np.array_equal(a[k], b[k]), k
        for sub in ("configurations", "reputations", os.path.join("plots", "snapshots"), "data"):
            assert os.path.isdir(tmp_path / "batch" / folder / sub)
