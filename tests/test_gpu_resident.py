"""GPU parity tests of the lattice-resident cluster kernel (csrc/spgg_resident.cuh): one
thread-block cluster keeps a small lattice in shared memory for a whole chunk of iterations.
Same arithmetic and Philox counters as the per-iteration kernels, so strategies, reputations
and Q-tables must be bit-identical to them and to the C oracle."""
import numpy as np
import pytest

from helpers import C1, C2, full_params

pytestmark = pytest.mark.gpu

EXACT = [c for c in range(18) if c != 10] + [31, 32, 33]


def _engine(*a, **k):
    import spgg_b200
    return spgg_b200.Engine(*a, **k)


def _init(L, seed):
    rs = np.random.RandomState(seed)
    Q0 = rs.uniform(-0.01, 0.01, (L, L, 2, 2)).astype(np.float32).astype(np.float64)
    S0 = rs.randint(0, 2, (L, L))
    return S0, Q0


CASES = [
    ("rep_m1_L100", dict(C1, L=100)),                                   # config 0/1 geometry: 8 blocks of 12-13 rows
    ("act_m2_L200", dict(C2, L=200)),                                   # config 2 geometry
    ("rep_m2_L36", dict(C1, L=36, use_second_order=True, r=3.6)),       # uneven blocks (4 of 5 rows, 4 of 4)
    ("rep_m2_L10", dict(C1, L=10, use_second_order=True)),              # one-CTA cluster, ghosts are its own rows
    ("act_m1_L102", dict(C2, L=102, use_second_order=False, r=3.0)),    # L % 4 != 0: partial last quad
    ("rep_m1_L17", dict(C1, L=17, rep_gain_C=0.5)),                     # odd side, blocks of 2-3 rows, half units
    ("rep_m2_L128", dict(C1, L=128, use_second_order=True, influence_factor=0.5)),
    ("rep_m1_L240", dict(C1, L=240, reward_weight_payoff=0.9)),         # near the shared-memory limit of 8 blocks
    ("act_m2_L256", dict(C2, L=256)),                                   # only fits as 16 blocks
    ("rep_m2_L336", dict(C1, L=336, use_second_order=True)),            # 16 blocks of 21 rows, near the limit
    # grid mode: one lattice over a cooperative grid of 148 CTAs, ghost rows through L2
    ("grid_rep_m1_L400", dict(C1, L=400)),
    ("grid_act_m2_L512", dict(C2, L=512)),                              # would otherwise take the TMA fast path
    ("grid_rep_m2_L362", dict(C1, L=362, use_second_order=True)),       # L % 4 != 0, blocks of 2-3 rows
    ("grid_rep_m2_L1000", dict(C1, L=1000, use_second_order=True, r=3.6)),
    ("grid_act_m1_L1036", dict(C2, L=1036, use_second_order=False)),    # 148 blocks of 7 rows: the largest lattice that fits
]


def test_cluster_of_8_equals_cluster_of_16(monkeypatch):
    """A lone lattice is spread over a (non-portable) cluster of 16 CTAs, batches use clusters
    of 8: the decomposition must be invisible (statistics rows included: integer columns
    exact, fp32 sums grouped differently)."""
    L, n = 200, 40
    p = full_params(dict(C1, L=L, use_second_order=True))
    S0, Q0 = _init(L, 23)
    outs = []
    for cs8 in (False, True):
        if cs8:
            monkeypatch.setenv("SPGG_RES_CS8", "1")
        else:
            monkeypatch.delenv("SPGG_RES_CS8", raising=False)
        eng = _engine(p, seeds=5, precision="fp32")
        eng.set_state(S0, np.zeros((L, L)), Q0)
        eng.step(n)
        outs.append(eng.get_state() + (eng.stats(),))
        eng.close()
    for a, b in zip(*outs):
        if a.ndim == 2 and a.shape[1] == 40:
            assert np.array_equal(a[:, EXACT], b[:, EXACT])
            np.testing.assert_allclose(a[:, 18:31], b[:, 18:31], rtol=1e-5, atol=1e-4)
        else:
            assert np.array_equal(a, b)


@pytest.mark.parametrize("name,p", CASES, ids=[c[0] for c in CASES])
def test_resident_equals_per_iteration_kernels(monkeypatch, name, p):
    """Resident cluster kernel vs the two-launches-per-iteration path: S, R, Q and every
    integer statistic bit-identical; fp32 partial sums are grouped differently."""
    p = full_params(p)
    L, n = p["L"], 45
    S0, Q0 = _init(L, 17)
    outs, launches = [], []
    for no_res in (False, True):
        if no_res:
            monkeypatch.setenv("SPGG_NO_RESIDENT", "1")
        else:
            monkeypatch.delenv("SPGG_NO_RESIDENT", raising=False)
        eng = _engine(p, seeds=777, precision="fp32")
        want = ("cooperative grid" if name.startswith("grid_") else "resident: cluster") if not no_res else "per iteration"
        assert want in eng.describe(), eng.describe()
        eng.set_state(S0, np.zeros((L, L)), Q0)
        l0 = eng.status().kernel_launches
        eng.step(n)
        launches.append(eng.status().kernel_launches - l0)
        outs.append(eng.get_state() + (eng.stats(),))
        eng.close()
    assert launches[0] == 2 and launches[1] >= n        # the resident path really ran: one launch per chunk (+ the row kernel)
    for a, b in zip(*outs):
        if a.ndim == 2 and a.shape[1] == 40:
            assert np.array_equal(a[:, EXACT], b[:, EXACT])
            # fp32 partial sums over up to 10^6 sites, grouped differently by the two layouts
            tol = dict(rtol=1e-5, atol=1e-4) if L <= 400 else dict(rtol=2e-4, atol=1e-3)
            np.testing.assert_allclose(a[:, 10], b[:, 10], **tol)
            np.testing.assert_allclose(a[:, 18:31], b[:, 18:31], **tol)
        else:
            assert np.array_equal(a, b)


@pytest.mark.parametrize("cfg", ["c1", "c2"])
def test_resident_3000_steps_vs_oracle(cfg):
    """Long run in a few chunks against the C restatement of the throughput arithmetic."""
    from oracle import c_oracle
    L, n = (100, 3000) if cfg == "c1" else (64, 3000)
    p = full_params(dict(C1 if cfg == "c1" else C2, L=L))
    S0, Q0 = _init(L, 5)
    eng = _engine(p, seeds=99, precision="fp32")
    eng.set_state(S0, np.zeros((L, L)), Q0)
    rows = []
    for c in (1, 999, 2000):
        eng.step(c)
        rows.append(eng.stats()[1:])
    rows = np.vstack(rows)
    S, R, Q = eng.get_state()
    sim = c_oracle.Sim(p, S0, np.zeros((L, L)), Q0, "fp32", seed=99)
    rows_o = sim.run(n)
    assert np.array_equal(S, sim.S)
    assert np.array_equal(R, sim.R.astype(np.float64))
    assert np.array_equal(Q.astype(np.float32), sim.Q)
    ints = [0, 1, 2, 3, 11, 12, 13, 14, 15, 16, 31, 32]
    assert np.array_equal(rows[:, ints], rows_o[:, ints])
    assert np.array_equal(rows[:, 33].astype(np.float32), rows_o[:, 33].astype(np.float32))
    np.testing.assert_allclose(rows[:, 4:11], rows_o[:, 4:11], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(rows[:, 18:31], rows_o[:, 18:31], rtol=1e-4, atol=1e-3)
    assert eng.status().iteration == n
    eng.close()


def test_resident_many_replicas_more_clusters_than_fit():
    """40 replicas of L=200 = 40 clusters of 8 CTAs: more than one wave of clusters; each
    replica equals its own single run."""
    L, n = 200, 12
    plist = [full_params(dict(C1, L=L, r=3.0 + 0.05 * i, influence_factor=(i % 3) * 0.5)) for i in range(40)]
    rs = np.random.RandomState(8)
    init = [(rs.randint(0, 2, (L, L)), rs.uniform(-0.01, 0.01, (L, L, 2, 2))) for _ in plist]
    eng = _engine(plist, seeds=list(range(500, 540)), precision="fp32")
    for i, (S0, Q0) in enumerate(init):
        eng.set_state(S0, np.zeros((L, L)), Q0, replica=i)
    eng.step(n)
    for i in (0, 7, 19, 39):
        S0, Q0 = init[i]
        single = _engine(plist[i], seeds=500 + i, precision="fp32")
        single.set_state(S0, np.zeros((L, L)), Q0)
        single.step(n)
        for a, b in zip(eng.get_state(i), single.get_state()):
            assert np.array_equal(a, b)
        # the lone lattice runs on a cluster of 16, the batch on clusters of 8: integer columns
        # exact, fp32 partial sums grouped differently
        ra, rb = eng.stats(i), single.stats()
        assert np.array_equal(ra[:, EXACT], rb[:, EXACT])
        np.testing.assert_allclose(ra[:, 18:31], rb[:, 18:31], rtol=1e-5, atol=1e-4)
        single.close()
    eng.close()


def test_resident_early_exit(monkeypatch):
    """spgg.py:405 inside the resident loop: every block must leave the loop at the same
    iteration, and a stopped replica must not disturb its batch neighbours."""
    L = 24
    p = full_params(dict(C1, L=L, epsilon=0.0, epsilon_min=0.0))
    Q0 = np.zeros((L, L, 2, 2))
    Q0[..., 1] = 1.0                      # everybody prefers to defect -> all D after iteration 1
    S0 = np.random.RandomState(0).randint(0, 2, (L, L))
    S1, Q1 = _init(L, 3)
    p_run = full_params(dict(C1, L=L))
    res = []
    for no_res in (False, True):
        if no_res:
            monkeypatch.setenv("SPGG_NO_RESIDENT", "1")
        else:
            monkeypatch.delenv("SPGG_NO_RESIDENT", raising=False)
        eng = _engine([p, p_run], seeds=[1, 2], precision="fp32")
        eng.set_state(S0, np.zeros((L, L)), Q0, replica=0)
        eng.set_state(S1, np.zeros((L, L)), Q1, replica=1)
        eng.step(10)
        st0, st1 = eng.status(0), eng.status(1)
        res.append((st0.stopped_at, st0.iteration, st1.stopped_at, st1.iteration,
                    eng.get_state(0), eng.get_state(1), eng.stats(0), eng.stats(1)))
        eng.step(5)                        # a stopped replica stays stopped across chunks
        assert eng.status(0).iteration == 1 and eng.status(1).iteration == 15
        eng.close()
    a, b = res
    assert a[:4] == b[:4] == (1, 1, -1, 10)
    for x, y in zip(a[4] + a[5], b[4] + b[5]):
        assert np.array_equal(x, y)
    assert (a[4][0] == 1).all()
    for x, y in ((a[6], b[6]), (a[7], b[7])):
        assert np.array_equal(x[:, EXACT], y[:, EXACT])
    # uniform initial lattice: nothing runs at all
    eng = _engine(p, precision="fp32")
    eng.set_state(np.zeros((L, L), np.uint8), np.zeros((L, L)), Q0)
    eng.step(5)
    assert eng.status().iteration == 0 and eng.status().stopped_at == 0
    eng.close()


def test_resident_then_replay_then_resident():
    """The planes a resident chunk writes back (ghost cells included) feed the per-iteration
    kernels: a replayed chunk in between must match the oracle fed the same draws."""
    from oracle import c_oracle
    L = 40
    p = full_params(dict(C1, L=L))
    S0, Q0 = _init(L, 12)
    rs = np.random.RandomState(4)
    u = rs.rand(6, L, L)
    b = rs.randint(0, 2, (6, L, L)).astype(np.uint8)
    eng = _engine(p, seeds=31, precision="fp32")
    eng.set_state(S0, np.zeros((L, L)), Q0)
    eng.step(9)                                   # resident, Philox
    eng.set_replay(u, b)
    eng.step(6)                                   # general kernel, replayed draws
    sim = c_oracle.Sim(p, S0, np.zeros((L, L)), Q0, "fp32", seed=31)
    sim.run(9)
    sim.run(6, lambda t, L_: (u[t - 10], b[t - 10]))
    S, R, Q = eng.get_state()
    assert np.array_equal(S, sim.S) and np.array_equal(R, sim.R.astype(np.float64))
    assert np.array_equal(Q.astype(np.float32), sim.Q)
    eng.close()


def test_grid_mode_chunks_and_early_exit(monkeypatch):
    """Grid mode (cooperative launch): chunking is invisible (statistics rows folded every 16
    iterations and at the end of a chunk), and the early exit leaves every block at the same
    iteration."""
    L = 420
    p = full_params(dict(C1, L=L))
    S0, Q0 = _init(L, 41)
    outs = []
    for chunks in ((37,), (1, 15, 16, 5)):
        eng = _engine(p, seeds=12, precision="fp32")
        eng.set_state(S0, np.zeros((L, L)), Q0)
        rows = []
        for c in chunks:
            eng.step(c)
            rows.append(eng.stats()[1:])
        outs.append(eng.get_state() + (np.vstack(rows),))
        assert eng.status().iteration == 37
        eng.close()
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    # everybody prefers to defect, no exploration: all D after iteration 1, iteration 2 breaks
    p0 = full_params(dict(C1, L=L, epsilon=0.0, epsilon_min=0.0))
    Qd = np.zeros((L, L, 2, 2))
    Qd[..., 1] = 1.0
    res = []
    for no_res in (False, True):
        if no_res:
            monkeypatch.setenv("SPGG_NO_RESIDENT", "1")
        else:
            monkeypatch.delenv("SPGG_NO_RESIDENT", raising=False)
        eng = _engine(p0, seeds=1, precision="fp32")
        eng.set_state(S0, np.zeros((L, L)), Qd)
        eng.step(20)
        st = eng.status()
        res.append((st.stopped_at, st.iteration) + eng.get_state() + (eng.stats(),))
        eng.step(3)
        assert eng.status().iteration == 1
        eng.close()
    a, b = res
    assert a[:2] == b[:2] == (1, 1)
    for x, y in zip(a[2:5], b[2:5]):
        assert np.array_equal(x, y)
    assert np.array_equal(a[5][:, EXACT], b[5][:, EXACT])
