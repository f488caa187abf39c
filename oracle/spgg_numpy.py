"""TEST INFRASTRUCTURE ONLY - CPU restatement (NumPy, fp64) of the reference's
SPGG step.  Never imported by the product path; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs use it.

Parity status: PINNED AGAINST THE EXECUTED REFERENCE.  The reference ships no
tests or golden vectors, so the pin is the reference
itself, run in the build container by ``oracle/ref_harness.py``:
``tests/test_oracle_vs_reference.py`` (container only) and the committed
fixtures ``tests/golden/*.npz`` (made by ``oracle/make_golden.py``) compare
this file bit-for-bit (S, R, Q, integer series) against it.

Every function cites the reference lines it restates
(paths relative to ``/root/reference``).  Conventions: axis 0 = row ``i``,
axis 1 = column ``j``, periodic; strategy/action 0 = cooperate, 1 = defect.
``at(X, di, dj)[i, j] == X[i - di, j - dj]`` (the value ``np.roll`` by
``(di, dj)`` brings to ``(i, j)``).
"""
from __future__ import annotations

import numpy as np

# neighbour offsets in the order the reference enumerates them
OFFSETS_M1 = ((1, 0), (-1, 0), (0, 1), (0, -1))                 # spgg.py:485
OFFSETS_M2 = OFFSETS_M1 + ((2, 0), (-2, 0), (0, 2), (0, -2),
                           (1, 1), (1, -1), (-1, 1), (-1, -1))  # spgg.py:479-483

SNAPSHOT_ITERS = (1, 10, 100, 1000, 5000, 10000, 20000, 30000, 40000)  # spgg.py:153


def at(X, di, dj):
    return np.roll(X, shift=(di, dj), axis=(0, 1))


def offsets_for(M):
    return OFFSETS_M2 if M == 2 else OFFSETS_M1


def group_counts(C):
    """Cooperators in the 5-site group centred on each site.
    spgg.py:23-36 (operand order: self, i+1, i-1, j+1, j-1; integers, so the
    order is irrelevant)."""
    return C + at(C, -1, 0) + at(C, 1, 0) + at(C, 0, -1) + at(C, 0, 1)


def normalised_payoff(S, r, c, cost):
    """spgg.py:230-259 summed over the five groups a site belongs to in the
    order of spgg.py:373-377 (own group, (i-1,j), (i+1,j), (i,j-1), (i,j+1)),
    then normalised with spgg.py:148-149."""
    C = (S == 0).astype(np.int64)
    D = 1 - C
    N = group_counts(C)
    total = None
    for (di, dj) in ((0, 0), (1, 0), (-1, 0), (0, 1), (0, -1)):
        Ng = at(N, di, dj)                       # group centred at (i-di, j-dj)
        share = r * c * Ng / 5
        term = (share - cost) * C + share * D
        total = term if total is None else total + term
    lo = r - 5
    hi = 4 * r
    return (total - lo) / (hi - lo)


def state_of(R, S, M, state_representation):
    """spgg.py:281-310."""
    if state_representation == "action":
        return (S == 0).astype(np.int64)
    if state_representation != "reputation":
        raise ValueError(f"Unknown state_representation: {state_representation}")
    acc = np.zeros_like(R, dtype=np.float64)
    for (di, dj) in ((0, 0),) + offsets_for(M):
        acc += at(R, di, dj)
    n = 1 + len(offsets_for(M))
    return (acc / n > 0).astype(np.int64)


def epsilon_sequence(eps0, decay, eps_min, n):
    """eps used at iteration t (t = 1..n) and the value recorded after it.
    algorithms.py:40-42 (iterated product, not pow)."""
    used = np.empty(n)
    after = np.empty(n)
    e = float(eps0)
    for t in range(n):
        used[t] = e
        e = max(e * decay, eps_min)
        after[t] = e
    return used, after


def _td_target(Qtab, ii, jj, s_new, algo, eps, u, b):
    """Value of the next state used by the TD error, per rule:
    qlearning  max_a Q[s',a]                                   algorithms.py:125
    sarsa      Q[s',a'] with a' drawn eps-greedily from Q[s']   algorithms.py:142-150,169 (draws u, b)
    expected_sarsa  sum_a pi(a|s') Q[s',a], pi eps-greedy       algorithms.py:212-224"""
    row = Qtab[ii, jj, s_new, :]
    if algo == "qlearning":
        return np.max(row, axis=2)
    greedy = np.argmax(row, axis=2)
    if algo == "sarsa":
        a_next = np.where(u < eps, b.astype(np.int64), greedy)
        return Qtab[ii, jj, s_new, a_next]
    if algo == "expected_sarsa":
        probs = np.full(row.shape, eps / 2)
        probs[ii, jj, greedy] = (1 - eps) + eps / 2
        return np.sum(probs * row, axis=2)
    raise ValueError(f"oracle: algorithm {algo!r} not restated")


def qlearning_step(S, R, Q, eps, u, b, p, algo="qlearning", extra=None):
    """One full iteration of spgg.py:368-592 for the Q-learning rule
    (algorithms.py:96-133) - or, with ``algo`` = 'sarsa' / 'expected_sarsa', the rules of
    algorithms.py:136-234 as spgg.py:431-473 applies them.  ``u``/``b`` are the iteration's
    action draws; ``extra`` = ((u2, b2), (u3, b3)) are SARSA's two further draw pairs
    (next action for the update, spgg.py:433, and for the NI statistic, spgg.py:452).
    Returns (S', R', Q', stats dict).  Q is updated on a copy."""
    L = S.shape[0]
    M = 2 if p["use_second_order"] else 1
    offs = offsets_for(M)
    r, c, cost = p["r"], p["c"], p["cost"]
    wP = p["reward_weight_payoff"]
    wR = 1 - wP                                                   # spgg.py:108
    alpha, gamma = p["alpha"], p["gamma"]
    kappa, lam_eps = p["influence_factor"], p["lambda_epsilon"]
    ii, jj = np.indices((L, L))
    st = {}

    P = normalised_payoff(S, r, c, cost)                          # spgg.py:373-377
    C_old = (S == 0)
    nC = int(C_old.sum())
    st["coop_rate"] = nC / (L * L)
    st["def_rate"] = int((S == 1).sum()) / (L * L)
    st["P_sum"] = P.sum()
    st["P_mean"] = np.mean(P)
    st["P_mean_C"] = np.mean(P[C_old]) if nC else 0
    st["P_mean_D"] = np.mean(P[~C_old]) if nC < L * L else 0
    st["rep_avg"] = np.mean(R)
    st["P"] = P

    dq = (algo == "double_qlearning")
    if dq:
        # two tables; the policy and every statistic use their mean (algorithms.py:263, 283-290)
        QA, QB = Q[0].copy(), Q[1].copy()
        Q = (QA + QB) / 2
    s_old = state_of(R, S, M, p["state_representation"])          # spgg.py:409
    greedy = np.argmax(Q[ii, jj, s_old, :], axis=2)               # algorithms.py:106-107
    a = np.where(u < eps, b.astype(np.int64), greedy)             # algorithms.py:105,109
    gain = np.where(a == 0, p["rep_gain_C"], -p["delta_R_D"])     # spgg.py:321
    R2 = np.clip(R + gain, p["R_min"], p["R_max"])                # spgg.py:322-323
    S2 = a.copy()
    st["switch_C_to_D"] = int(((S == 0) & (S2 == 1)).sum())       # spgg.py:419
    st["switch_D_to_C"] = int(((S == 1) & (S2 == 0)).sum())       # spgg.py:420

    s_new = state_of(R2, S2, M, p["state_representation"])        # spgg.py:423
    rep_reward = np.where(a == 0, 0.5, 0)                         # spgg.py:424
    st["payoff_component"] = np.mean(wP * P)
    st["rep_component"] = np.mean(wR * rep_reward)
    rew = wP * P + wR * rep_reward                                # spgg.py:427

    if dq:
        # algorithms.py:292-341: rand(L,L) < 0.5 picks the table to update; its target is the
        # other table's value at the other table's greedy action, i.e. that table's row maximum
        upd1 = extra[0] < 0.5
        qa, qb = QA[ii, jj, s_old, a], QB[ii, jj, s_old, a]
        nxt_a = np.max(QB[ii, jj, s_new, :], axis=2)
        nxt_b = np.max(QA[ii, jj, s_new, :], axis=2)
        td_a = rew + gamma * nxt_a - qa
        td_b = rew + gamma * nxt_b - qb
        QA[ii[upd1], jj[upd1], s_old[upd1], a[upd1]] += alpha * td_a[upd1]
        QB[ii[~upd1], jj[~upd1], s_old[~upd1], a[~upd1]] += alpha * td_b[~upd1]
        Q2 = (QA + QB) / 2
        q_cur = Q2[ii, jj, s_old, a]                              # spgg.py:464-468
        td2 = rew + gamma * np.max(Q2[ii, jj, s_new, :], axis=2) - q_cur
    else:
        Q2 = Q.copy()
        q0 = Q[ii, jj, s_old, a]
        x1 = extra[0] if extra else (None, None)
        x2 = extra[1] if extra else (None, None)
        nxt = _td_target(Q, ii, jj, s_new, algo, eps, *x1)
        td = rew + gamma * nxt - q0                               # algorithms.py:128
        Q2[ii, jj, s_old, a] = q0 + alpha * td                    # algorithms.py:131

        # TD error on the updated table, only feeds the NI statistic (spgg.py:446-473)
        q_cur = Q2[ii, jj, s_old, a]
        td2 = rew + gamma * _td_target(Q2, ii, jj, s_new, algo, eps, *x2) - q_cur

    diffs = np.stack([at(rew, di, dj) - rew for (di, dj) in offs])  # spgg.py:486
    best = diffs.max(axis=0)
    gmax = np.max(np.abs(diffs))                                  # spgg.py:488
    lam = kappa * np.maximum(0, best) / (gmax + lam_eps)          # spgg.py:489
    kstar = np.argmax(diffs, axis=0)                              # first max
    nbr_a = np.stack([at(a, di, dj) for (di, dj) in offs])
    a_star = nbr_a[kstar, ii, jj]
    nu = lam * np.where(a_star == a, 1.0, -1.0)                   # spgg.py:494-495
    if dq:                                                        # spgg.py:498-505
        QA[ii, jj, s_old, a] += nu
        QB[ii, jj, s_old, a] += nu
        Q2 = (QA + QB) / 2
    else:
        Q2[ii, jj, s_old, a] += nu                                # spgg.py:509
    st["gmax"] = gmax

    pct = np.abs(nu) / (np.abs(alpha * td2) + np.abs(nu) + 1e-8) * 100
    st["neighbor_influence_percent"] = np.mean(pct)               # spgg.py:512-513
    pos = best > 0
    st["best_neighbor_second_order_percent"] = (
        np.mean((kstar >= 4)[pos]) * 100 if pos.any() else 0)     # spgg.py:516-526
    cm = (a == 0)
    if cm.any():                                                  # spgg.py:529-538
        st["reputation_reward_ratio"] = np.mean(
            np.abs(wR * rep_reward[cm]) / (np.abs(rew[cm]) + 1e-9) * 100)
    else:
        st["reputation_reward_ratio"] = np.nan
    st["avg_reward_C"] = np.mean(rew[cm]) if cm.any() else 0      # spgg.py:542
    st["avg_reward_D"] = np.mean(rew[~cm]) if (~cm).any() else 0  # spgg.py:543

    names = ("q_s0_c", "q_s0_d", "q_s1_c", "q_s1_d")
    was_C, was_D = (S == 0), (S == 1)
    for k, nm in enumerate(names):                                # spgg.py:562-583
        plane = Q2[:, :, k // 2, k % 2]
        st["avg_" + nm] = np.mean(plane)
        st["cooperators_" + nm] = np.mean(plane[was_C]) if was_C.any() else np.nan
        st["defectors_" + nm] = np.mean(plane[was_D]) if was_D.any() else np.nan
    D2 = (S2 == 1).astype(np.int64)
    nd = group_counts(D2)                                         # spgg.py:586-592
    for k in range(6):
        st[f"group_comp_d{k}"] = (int((nd == k).sum()) / (L * L)) * 100
    st["rew"] = rew
    if dq:
        return S2, R2, (QA, QB), st
    return S2, R2, Q2, st


def simulate(p, S0, R0, Q0, draws):
    """Run ``p['iterations']`` iterations like spgg.py:368-592 and assemble the
    series under the reference's HDF5 dataset names (spgg.py:595-629).
    ``draws(t, L)`` returns ``(u, b)`` for iteration ``t`` (1-based)."""
    L = S0.shape[0]
    S, R = S0.astype(np.int64).copy(), R0.astype(np.float64).copy()
    Q = (Q0[0].copy(), Q0[1].copy()) if isinstance(Q0, tuple) else Q0.copy()   # Double-Q: (table 1, table 2)
    eps = p["epsilon"]
    series = {}

    def push(k, v):
        series.setdefault(k, []).append(v)

    snaps = {}
    P_last = None
    for t in range(1, p["iterations"] + 1):
        if (S == 0).all() or (S == 1).all():
            # spgg.py:381-406: pre-action records are taken, then the loop breaks
            P_last = normalised_payoff(S, p["r"], p["c"], p["cost"])
            nC = int((S == 0).sum())
            push("coop_rate_history", nC / (L * L))
            push("it_records_final", (nC / (L * L), int((S == 1).sum()) / (L * L), P_last.sum(),
                                      np.mean(P_last),
                                      np.mean(P_last) if nC else 0,
                                      np.mean(P_last) if nC == 0 else 0))
            push("rep_avg_history_final", np.mean(R))
            if t in SNAPSHOT_ITERS:
                snaps[t] = (R.copy(), S.copy())
            break
        if t in SNAPSHOT_ITERS:
            snaps[t] = (R.copy(), S.copy())
        algo = str(p.get("algorithm", "qlearning")).lower()
        if algo == "sarsa":      # three draw pairs per iteration (spgg.py:410, 433, 452)
            u, b, u2, b2, u3, b3 = draws(t, L)
            S, R, Q, st = qlearning_step(S, R, Q, eps, u, b, p, algo, ((u2, b2), (u3, b3)))
        elif algo == "double_qlearning":   # rand, randint, then rand for the table choice
            u, b, u2 = draws(t, L)
            S, R, Q, st = qlearning_step(S, R, Q, eps, u, b, p, algo, (u2,))
        else:
            u, b = draws(t, L)
            S, R, Q, st = qlearning_step(S, R, Q, eps, u, b, p, algo)
        P_last = st["P"]
        eps = max(eps * p["epsilon_decay"], p["epsilon_min"])
        push("coop_rate_history", st["coop_rate"])
        push("it_records_final", (st["coop_rate"], st["def_rate"], st["P_sum"],
                                  st["P_mean"], st["P_mean_C"], st["P_mean_D"]))
        push("rep_avg_history_final", st["rep_avg"])
        push("epsilon_history_final", eps)
        for k in ("switch_C_to_D", "switch_D_to_C", "neighbor_influence_percent",
                  "best_neighbor_second_order_percent", "reputation_reward_ratio"):
            push(k, st[k])
        push("payoff_component_history", st["payoff_component"])
        push("rep_component_history", st["rep_component"])
        push("avg_reward_C_history", st["avg_reward_C"])
        push("avg_reward_D_history", st["avg_reward_D"])
        push("gmax", st["gmax"])
        for k in range(6):
            push(f"group_comp_d{k}_history", st[f"group_comp_d{k}"])
        for nm in ("q_s0_c", "q_s0_d", "q_s1_c", "q_s1_d"):
            push(f"avg_{nm}_history", st["avg_" + nm])
            push(f"cooperators_{nm}_history", st["cooperators_" + nm])
            push(f"defectors_{nm}_history", st["defectors_" + nm])
    out = {k: np.array(v) for k, v in series.items()}
    if isinstance(Q, tuple):
        out["q1_final"], out["q2_final"] = Q
        Q = (Q[0] + Q[1]) / 2
    out["Sn_final"], out["R_final"], out["q_final"] = S, R, Q
    out["snapshots"] = snaps
    out["P_last"] = P_last
    return out


def legacy_draws(seed, L, algorithm="qlearning"):
    """Reconstruct the reference's global-MT19937 stream for a pinned ctor seed
    (SURVEY.md section 8c): ctor ``uniform(L,L,2,2)`` then ``randint(L,L)``
    (spgg.py:121,162); each step ``rand(L,L)`` then ``randint(0,2,(L,L))``
    (algorithms.py:105,108).  Returns (Q0, S0, draws callable)."""
    rs = np.random.RandomState(seed)
    Q0 = rs.uniform(low=-0.01, high=0.01, size=(L, L, 2, 2))
    S0 = rs.randint(0, 2, size=(L, L))
    state = {"t": 0}

    def draws(t, L_):
        assert t == state["t"] + 1, "draws must be consumed in order"
        state["t"] = t
        u = rs.rand(L_, L_)
        b = rs.randint(0, 2, size=(L_, L_)).astype(np.uint8)
        return u, b

    return Q0, S0, draws
