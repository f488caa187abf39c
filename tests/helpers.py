"""Shared helpers for the parity tests (test infrastructure)."""
import json
import os

import numpy as np

# runner.py:88-101 fixes these for every CLI experiment of the reference
RUNNER_FIXED = dict(c=1, cost=1, gamma=0.9, epsilon=0.5, epsilon_decay=0.99, epsilon_min=0.01,
                    lambda_epsilon=0.01, delta_R_C=1, delta_R_D=1, R_min=-10, R_max=10, alpha=0.8)

C1 = dict(RUNNER_FIXED, r=3.0, influence_factor=1.0, use_second_order=False,
          reward_weight_payoff=0.95, rep_gain_C=1.0, state_representation="reputation")
C2 = dict(RUNNER_FIXED, r=4.0, influence_factor=1.0, use_second_order=True,
          reward_weight_payoff=1.0, rep_gain_C=1.0, state_representation="action")


def load_golden(golden_dir, name):
    z = np.load(os.path.join(golden_dir, f"replay_{name}.npz"), allow_pickle=False)
    params = json.loads(str(z["params_json"]))
    return z, params


def full_params(p):
    """Fill the reference ctor defaults (spgg.py:50-56) so oracle and engine see the
    same numbers."""
    d = dict(r=2, c=1, cost=0.5, L=50, iterations=1000, alpha=0.1, gamma=0.9, epsilon=0.5,
             epsilon_decay=0.995, epsilon_min=0.01, influence_factor=1.0, use_second_order=True,
             lambda_epsilon=0.01, delta_R_C=1, delta_R_D=1, R_min=-10, R_max=10,
             reward_weight_payoff=1.0, rep_gain_C=0.5, state_representation="reputation")
    d.update(p)
    return d


def legacy_stream(seed, L, n):
    """ctor + per-step draws of the reference for a pinned seed."""
    rs = np.random.RandomState(seed)
    Q0 = rs.uniform(-0.01, 0.01, (L, L, 2, 2))
    S0 = rs.randint(0, 2, (L, L))
    u = np.empty((n, L, L))
    b = np.empty((n, L, L), np.uint8)
    for t in range(n):
        u[t] = rs.rand(L, L)
        b[t] = rs.randint(0, 2, (L, L))
    return Q0, S0, u, b


def launch_ranks(argv, world, timeout=600, extra_env=None):
    """Start ``world`` copies of tests/strip_worker.py (one rank each, rendezvous on
    127.0.0.1) and return their (returncode, output) pairs."""
    import socket
    import subprocess
    import sys
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "strip_worker.py")
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        env.update(extra_env or {})
        procs.append(subprocess.Popen([sys.executable, worker] + [str(a) for a in argv], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    out = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            p.kill()
            o, _ = p.communicate()
            o += "\n[timeout]"
        out.append((p.returncode, o))
    return out
