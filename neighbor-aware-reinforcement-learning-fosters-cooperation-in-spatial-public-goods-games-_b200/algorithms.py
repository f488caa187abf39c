"""Host-side mirror of the reference's RL-algorithm interface
(``src/model/algorithms.py``: ``RLAlgorithm`` :10-93, ``QLearning`` :96, ``SARSA`` :136,
``ExpectedSARSA`` :181, ``DoubleQLearning`` :237, ``create_algorithm`` :344).

In this build the policy and the TD rule are not Python callables: they are compiled
into the fused CUDA step (``csrc/spgg_kernels.cuh``), selected by ``kernel_tag``.  The
classes keep the reference's constructor signature, attributes and ``decay_epsilon``
so code that inspects ``spgg.algorithm`` keeps working; ``select_action`` /
``update_q_table`` cannot run on host arrays (there is no CPU path) and say so.
"""
from __future__ import annotations


class RLAlgorithm:
    """Hyper-parameter holder; same attributes as algorithms.py:34-38."""

    kernel_tag = None       # name of the compiled TD rule, None = not built into the kernel
    name = "abstract"

    def __init__(self, alpha, gamma, epsilon, epsilon_decay, epsilon_min, **kwargs):
        self.alpha = alpha
        self.gamma = gamma
        self.epsilon = epsilon
        self.epsilon_decay = epsilon_decay
        self.epsilon_min = epsilon_min

    def decay_epsilon(self):
        """algorithms.py:40-42 (iterated product, floor at epsilon_min)."""
        self.epsilon = max(self.epsilon * self.epsilon_decay, self.epsilon_min)

    def _fused(self, what):
        raise RuntimeError(
            f"{type(self).__name__}.{what} is fused into the CUDA lattice step "
            "(k_step in csrc/spgg_kernels.cuh); it cannot be applied to host arrays and "
            "there is no CPU fallback. Drive the simulation through SPGG.run().")

    def select_action(self, q_table, states, L, **kwargs):
        self._fused("select_action")

    def update_q_table(self, q_table, old_states, actions, rewards, new_states, **kwargs):
        self._fused("update_q_table")


class QLearning(RLAlgorithm):
    """Q(s,a) += alpha*(r + gamma*max_a' Q(s',a') - Q(s,a))   (algorithms.py:112-133)."""
    kernel_tag = "qlearning"
    name = "qlearning"


class SARSA(RLAlgorithm):
    """TD target Q(s',a') with a' drawn from the epsilon-greedy policy (algorithms.py:152-178;
    three draw pairs per iteration, spgg.py:410,433,452)."""
    kernel_tag = "sarsa"
    name = "sarsa"


class ExpectedSARSA(RLAlgorithm):
    """Expected TD target under the epsilon-greedy policy (algorithms.py:197-234)."""
    kernel_tag = "expected_sarsa"
    name = "expected_sarsa"


class DoubleQLearning(RLAlgorithm):
    """Two tables; a fair draw per site picks the one to update, its target comes from the
    other (algorithms.py:292-341); the policy and the statistics use their mean (:263)."""
    kernel_tag = "double_qlearning"
    name = "double_qlearning"

    def __init__(self, alpha, gamma, epsilon, epsilon_decay, epsilon_min, **kwargs):
        super().__init__(alpha, gamma, epsilon, epsilon_decay, epsilon_min, **kwargs)
        self.q_table_1 = None
        self.q_table_2 = None

    def initialize_q_tables(self, shape, rng=None):
        """algorithms.py:249-260 (two uniform(-0.01, 0.01) draws, table 1 first)."""
        import numpy as np
        rng = np.random if rng is None else rng
        self.q_table_1 = rng.uniform(low=-0.01, high=0.01, size=shape)
        self.q_table_2 = rng.uniform(low=-0.01, high=0.01, size=shape)

    def get_combined_q_table(self):
        """algorithms.py:262-266."""
        if self.q_table_1 is None or self.q_table_2 is None:
            raise ValueError("Q-tables not initialized. Call initialize_q_tables first.")
        return (self.q_table_1 + self.q_table_2) / 2


_FACTORY = {
    "qlearning": QLearning, "q-learning": QLearning,
    "sarsa": SARSA,
    "expected_sarsa": ExpectedSARSA, "expected-sarsa": ExpectedSARSA,
    "double_qlearning": DoubleQLearning, "double-q-learning": DoubleQLearning,
}


def create_algorithm(algorithm_name, alpha, gamma, epsilon, epsilon_decay, epsilon_min, **kwargs):
    """Same names and error as algorithms.py:344-383."""
    key = algorithm_name.lower()
    if key not in _FACTORY:
        raise ValueError(f"Unknown algorithm: {key}. "
                         f"Supported: 'qlearning', 'sarsa', 'expected_sarsa', 'double_qlearning'")
    return _FACTORY[key](alpha, gamma, epsilon, epsilon_decay, epsilon_min, **kwargs)
