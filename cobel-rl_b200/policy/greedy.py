"""Epsilon-greedy policies (reference: policy/greedy.py:10-147)."""
import torch

from .policy import Policy


class EpsilonGreedy(Policy):
    """``p = eps/n_valid + (1-eps)*tie/n_ties`` over the unmasked actions (greedy.py:60-88)."""
    kind = 0
    param_name = 'epsilon'

    def __init__(self, epsilon=0.1, rng=None):
        super().__init__(rng)
        e = torch.as_tensor(epsilon, dtype=torch.float64)
        assert bool(((e >= 0.0) & (e <= 1.0)).all())
        self.epsilon = epsilon



class ExclusiveEpsilonGreedy(EpsilonGreedy):
    """Exploration mass only on the non-greedy actions (greedy.py:117-147)."""
    kind = 1

