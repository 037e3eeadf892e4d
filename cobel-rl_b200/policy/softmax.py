"""Softmax policy (reference: policy/softmax.py:10-88)."""
import torch

from .policy import Policy


class Softmax(Policy):
    """``p = exp(beta*(v - max v)) / sum`` over the unmasked actions (softmax.py:60-88)."""
    kind = 2
    param_name = 'beta'

    def __init__(self, beta=1.0, rng=None):
        super().__init__(rng)
        assert bool((torch.as_tensor(beta, dtype=torch.float64) > 0.0).all()), 'Inverse temperature must be non-negative!'
        self.beta = beta

