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

    def get_action_probs(self, v, mask=None):
        v, m = self._prep(v, mask)
        b = torch.as_tensor(self.beta, dtype=torch.float64).to(v.device)
        if b.dim() > 0 and v.dim() > 1:
            b = b.reshape(-1, 1)
        vmax = torch.where(m, v, torch.full_like(v, -float('inf'))).amax(dim=-1, keepdim=True)
        e = torch.where(m, torch.exp((v - vmax) * b), torch.zeros_like(v))
        return e / e.sum(dim=-1, keepdim=True)
