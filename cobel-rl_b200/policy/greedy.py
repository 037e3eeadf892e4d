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

    def _eps(self, v):
        e = torch.as_tensor(self.epsilon, dtype=torch.float64).to(v.device)
        return e.reshape(-1, 1) if (e.dim() > 0 and v.dim() > 1) else e

    def get_action_probs(self, v, mask=None):
        v, m = self._prep(v, mask)
        eps = self._eps(v)
        vmax = torch.where(m, v, torch.full_like(v, -float('inf'))).amax(dim=-1, keepdim=True)
        ties = (v == vmax) & m
        nv = m.sum(dim=-1, keepdim=True).to(torch.float64)
        nt = ties.sum(dim=-1, keepdim=True).to(torch.float64)
        p = eps / nv + ((1.0 - eps) * ties.to(torch.float64)) / nt
        return torch.where(m, p, torch.zeros_like(p))


class ExclusiveEpsilonGreedy(EpsilonGreedy):
    """Exploration mass only on the non-greedy actions (greedy.py:117-147)."""
    kind = 1

    def get_action_probs(self, v, mask=None):
        v, m = self._prep(v, mask)
        eps = self._eps(v)
        vmax = torch.where(m, v, torch.full_like(v, -float('inf'))).amax(dim=-1, keepdim=True)
        ties = (v == vmax) & m
        nv = m.sum(dim=-1, keepdim=True)
        nt = ties.sum(dim=-1, keepdim=True)
        d = torch.clamp(nv - nt, min=1).to(torch.float64)
        p = ((1.0 - eps) * ties.to(torch.float64)) / nt.to(torch.float64) \
            + (eps * (~ties & m).to(torch.float64)) / d
        return torch.where(m, p, torch.zeros_like(p))
