"""Policy base class (reference: policy/policy.py:10-67).

Inside ``Agent.train()/test()`` the policy is evaluated by the fused CUDA kernels
(csrc/common.cuh: action_probs / draw_categorical); the object only carries the
kind and the per-agent parameter.  ``get_action_probs`` / ``select_action`` are
kept for interactive use and act on ``[N, A]`` batches (or one ``[A]`` row for a
single-agent stream).
"""
import abc

import torch

from .. import _lib
from ..stream import BatchStream


class Policy(abc.ABC):
    kind = -1           # COBEL_POLICY_* (include/cobel_b200.h)
    param_name = ''

    def __init__(self, rng=None):
        assert rng is None or isinstance(rng, BatchStream), 'rng must be a cobel_rl_b200.BatchStream'
        self.rng = rng

    def _param(self):
        return getattr(self, self.param_name)

    def c_struct(self, stream, keep):
        """``CobelPolicy`` for N agents; ``keep`` collects tensors that must outlive the launch."""
        par = stream.param(self._param(), self.param_name)
        keep.append(par)
        return _lib.Policy(self.kind, 0, par.data_ptr())

    @abc.abstractmethod
    def get_action_probs(self, v, mask=None):
        ...

    def select_action(self, v, mask=None):
        """Probabilities, then one inverse-CDF draw per agent (policy/greedy.py:58):
        ``searchsorted(cumsum(p) / cumsum(p)[-1], u, side='right')``."""
        assert self.rng is not None, 'select_action outside an agent needs rng=BatchStream(...)'
        p = self.get_action_probs(v, mask)
        single = p.dim() == 1
        p2 = p.reshape(-1, p.shape[-1])
        cdf = torch.cumsum(p2, dim=1)
        cdf = cdf / cdf[:, -1:]
        u = self.rng.next(1).to(p2.device)
        a = (cdf <= u).sum(dim=1)
        return int(a[0]) if single else a

    @staticmethod
    def _prep(v, mask):
        v = torch.as_tensor(v, dtype=torch.float64)
        m = torch.ones_like(v, dtype=torch.bool) if mask is None else torch.as_tensor(mask).to(v.device).bool()
        assert bool((m.sum(dim=-1) > 0).all()), 'The action mask masks all actions!'
        return v, m
