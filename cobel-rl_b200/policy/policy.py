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

    def _rows(self, v, mask):
        """(values [R, A] on the stream's device, mask bytes or None, leading shape, stream)."""
        st = self.rng
        assert st is not None, 'policies act through a cobel_rl_b200.BatchStream (rng=...)'
        v = torch.as_tensor(v, dtype=torch.float64, device=st.device)
        shape = v.shape
        v2 = v.reshape(-1, shape[-1]).contiguous()
        m = None
        if mask is not None:
            m = torch.as_tensor(mask, device=st.device).bool().expand(shape).reshape(-1, shape[-1])
            assert bool((m.sum(dim=-1) > 0).all()), 'The action mask masks all actions!'
            m = m.to(torch.uint8).contiguous()
        return v2, m, shape, st

    def get_action_probs(self, v, mask=None):
        """``get_action_probs`` of the reference for ``[A]`` (single agent), ``[N, A]`` or ``[N, R, A]`` values
        (policy/greedy.py:60-88,117-147, policy/softmax.py:60-88), evaluated by csrc/ops.cu."""
        from ..stream import cuda_stream
        v2, m, shape, st = self._rows(v, mask)
        rows = v2.shape[0]
        assert rows % st.n_agents == 0, 'values must have one row (or R rows) per agent'
        keep = []
        pol = self.c_struct(st, keep)
        out = torch.empty_like(v2)
        _lib.call('cobel_policy_probs', st.device, pol, st.n_agents, rows // st.n_agents, shape[-1], v2.data_ptr(),
                  _lib.ptr(m), out.data_ptr(), cuda_stream(st.device))
        return out.reshape(shape)

    def select_action(self, v, mask=None):
        """Probabilities, then one inverse-CDF draw per agent (policy/greedy.py:40-58):
        ``searchsorted(cumsum(p) / cumsum(p)[-1], u, side='right')``; one launch of csrc/ops.cu."""
        from ..stream import cuda_stream
        v2, m, shape, st = self._rows(v, mask)
        assert v2.shape[0] == st.n_agents, 'select_action takes one row of values per agent'
        keep = []
        pol, s = self.c_struct(st, keep), st.c_struct()
        a = torch.empty(st.n_agents, dtype=torch.int32, device=st.device)
        _lib.call('cobel_policy_select', st.device, pol, s, st.n_agents, shape[-1], v2.data_ptr(), _lib.ptr(m),
                  a.data_ptr(), cuda_stream(st.device))
        return int(a[0]) if st.single else a.long()
