"""Abstract environment (reference: interface/interface.py:37-86).

Batched convention: every method acts on all N agents of the attached
``BatchStream`` at once; with a single-agent stream the return values have the
reference's scalar types.
"""
import abc

import numpy as np
import torch

from .. import _lib
from ..stream import BatchStream


class Interface(abc.ABC):
    def __init__(self, widget=None, rng=None):
        assert widget is None, 'visualisation widgets are out of scope of the B200 path'
        self.widget = None
        self.rng = BatchStream() if rng is None else rng
        assert isinstance(self.rng, BatchStream), 'rng must be a cobel_rl_b200.BatchStream'

    @abc.abstractmethod
    def step(self, action):
        ...

    @abc.abstractmethod
    def reset(self):
        ...

    @abc.abstractmethod
    def get_position(self):
        ...

    # -- table view consumed by the fused kernels (include/cobel_b200.h: CobelWorld)
    def _set_tables(self, succ, reward, terminal, starts, sas=None):
        dev = self.rng.device
        self._tp = None
        if sas is not None:
            # non-deterministic world: CSR of the non-zero entries of sas[s,a,:], ascending next state
            sas = np.asarray(sas, dtype=np.float64)
            S, A, _ = sas.shape
            flat = sas.reshape(S * A, S)
            nz_row, nz_col = np.nonzero(flat)
            off = np.zeros(S * A + 1, dtype=np.int32)
            np.cumsum(np.bincount(nz_row, minlength=S * A), out=off[1:])
            self._tp = (torch.as_tensor(off).to(dev), torch.as_tensor(nz_col.astype(np.int32)).to(dev),
                        torch.as_tensor(flat[nz_row, nz_col]).contiguous().to(dev))
        # largest |s' - s| over all possible transitions (PMA's banded update_sr, include/cobel_b200.h sr_band)
        succ_np = np.asarray(succ, dtype=np.int64)
        self.transition_band = int(np.abs(succ_np - np.arange(succ_np.shape[0])[:, None]).max()) if succ_np.size else 0
        if sas is not None:
            self.transition_band = max(self.transition_band, int(np.abs(nz_col - nz_row // A).max()))
        self._succ = torch.as_tensor(succ, dtype=torch.int32).contiguous().to(dev)
        self._reward = torch.as_tensor(reward, dtype=torch.float64).contiguous().to(dev)
        self._terminal = torch.as_tensor(terminal).to(torch.uint8).contiguous().to(dev)
        self._starts = torch.as_tensor(starts, dtype=torch.int32).contiguous().to(dev)
        assert self._succ.dim() == 2 and self._starts.numel() > 0

    @property
    def n_states(self):
        return self._succ.shape[0]

    @property
    def n_actions(self):
        return self._succ.shape[1]

    def c_world(self):
        tp = self._tp if self._tp is not None else (None, None, None)
        return _lib.World(self.n_states, self.n_actions, self._starts.numel(), 0, self._succ.data_ptr(),
                          self._reward.data_ptr(), self._terminal.data_ptr(), self._starts.data_ptr(),
                          _lib.ptr(tp[0]), _lib.ptr(tp[1]), _lib.ptr(tp[2]))

    # -- Interface.reset() / step() for all agents: one launch of csrc/ops.cu each
    def _launch_reset(self):
        from ..stream import cuda_stream
        st = self.rng
        cur = torch.empty(st.n_agents, dtype=torch.int32, device=st.device)
        w, s = self.c_world(), st.c_struct()
        _lib.call('cobel_env_reset', st.device, w, s, st.n_agents, cur.data_ptr(), cuda_stream(st.device))
        self._current = cur
        return cur

    def _launch_step(self, action):
        from ..stream import cuda_stream
        st = self.rng
        a = torch.as_tensor(action, device=st.device).reshape(-1).to(torch.int32)
        if a.numel() == 1:
            a = a.expand(st.n_agents)
        a = a.contiguous()
        assert a.numel() == st.n_agents, 'one action per agent'
        if bool(((a < 0) | (a >= self.n_actions)).any()):
            raise IndexError('action out of range')          # the reference indexes sas[s, a] / neighbors[a]
        cur = self._current.clone()
        reward = torch.empty(st.n_agents, dtype=torch.float64, device=st.device)
        end = torch.empty(st.n_agents, dtype=torch.uint8, device=st.device)
        w, s = self.c_world(), st.c_struct()
        _lib.call('cobel_env_step', st.device, w, s, st.n_agents, cur.data_ptr(), a.data_ptr(), reward.data_ptr(),
                  end.data_ptr(), cuda_stream(st.device))
        self._current = cur
        return cur, reward, end.bool()

    def _out(self, t):
        """Squeeze the agent axis for single-agent streams (reference return types)."""
        if not self.rng.single:
            return t
        v = t[0]
        return v.item() if v.dim() == 0 else v
