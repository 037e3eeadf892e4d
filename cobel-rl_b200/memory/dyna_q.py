"""Dyna-Q memory (reference: memory/dyna_q.py:17-157).

Holds, per agent, the learned model the replay samples from: EMA reward table,
last observed next state and the non-terminal flag of every (state, action).
Inside ``DynaQ.train()`` the tables are read and written by the fused kernel
(csrc/dynaq.cu); ``store`` / ``retrieve`` / ``retrieve_batch`` stay available for
interactive use on the whole batch of agents.
"""
import torch

from .. import _lib
from ..experience import Experience, ExperienceBatch  # noqa: F401  (Experience: memory/dyna_q.py:8-14)
from ..stream import BatchStream, cuda_stream


class TableMemory:
    """Common per-agent tables ``rewards / states / terminals`` with a leading agent axis."""

    def __init__(self, states, actions, learning_rate, rng, init_self_loops):
        assert rng is None or isinstance(rng, BatchStream), 'rng must be a cobel_rl_b200.BatchStream'
        self.rng = rng
        self.number_of_states = int(states)
        self.number_of_actions = int(actions)
        self.learning_rate = learning_rate
        self._init_self_loops = init_self_loops
        self._alloc_for = None
        if rng is not None:
            self._allocate(rng)

    def _allocate(self, stream):
        if self._alloc_for is stream:
            return
        assert self._alloc_for is None, 'memory is already bound to another BatchStream'
        n, S, A, dev = stream.n_agents, self.number_of_states, self.number_of_actions, stream.device
        self._rewards = torch.zeros((n, S, A), dtype=torch.float64, device=dev)
        if self._init_self_loops:   # memory/dyna_q.py:74
            self._states = torch.arange(S, dtype=torch.int32, device=dev).reshape(1, S, 1).repeat(n, 1, A).contiguous()
        else:                       # memory/pma.py:136
            self._states = torch.zeros((n, S, A), dtype=torch.int32, device=dev)
        self._terminals = torch.zeros((n, S, A), dtype=torch.int32, device=dev)
        self._alloc_for = stream
        self.rng = stream
        self._allocate_extra(stream)

    def _allocate_extra(self, stream):
        pass

    def _view(self, t):
        return t[0] if self._alloc_for.single else t

    # public attributes with the reference's names
    rewards = property(lambda self: self._view(self._rewards))
    states = property(lambda self: self._view(self._states))
    terminals = property(lambda self: self._view(self._terminals))

    def _batched(self, x, dtype):
        t = torch.as_tensor(x, dtype=dtype, device=self._alloc_for.device).reshape(-1)
        return t.expand(self._alloc_for.n_agents) if t.numel() == 1 else t

    def _table_params(self, keep, Q=None, lr=None, gamma=None):
        """``CobelDynaQParams`` around the memory tables (and, for the agent-level calls, its Q table)."""
        st = self._alloc_for
        S, A = self.number_of_states, self.number_of_actions
        mlr = st.param(self.learning_rate, 'memory learning_rate')
        keep.append(mlr)
        world = _lib.World(S, A, 0, 0, None, None, None, None, None, None, None)
        return _lib.DynaQParams(st.n_agents, world, st.c_struct(), _lib.Policy(0, 0, None), _lib.Trace(),
                                _lib.ptr(Q), self._rewards.data_ptr(), self._states.data_ptr(),
                                self._terminals.data_ptr(), None, 0, _lib.ptr(lr), _lib.ptr(gamma), mlr.data_ptr(),
                                0, 0, 0, 1, 0, 0)

    def _check_experience(self, batch):
        S, A = self.number_of_states, self.number_of_actions
        bad = ((batch.state < 0) | (batch.state >= S) | (batch.action < 0) | (batch.action >= A) |
               (batch.next_state < 0) | (batch.next_state >= S))
        if bool(bad.any()):
            raise IndexError('experience with a state / action outside the tables')

    def store(self, experience):
        """memory/dyna_q.py:77-96 for all agents (fields are scalars or ``[N]`` tensors); one launch of csrc/ops.cu."""
        st = self._alloc_for
        batch = ExperienceBatch.from_dicts(st, experience)
        self._check_experience(batch)
        keep = []
        p, e = self._table_params(keep), batch.c_struct()
        _lib.call('cobel_dynaq_op', st.device, p, _lib.OP_STORE, e, cuda_stream(st.device))

    def retrieve(self, state, action):
        """memory/dyna_q.py:98-120."""
        st = self._alloc_for
        n = torch.arange(st.n_agents, device=st.device)
        s = self._batched(state, torch.int64)
        a = self._batched(action, torch.int64)
        sq = (lambda t: t[0].item()) if st.single else (lambda t: t)
        return {'state': state, 'action': action, 'reward': sq(self._rewards[n, s, a]),
                'next_state': sq(self._states[n, s, a]), 'terminal': sq(self._terminals[n, s, a])}


class DynaQMemory(TableMemory):
    def __init__(self, states, actions, learning_rate=0.9, rng=None):
        super().__init__(states, actions, learning_rate, rng, init_self_loops=True)

    def retrieve_batch(self, batch_size=32):
        """memory/dyna_q.py:122-157: ``batch_size`` uniform draws over S*A per agent, C-order unravel.  Returns the
        list of ``batch_size`` Experience dicts (fields ``[N]`` tensors)."""
        return self._retrieve(batch_size).to_dicts()

    def _retrieve(self, batch_size):
        st = self._alloc_for
        batch = ExperienceBatch(st, batch_size)
        if batch_size > 0:
            keep = []
            p, e = self._table_params(keep), batch.c_struct()
            _lib.call('cobel_dynaq_op', st.device, p, _lib.OP_RETRIEVE_BATCH, e, cuda_stream(st.device))
        return batch
