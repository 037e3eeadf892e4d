"""Dyna-Q memory (reference: memory/dyna_q.py:17-157).

Holds, per agent, the learned model the replay samples from: EMA reward table,
last observed next state and the non-terminal flag of every (state, action).
Inside ``DynaQ.train()`` the tables are read and written by the fused kernel
(csrc/dynaq.cu); ``store`` / ``retrieve`` / ``retrieve_batch`` stay available for
interactive use on the whole batch of agents.
"""
import torch

from ..stream import BatchStream


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

    def store(self, experience):
        """memory/dyna_q.py:77-96 for all agents (fields are scalars or ``[N]`` tensors)."""
        st = self._alloc_for
        n = torch.arange(st.n_agents, device=st.device)
        s = self._batched(experience['state'], torch.int64)
        a = self._batched(experience['action'], torch.int64)
        r = self._batched(experience['reward'], torch.float64)
        lr = st.param(self.learning_rate, 'learning_rate')
        cur = self._rewards[n, s, a]
        self._rewards[n, s, a] = cur + lr * (r - cur)
        self._states[n, s, a] = self._batched(experience['next_state'], torch.int32)
        self._terminals[n, s, a] = self._batched(experience['terminal'], torch.int32)

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
        """memory/dyna_q.py:122-157: ``batch_size`` uniform draws over S*A per agent,
        C-order unravel.  Returns a dict of ``[N, batch]`` tensors."""
        st = self._alloc_for
        S, A = self.number_of_states, self.number_of_actions
        u = st.next(batch_size)
        idx = torch.clamp((u * (S * A)).floor().to(torch.int64), max=S * A - 1)
        s, a = idx // A, idx % A
        n = torch.arange(st.n_agents, device=st.device).reshape(-1, 1)
        return {'state': s, 'action': a, 'reward': self._rewards[n, s, a],
                'next_state': self._states[n, s, a], 'terminal': self._terminals[n, s, a]}
