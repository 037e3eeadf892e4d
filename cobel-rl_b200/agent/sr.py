"""Successor-representation agent (reference: agent/sr.py:25-324).

``train()`` / ``test()`` run the reference's loop (Q from the SR and the learned reward
vector, action selection, environment step, reward / transition-model / SR row update) for all
N agents in one launch of ``cobel_sr_run`` (csrc/sr.cu).  The reference's dense one-hot
transition model ``transitions[S,A,S]`` is held as its arg-max ``model[S,A]``.
"""
import numpy as np
import torch

from .. import _lib
from ..spaces import Discrete
from .agent import Agent, launch_stream


class SR(Agent):
    def __init__(self, observation_space, action_space, policy, policy_test=None, learning_rate=0.1,
                 gamma=0.99, custom_callbacks=None):
        assert type(observation_space) is Discrete, 'SR requires a discrete observation space!'
        assert type(action_space) is Discrete, 'SR requires a discrete action space!'
        super().__init__(observation_space, action_space, custom_callbacks)
        self.policy = policy
        self.policy_test = policy if policy_test is None else policy_test
        self.learning_rate = learning_rate
        self.gamma = gamma
        self.mask_actions = False
        stream = self._find_stream(self.policy, self.policy_test)
        if stream is not None:
            self._bind(stream)

    def _allocate(self, stream):
        S, A, n, dev = int(self.observation_space.n), int(self.action_space.n), stream.n_agents, stream.device
        self._SR = torch.eye(S, dtype=torch.float64, device=dev).repeat(n, 1, 1).contiguous()     # sr.py:130
        self._model = torch.arange(S, dtype=torch.int32, device=dev).reshape(1, S, 1).repeat(n, 1, A).contiguous()
        self._rewards = torch.zeros((n, S), dtype=torch.float64, device=dev)                       # sr.py:136
        self._action_mask = torch.ones((S, A), dtype=torch.bool, device=dev)

    SR = property(lambda self: self._view(self._SR))
    rewards = property(lambda self: self._view(self._rewards))
    model = property(lambda self: self._view(self._model))

    @property
    def transitions(self):
        """The reference's one-hot ``(S, A, S)`` model (sr.py:131-135), materialised on demand."""
        S = self._SR.shape[1]
        return self._view(torch.nn.functional.one_hot(self._model.long(), S).to(torch.float64))

    @property
    def action_mask(self):
        return self._action_mask

    @action_mask.setter
    def action_mask(self, value):
        self._action_mask = torch.as_tensor(value, device=self._stream.device).bool().contiguous()

    def _run(self, interface, trials, steps, learn):
        if self._stream is None:
            self._bind(interface.rng)
        st = self._stream
        assert interface.rng is st, 'environment and agent must share one BatchStream'
        assert interface.n_states == self._SR.shape[1] and interface.n_actions == self._model.shape[2]
        pol = self.policy if learn else self.policy_test
        results = []
        for _, n_tr in self._chunks(trials):
            keep = []
            tr, res = self._make_trace(n_tr, steps, 0, 0, 0, keep)
            lr, gm = st.param(self.learning_rate, 'learning_rate'), st.param(self.gamma, 'gamma')
            mptr, mstride = self._mask_args(keep)
            p = _lib.SRParams(st.n_agents, interface.c_world(), st.c_struct(), pol.c_struct(st, keep), tr,
                              self._SR.data_ptr(), self._rewards.data_ptr(), self._model.data_ptr(), mptr, mstride,
                              lr.data_ptr(), gm.data_ptr(), n_tr, steps, 1 if learn else 0, 0)
            _lib.check(_lib.lib().cobel_sr_run(p, launch_stream(st)))
            self._check_flags(res)
            self._fire_trial_callbacks(res, self.current_trial)
            self.current_trial += n_tr
            results.append(res)
            if self.stop:
                break
        self.last_run = self._merge(results)
        return self.last_run

    def train(self, interface, trials, steps):
        """agent/sr.py:142-197 for all agents."""
        return self._run(interface, trials, steps, learn=True)

    def test(self, interface, trials, steps):
        """agent/sr.py:199-253 for all agents."""
        return self._run(interface, trials, steps, learn=False)

    def retrieve_q(self, state):
        """agent/sr.py:288-308 (host-side convenience; same pairwise order is NOT guaranteed here)."""
        st = self._stream
        s = torch.as_tensor(state, device=st.device).reshape(-1).long().expand(st.n_agents)
        n = torch.arange(st.n_agents, device=st.device)
        values = (self._SR * self._rewards.unsqueeze(1)).sum(dim=2)          # [N, S]
        m = self._model[n, s].long()                                          # [N, A]
        return self._view(torch.gather(values, 1, m))

    def predict_on_batch(self, batch):
        idx = np.array(batch).astype(int).reshape(-1)
        out = torch.stack([self._stream.single and self.retrieve_q(int(s)) or self.retrieve_q(int(s)) for s in idx], dim=-2)
        return out.cpu().numpy() if self._stream.single else out
