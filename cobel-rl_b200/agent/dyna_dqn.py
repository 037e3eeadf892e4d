"""Dyna hybrids: tabular Dyna memory + deep function approximation (reference: agent/dyna_q.py:333-1150).

``DynaDQN`` (agent/dyna_q.py:333-708) and ``DynaDSR`` (:711-1150) sample replay batches from the tabular
``DynaQMemory`` and train PyTorch networks on them.  On the batched path the table side -- ``Interface.step``,
``Policy.select_action``, ``DynaQMemory.store`` / ``retrieve_batch`` -- is one CUDA launch each for ALL N agents
(csrc/ops.cu), the per-agent random streams are the ``BatchStream`` contract of the tabular agents, and the
networks are N independent copies evaluated together (``network.BatchedTorchNetwork``).  The Python loop below is
the reference's ``train`` / ``test`` loop with every variable carrying an agent axis; the agents of a trial advance
in lock-step and an agent whose trial has ended idles until the others finish (it consumes no draws, stores
nothing and its networks and optimizer state are not touched), so every agent follows exactly the trajectory of
a single-agent run under its own stream.

Callbacks: all hooks of the reference fire, per step, with batched values (``[N]`` tensors; ``logs['active']``
marks the agents still inside the trial).
"""
import numpy as np
import torch

from ..memory.dyna_q import DynaQMemory
from ..spaces import Discrete
from .agent import Agent


class _HybridBase(Agent):
    """The trial x step loop shared by the two hybrids (agent/dyna_q.py:490-622, 888-1020)."""

    def _bind_stream(self, *candidates):
        stream = self._find_stream(*candidates)
        assert stream is not None, 'the policy or the memory must carry the BatchStream (rng=...)'
        self._stream = stream
        self.M._allocate(stream)
        return stream

    def _allocate(self, stream):
        pass

    # -- masked single steps of the batched stand-alone API ------------------------------------------------
    def _restore_draws(self, before, active):
        st = self._stream
        st.draw_count.copy_(torch.where(active, st.draw_count, before))

    def _select(self, policy, q, state, active):
        st = self._stream
        before = st.draw_count.clone()
        mask = self._action_mask[state] if self.mask_actions else None
        a = policy.select_action(q, mask)
        a = torch.as_tensor(a, device=st.device).reshape(-1).long()
        self._restore_draws(before, active)
        return a

    def _env_step(self, interface, action, active):
        st = self._stream
        before, cur = st.draw_count.clone(), interface._current.clone()
        nxt, reward, end, _, _ = interface.step(action)
        nxt = torch.as_tensor(nxt, device=st.device).reshape(-1).long()
        reward = torch.as_tensor(reward, device=st.device, dtype=torch.float64).reshape(-1)
        end = torch.as_tensor(end, device=st.device).reshape(-1).bool()
        interface._current = torch.where(active, interface._current, cur)
        self._restore_draws(before, active)
        return nxt, reward, end

    def _store(self, state, action, reward, nxt, end, active):
        """M.store for the active agents; the others re-store what the table already holds (an exact no-op:
        ``r + lr * (r - r) == r``, memory/dyna_q.py:92-96)."""
        M, n = self.M, torch.arange(self._stream.n_agents, device=self._stream.device)
        r0, s0, t0 = M._rewards[n, state, action], M._states[n, state, action].long(), M._terminals[n, state, action]
        M.store({'state': state, 'action': action, 'reward': torch.where(active, reward, r0),
                 'next_state': torch.where(active, nxt, s0),
                 'terminal': torch.where(active, 1 - end.to(torch.int32), t0.to(torch.int32))})

    def _retrieve(self, batch_size, active):
        st = self._stream
        before = st.draw_count.clone()
        batch = self.M._retrieve(batch_size)
        self._restore_draws(before, active)
        return batch

    def _reset(self, interface):
        state, _ = interface.reset()
        return torch.as_tensor(state, device=self._stream.device).reshape(-1).long()

    def _run(self, interface, trials, steps, batch_size, no_replay, learn):
        st = self._stream
        assert interface.rng is st, 'environment and agent must share one BatchStream'
        n, dev = st.n_agents, st.device
        policy = self.policy if learn else self.policy_test
        for trial in range(trials):
            logs = self.callbacks.on_trial_begin({'trial_reward': torch.zeros(n, dtype=torch.float64, device=dev),
                                                  'trial': self.current_trial, 'trial_session': trial})
            state = self._reset(interface)
            active = torch.ones(n, dtype=torch.bool, device=dev)
            last_step = torch.zeros(n, dtype=torch.int64, device=dev)
            for step in range(steps):
                logs['step'], logs['active'] = step, active
                logs = self.callbacks.on_step_begin(logs)
                action = self._select(policy, self.retrieve_q(state), state, active)
                nxt, reward, end = self._env_step(interface, action, active)
                if learn:
                    self._store(state, action, reward, nxt, end, active)
                experience = {'state': state, 'action': action, 'reward': reward, 'next_state': nxt,
                              'terminal': 1 - end.to(torch.int32)}
                state = torch.where(active, nxt, state)
                if learn and not no_replay and not self.episodic_replay:
                    self.replay(batch_size, active)
                logs['trial_reward'] = logs['trial_reward'] + torch.where(active, reward, torch.zeros_like(reward))
                logs.update(experience)
                logs = self.callbacks.on_step_end(logs)
                last_step = torch.where(active, torch.full_like(last_step, step), last_step)
                active = active & ~end
                if not bool(active.any()):
                    break
            self.current_trial += 1
            logs['steps'] = last_step
            if learn and not no_replay and self.episodic_replay:
                self.replay(batch_size)
            logs = self.callbacks.on_trial_end(logs)
            if self.stop:
                break

    def train(self, interface, trials, steps, batch_size=32, no_replay=False):
        """agent/dyna_q.py:490-564 / 888-962 for all agents."""
        self._run(interface, trials, steps, batch_size, no_replay, learn=True)

    def test(self, interface, trials, steps):
        """agent/dyna_q.py:566-622 / 964-1020 for all agents."""
        self._run(interface, trials, steps, 0, True, learn=False)

    @property
    def action_mask(self):
        return self._action_mask

    @action_mask.setter
    def action_mask(self, value):
        self._action_mask = torch.as_tensor(value, device=self._stream.device).bool().contiguous()

    def _observe(self, idx):
        """``self.observations[idx]`` for an index tensor ``[N, ...]`` -> ``[N, ..., obs]``."""
        return self._observations[idx]

    def _blend_targets(self, online, target, active):
        """agent/dyna_q.py:690-708: soft (``target_update < 1``) or periodic hard update of a target network."""
        act = torch.ones(self._stream.n_agents, dtype=torch.bool, device=self._stream.device) if active is None else active
        if self.target_update < 1.0:
            wt, wo = target.get_weights(), online.get_weights()
            target.set_weights([t + self.target_update * (o - t) for t, o in zip(wt, wo)], active=act)
            return None
        return act & (self.last_update == self.target_update)


class DynaDQN(_HybridBase):
    def __init__(self, observation_space, action_space, policy, model, observations=None, policy_test=None, gamma=0.99,
                 memory=None, custom_callbacks=None):
        assert type(observation_space) is Discrete, 'DynaDQN requires a discrete observation space!'
        assert type(action_space) is Discrete, 'DynaDQN requires a discrete action space!'
        super().__init__(observation_space, action_space, custom_callbacks)
        self.policy = policy
        self.policy_test = policy if policy_test is None else policy_test
        self.model_target = model                       # agent/dyna_q.py:467-469
        self.model_online = self.model_target.clone()
        S, A = int(observation_space.n), int(action_space.n)
        self.M = DynaQMemory(S, A) if memory is None else memory
        st = self._bind_stream(self.policy, self.policy_test, self.M)
        assert model.n_agents == st.n_agents, 'one network copy per agent'
        obs = np.eye(S) if observations is None else observations
        self._observations = torch.as_tensor(np.asarray(obs), device=st.device)
        self.target_update = 10 ** -2
        self.last_update = torch.zeros(st.n_agents, dtype=torch.int64, device=st.device)
        self.DDQN = False
        self.gamma = gamma
        self._action_mask = torch.ones((S, A), dtype=torch.bool, device=st.device)
        self.mask_actions = False
        self.episodic_replay = False

    observations = property(lambda self: self._observations)

    def retrieve_q(self, state):
        """agent/dyna_q.py:624-640: Q-values of each agent's current state, ``[N, A]``."""
        s = torch.as_tensor(state, device=self._stream.device).reshape(-1, 1).long()
        return self.model_online.predict_on_batch(self._observe(s))[:, 0]

    def predict_on_batch(self, batch):
        """agent/dyna_q.py:642-658: ``[N, B, A]`` for a batch of B states (the same for every agent)."""
        idx = torch.as_tensor(np.array(batch).astype(int), device=self._stream.device).reshape(1, -1)
        return self.model_online.predict_on_batch(self._observe(idx.expand(self._stream.n_agents, -1)))

    def replay(self, batch_size, active=None):
        """agent/dyna_q.py:660-708 for all (``active``) agents."""
        st = self._stream
        act = torch.ones(st.n_agents, dtype=torch.bool, device=st.device) if active is None else active
        b = self._retrieve(batch_size, act)
        states, next_states = self._observe(b.state.long()), self._observe(b.next_state.long())
        targets = self.model_online.predict_on_batch(states)
        boot = self.model_target.predict_on_batch(next_states)
        pick = (self.model_online.predict_on_batch(next_states) if self.DDQN else boot).argmax(dim=2)
        boot = boot.gather(2, pick.unsqueeze(-1)).squeeze(-1)
        gamma = st.param(self.gamma, 'gamma').reshape(-1, 1).to(boot.dtype)
        value = b.reward.to(boot.dtype) + boot * (b.terminal != 0).to(boot.dtype) * gamma
        # experiences of a batch that share (state, action) write the same slot in order: the last one wins
        # (targets[arange, actions] = ..., agent/dyna_q.py:684-686); states are one-hot rows here, so the slot is
        # per sample and there is nothing to resolve
        targets.scatter_(2, b.action.long().unsqueeze(-1), value.unsqueeze(-1))
        self.model_online.train_on_batch(states, targets, active=act)
        self.last_update = self.last_update + act.long()
        hard = self._blend_targets(self.model_online, self.model_target, act)
        if hard is not None and bool(hard.any()):
            self.model_target.set_weights(self.model_online.get_weights(), active=hard)
            self.last_update = torch.where(hard, torch.zeros_like(self.last_update), self.last_update)


class DynaDSR(_HybridBase):
    def __init__(self, observation_space, action_space, policy, model_sr, model_reward, observations=None,
                 policy_test=None, gamma=0.99, memory=None, custom_callbacks=None):
        assert type(observation_space) is Discrete, 'DynaDSR requires a discrete observation space!'
        assert type(action_space) is Discrete, 'DynaDSR requires a discrete action space!'
        super().__init__(observation_space, action_space, custom_callbacks)
        self.policy = policy
        self.policy_test = policy if policy_test is None else policy_test
        S, A = int(observation_space.n), int(action_space.n)
        self.models_target = {a: model_sr.clone() for a in range(A)}      # agent/dyna_q.py:857-863
        self.models_online = {a: model_sr.clone() for a in range(A)}
        self.model_reward = model_reward
        self.M = DynaQMemory(S, A) if memory is None else memory
        st = self._bind_stream(self.policy, self.policy_test, self.M)
        assert model_sr.n_agents == st.n_agents and model_reward.n_agents == st.n_agents, 'one network copy per agent'
        obs = np.eye(S) if observations is None else observations
        self._observations = torch.as_tensor(np.asarray(obs), device=st.device)
        self.target_update = 10 ** -2
        self.last_update = torch.zeros(st.n_agents, dtype=torch.int64, device=st.device)
        self.use_DR = False
        self.use_follow_up_state = False
        self.ignore_terminality = True
        self.gamma = gamma
        self._action_mask = torch.ones((S, A), dtype=torch.bool, device=st.device)
        self.mask_actions = False
        self.episodic_replay = False

    observations = property(lambda self: self._observations)

    def _q(self, obs):
        """``model_reward(model_a(obs))`` for every action: ``[N, B, A]`` (agent/dyna_q.py:1022-1062)."""
        cols = [self.model_reward.predict_on_batch(m.predict_on_batch(obs))[..., 0] for m in self.models_online.values()]
        return torch.stack(cols, dim=-1)

    def retrieve_q(self, state):
        s = torch.as_tensor(state, device=self._stream.device).reshape(-1, 1).long()
        return self._q(self._observe(s))[:, 0]

    def predict_on_batch(self, batch):
        idx = torch.as_tensor(np.array(batch).astype(int), device=self._stream.device).reshape(1, -1)
        return self._q(self._observe(idx.expand(self._stream.n_agents, -1)))

    def replay(self, batch_size, active=None):
        """agent/dyna_q.py:1064-1150 for all (``active``) agents."""
        st = self._stream
        act = torch.ones(st.n_agents, dtype=torch.bool, device=st.device) if active is None else active
        b = self._retrieve(batch_size, act)
        states, next_states = self._observe(b.state.long()), self._observe(b.next_state.long())
        dt = states.dtype
        future_sr = {a: m.predict_on_batch(next_states) for a, m in self.models_target.items()}
        future_val = torch.stack([self.model_reward.predict_on_batch(sr)[..., 0] for sr in future_sr.values()], dim=0)
        nonterm = (b.terminal != 0).to(dt)                                   # bool(experience['terminal'])
        ufs, ign = float(self.use_follow_up_state), float(self.ignore_terminality)
        boot = next_states * ((1 - ufs) * (1 - ign)) * (1 - nonterm).unsqueeze(-1)
        gate = torch.clamp(nonterm + ign, max=1.0).unsqueeze(-1)
        srs = torch.stack(list(future_sr.values()), dim=0)                   # [A, N, B, F]
        if not self.use_DR:                                                  # Deep SR: the best follow-up action's stream
            best = future_val.argmax(dim=0)                                  # [N, B]
            pick = srs.gather(0, best.reshape((1,) + best.shape + (1,)).expand((1,) + srs.shape[1:]))[0]
            boot = boot + pick * gate
        else:                                                                # Deep DR: the mean over the actions
            boot = boot + srs.mean(dim=0) * gate
        gamma = st.param(self.gamma, 'gamma').reshape(-1, 1, 1).to(dt)
        targets = (next_states if self.use_follow_up_state else states) + gamma * boot
        for a, model in self.models_online.items():                          # one sub-batch per action
            model.train_on_batch(states, targets, active=act, sample_mask=(b.action == a))
        self.model_reward.train_on_batch(next_states, b.reward.to(dt), active=act)
        self.last_update = self.last_update + act.long()
        hard = None
        for a in self.models_online:
            hard = self._blend_targets(self.models_online[a], self.models_target[a], act)
            if hard is not None and bool(hard.any()):
                self.models_target[a].set_weights(self.models_online[a].get_weights(), active=hard)
        if hard is not None:
            self.last_update = torch.where(hard, torch.zeros_like(self.last_update), self.last_update)
