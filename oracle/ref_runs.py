"""Drive the *unmodified* reference under a pre-drawn uniform stream.

Test infrastructure (see oracle/__init__.py); build container only (needs
/root/reference).  Every reference object of one run receives the same
``StreamRNG`` as ``rng=``; trajectories are captured through the reference's
own callback hooks (``on_step_end``, ``on_trial_end``, ``on_replay_end``).
Returns records shaped like ``oracle.tabular.Record.arrays()`` plus the final
tables, so reference, restatement and CUDA outputs compare field by field.
"""
import numpy as np

from . import ref_loader
from .stream_rng import StreamRNG


def _policy(cobel, spec, rng):
    kind, par = spec
    cls = {'eps': cobel.policy.EpsilonGreedy, 'xeps': cobel.policy.ExclusiveEpsilonGreedy,
           'softmax': cobel.policy.Softmax}[kind]
    return cls(par, rng=rng)


class _Capture:
    def __init__(self, key=lambda s: s):
        self.s, self.a, self.s2, self.r = [], [], [], []
        self.trial_steps, self.trial_reward, self.replay, self.replay_len = [], [], [], []
        self.key = key

    def on_step_end(self, logs):
        self.s.append(self.key(logs['state'])); self.a.append(int(logs['action']))
        self.s2.append(self.key(logs['next_state'])); self.r.append(float(logs['reward']))

    def on_trial_end(self, logs):
        self.trial_steps.append(int(logs['steps'])); self.trial_reward.append(float(logs['trial_reward']))

    def callbacks(self):
        return {'on_step_end': [self.on_step_end], 'on_trial_end': [self.on_trial_end]}

    def arrays(self):
        return {
            'states': np.array(self.s, dtype=np.int32), 'actions': np.array(self.a, dtype=np.int32),
            'next_states': np.array(self.s2, dtype=np.int32), 'rewards': np.array(self.r, dtype=np.float64),
            'trial_steps': np.array(self.trial_steps, dtype=np.int32),
            'trial_reward': np.array(self.trial_reward, dtype=np.float64),
            'replay': np.array(self.replay, dtype=np.int32),
            'replay_len': np.array(self.replay_len, dtype=np.int32),
        }


def _gridworld(cobel, world, rng):
    # NOTE Gridworld.__init__ performs one reset, i.e. consumes one draw (gridworld.py:89)
    return cobel.interface.Gridworld(world, rng=rng)


def run_dynaq(world, u, trials, steps, batch, *, policy=('eps', 0.1), policy_test=None, lr=0.99,
              gamma=0.99, mem_lr=0.9, mask_actions=False, no_replay=False, episodic_replay=False,
              test_trials=0, action_mask=None):
    cobel = ref_loader.load()
    rng = StreamRNG(u)
    env = _gridworld(cobel, world, rng)
    S = world['states']
    cap = _Capture()
    mem = cobel.memory.DynaQMemory(S, 4, mem_lr, rng=rng)
    # capture replay indices by wrapping retrieve_batch
    orig = mem.retrieve_batch

    def retrieve_batch(n=32):
        b = orig(n)
        cap.replay.extend(int(e['state']) * 4 + int(e['action']) for e in b)
        cap.replay_len.append(len(b))
        return b
    mem.retrieve_batch = retrieve_batch
    agent = cobel.agent.DynaQ(env.observation_space, env.action_space, _policy(cobel, policy, rng),
                              None if policy_test is None else _policy(cobel, policy_test, rng),
                              lr, gamma, mem, cap.callbacks())
    agent.mask_actions = mask_actions
    if action_mask is not None:
        agent.action_mask = np.array(action_mask, dtype=bool)
    agent.episodic_replay = episodic_replay
    agent.train(env, trials, steps, batch, no_replay)
    out = cap.arrays()
    out.update(Q=agent.Q.copy(), Mr=mem.rewards.copy(), Ms=mem.states.astype(np.int32),
               Mt=mem.terminals.astype(np.int32), draws=rng.k)
    if test_trials:
        cap2 = _Capture()
        agent.callbacks.custom_callbacks = cap2.callbacks()
        agent.test(env, test_trials, steps)
        t = cap2.arrays()
        out.update({'test_' + k: v for k, v in t.items() if not k.startswith('replay')})
        out['draws_after_test'] = rng.k
    return out


def run_q_gridworld(world, u, trials, steps, batch, *, policy=('eps', 0.1), lr=0.9, gamma=0.8):
    cobel = ref_loader.load()
    rng = StreamRNG(u)
    env = _gridworld(cobel, world, rng)
    S = world['states']
    cap = _Capture(key=lambda t: int(t[0]))
    agent = cobel.agent.QAgent(env.observation_space, env.action_space, _policy(cobel, policy, rng),
                               None, lr, gamma, cap.callbacks(), rng=rng)
    _wrap_q_replay(agent, cap)
    agent.train(env, trials, steps, batch)
    out = cap.arrays()
    Q = np.zeros((S, 4))
    for k, v in agent.Q.items():
        Q[int(k[0])] = v
    out.update(Q=Q, draws=rng.k, log_len=len(agent.M))
    return out


def _wrap_q_replay(agent, cap):
    orig_choice = agent.rng.choice

    class _R:
        def __getattr__(self, name):
            return getattr(agent_rng, name)

        def choice(self, a, size=None, p=None):
            idx = orig_choice(a, size, p)
            cap.replay.extend(int(i) for i in np.atleast_1d(idx))
            cap.replay_len.append(len(np.atleast_1d(idx)))
            return idx
    agent_rng = agent.rng
    agent.rng = _R()


def run_q_topology(nodes, starting_nodes, u, trials, steps, batch, *, policy=('eps', 0.1), lr=0.9, gamma=0.8):
    """QAgent on Topology with pose observations (agent/q.py:151-158)."""
    cobel = ref_loader.load()
    rng = StreamRNG(u)
    env = cobel.interface.Topology(nodes, starting_nodes, rng=rng)   # one draw in __init__
    ids = list(nodes.keys())
    pose_to_idx = {tuple(np.array(nodes[n]['pose']).flatten()): i for i, n in enumerate(ids)}
    cap = _Capture(key=lambda t: pose_to_idx[tuple(t)])
    agent = cobel.agent.QAgent(env.observation_space, env.action_space, _policy(cobel, policy, rng),
                               None, lr, gamma, cap.callbacks(), rng=rng)
    _wrap_q_replay(agent, cap)
    agent.train(env, trials, steps, batch)
    out = cap.arrays()
    Q = np.zeros((len(ids), int(env.action_space.n)))
    for k, v in agent.Q.items():
        Q[pose_to_idx[tuple(k)]] = v
    out.update(Q=Q, draws=rng.k, log_len=len(agent.M))
    return out


def run_sr(world, u, trials, steps, *, policy=('eps', 0.1), lr=0.1, gamma=0.99, mask_actions=False,
           action_mask=None):
    cobel = ref_loader.load()
    rng = StreamRNG(u)
    env = _gridworld(cobel, world, rng)
    cap = _Capture()
    agent = cobel.agent.SR(env.observation_space, env.action_space, _policy(cobel, policy, rng),
                           None, lr, gamma, cap.callbacks())
    agent.mask_actions = mask_actions
    if action_mask is not None:
        agent.action_mask = np.array(action_mask, dtype=bool)
    agent.train(env, trials, steps)
    out = cap.arrays()
    out.update(SR=agent.SR.copy(), rew=agent.rewards.copy(),
               model=np.argmax(agent.transitions, axis=2).astype(np.int32), draws=rng.k)
    return out


def run_sfma(world, D, u, trials, steps, batch, *, policy=('eps', 0.1), lr=0.99, gamma=0.99,
             mem_lr=0.9, mask_actions=False, mode='default', recency=False, start_replay=False,
             nb_replays=1, metric=None, action_mask=None, random_replay=False, dynamic=False, mem_flags=None):
    cobel = ref_loader.load()
    rng = StreamRNG(u)
    env = _gridworld(cobel, world, rng)
    S = world['states']
    cap = _Capture()

    class _Metric:
        pass
    if metric is None:
        metric = _Metric()
        metric.D = D
    mem = cobel.memory.SFMAMemory(metric, S, 4, learning_rate=mem_lr, rng=rng)
    mem.mode = mode
    mem.recency = recency
    for k, v in (mem_flags or {}).items():      # reward_mod_local, reward_mod, state_mod, C_normalize, ...
        assert hasattr(mem, k), k
        setattr(mem, k, v)
    cbs = cap.callbacks()

    def on_replay_end(logs):
        cap.replay.extend(int(e['action']) * S + int(e['state']) for e in logs['replay'])
        cap.replay_len.append(len(logs['replay']))
    cbs['on_replay_end'] = [on_replay_end]
    agent = cobel.agent.SFMA(env.observation_space, env.action_space, _policy(cobel, policy, rng),
                             mem, None, lr, gamma, cbs, rng=rng)
    agent.mask_actions = mask_actions
    if action_mask is not None:
        agent.action_mask = np.array(action_mask, dtype=bool)
    agent.start_replay = start_replay
    agent.random = random_replay
    agent.nb_replays = nb_replays
    agent.dynamic = dynamic
    modes = []
    if dynamic:
        cbs['on_trial_end'] = list(cbs.get('on_trial_end', [])) + [lambda logs: modes.append(['reverse', 'default'].index(logs['replay_mode']))]
    agent.train(env, trials, steps, batch)
    out = cap.arrays()
    if dynamic:
        out['modes'] = np.array(modes, dtype=np.int32)
        out['td'] = np.float64(agent.td)
    out.update(Q=agent.Q.copy(), Mr=mem.rewards.copy(), Ms=mem.states.astype(np.int32),
               Mt=mem.terminals.astype(np.int32), C=mem.C.copy(), T=mem.T.copy(), I=mem.I.copy(),
               draws=rng.k)
    return out


def run_pma(world, u, trials, steps, batch, *, policy=('eps', 0.1), mem_policy=('eps', 0.1), lr=0.9,
            gamma=0.99, mem_lr=0.9, lr_q=0.9, gamma_sr=0.9, gamma_q=0.99, mask_actions=True,
            prefill=False, min_gain_mode='original', action_mask=None, equal_need=False, equal_gain=False,
            ignore_barriers=True, allow_loops=False):
    cobel = ref_loader.load()
    rng = StreamRNG(u)
    env = _gridworld(cobel, world, rng)
    S = world['states']
    cap = _Capture()
    mem = cobel.memory.PMAMemory(world['sas'], _policy(cobel, mem_policy, rng), mem_lr, lr_q,
                                 gamma_sr, gamma_q, rng=rng)
    mem.min_gain_mode = min_gain_mode
    mem.equal_need, mem.equal_gain, mem.ignore_barriers = equal_need, equal_gain, ignore_barriers
    mem.allow_loops = allow_loops
    if prefill:   # unit_tests/test_pma.py:69-73
        for s in range(S):
            for a in range(4):
                mem.states[s, a] = np.argmax(world['sas'][s, a])
        mem.compute_update_mask()
    cbs = cap.callbacks()

    def on_replay_end(logs):
        cap.replay.extend(int(e['action']) * S + int(e['state']) for e in logs['replay'])
        cap.replay_len.append(len(logs['replay']))
    cbs['on_replay_end'] = [on_replay_end]
    agent = cobel.agent.PMA(env.observation_space, env.action_space, _policy(cobel, policy, rng),
                            mem, None, lr, gamma, cbs)
    agent.mask_actions = mask_actions
    if action_mask is not None:
        agent.action_mask = np.array(action_mask, dtype=bool)
    agent.train(env, trials, steps, batch)
    out = cap.arrays()
    out.update(Q=agent.Q.copy(), Mr=mem.rewards.copy(), Ms=mem.states.astype(np.int32),
               Mt=mem.terminals.astype(np.int32), T=mem.T.copy(), SR=mem.SR.copy(),
               update_mask=mem.update_mask.copy(), draws=rng.k)
    return out
