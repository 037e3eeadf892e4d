"""GPU: the Dyna hybrids (DynaDQN / DynaDSR, agent/dyna_q.py:333-1150) on the batched path against golden runs of
the UNMODIFIED reference classes (oracle/make_golden_hybrid.py: reference TorchNetwork on the CPU, fp64, Adam).

The golden agent is local agent 1 of a batch of three agents with different streams and different initial weights,
so the lock-step loop's masking is exercised: the other agents end their trials at other steps.  Integer data
(trajectory, replayed experiences, draw counts, memory tables) must be identical; the network outputs agree to the
accuracy of fp64 matrix products evaluated in a different order (1e-9 asserted, ~1e-13 seen)."""
import numpy as np
import pytest
import torch

from oracle import cases
from oracle.hybrid_models import seeded
from oracle.make_golden_hybrid import HYBRID_CASES, MODEL_SEED, WORLD
from helpers import load_golden, make_world

pytestmark = pytest.mark.gpu


class _Cap:
    def __init__(self, local):
        self.i, self.s, self.a, self.s2, self.r, self.steps, self.rew = local, [], [], [], [], [], []

    def step(self, logs):
        if bool(logs['active'][self.i]):
            self.s.append(int(logs['state'][self.i])); self.a.append(int(logs['action'][self.i]))
            self.s2.append(int(logs['next_state'][self.i])); self.r.append(float(logs['reward'][self.i]))

    def trial(self, logs):
        self.steps.append(int(logs['steps'][self.i])); self.rew.append(float(logs['trial_reward'][self.i]))


@pytest.mark.parametrize('name', sorted(HYBRID_CASES))
def test_hybrid_matches_reference_golden(name):
    import cobel_rl_b200 as cb
    from cobel_rl_b200.agent import DynaDQN, DynaDSR
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.memory import DynaQMemory
    from cobel_rl_b200.network import BatchedTorchNetwork
    from cobel_rl_b200.policy import EpsilonGreedy
    from oracle.tabular import valid_move_mask
    kind, agent_id, kw = HYBRID_CASES[name]
    gold = load_golden(name)
    n, local = 3, 1
    base = agent_id - local
    stream = cb.BatchStream(n, seed=cases.SEED, device='cuda:0', agent_id_base=base)
    env = Gridworld(make_world(WORLD), rng=stream)
    mem = DynaQMemory(env.n_states, 4, 0.9, rng=stream)
    pol, pol_test = EpsilonGreedy(0.1, rng=stream), EpsilonGreedy(0.0, rng=stream)
    cap = _Cap(local)
    cbs = {'on_step_end': [cap.step], 'on_trial_end': [cap.trial]}
    if kind == 'dqn':
        net = BatchedTorchNetwork([seeded(4, MODEL_SEED + base + i) for i in range(n)], device='cuda:0')
        ag = DynaDQN(env.observation_space, env.action_space, pol, net, policy_test=pol_test, gamma=0.9, memory=mem,
                     custom_callbacks=cbs)
        ag.DDQN, ag.mask_actions = kw['ddqn'], kw['mask']
        if kw['mask']:
            ag.action_mask = valid_move_mask(env._succ.cpu().numpy())
    else:
        sr = BatchedTorchNetwork([seeded(25, MODEL_SEED + base + i) for i in range(n)], device='cuda:0')
        rw = BatchedTorchNetwork([seeded(1, MODEL_SEED + 100 + base + i) for i in range(n)], device='cuda:0')
        ag = DynaDSR(env.observation_space, env.action_space, pol, sr, rw, policy_test=pol_test, gamma=0.9, memory=mem,
                     custom_callbacks=cbs)
        ag.use_DR, ag.use_follow_up_state, ag.ignore_terminality = kw['use_DR'], kw['use_follow_up_state'], kw['ignore_terminality']
    ag.target_update = kw['target_update']
    ag.train(env, kw['trials'], kw['steps'], kw['batch'])
    torch.cuda.synchronize()
    nt = int(gold['n_train_steps'])
    assert cap.s == gold['states'][:nt].tolist() and cap.a == gold['actions'][:nt].tolist(), name
    assert cap.s2 == gold['next_states'][:nt].tolist() and cap.r == gold['rewards'][:nt].tolist()
    assert int(stream.draw_count[local]) == int(gold['draws_train'])
    assert np.array_equal(mem.rewards[local].cpu().numpy(), gold['Mr']) and np.array_equal(mem.states[local].cpu().numpy(), gold['Ms'])
    assert np.array_equal(mem.terminals[local].cpu().numpy(), gold['Mt'])
    q = ag.predict_on_batch(np.arange(env.n_states))[local].cpu().numpy()
    err = np.abs(q - gold['q_train']).max() / np.abs(gold['q_train']).max()
    assert err < 1e-9, 'Q predictions differ from the reference by %.3e (relative to the largest value)' % err
    ag.test(env, 2, kw['steps'])
    torch.cuda.synchronize()
    assert cap.s == gold['states'].tolist() and cap.a == gold['actions'].tolist()
    assert cap.steps == gold['trial_steps'].tolist() and cap.rew == gold['trial_reward'].tolist()
    assert int(stream.draw_count[local]) == int(gold['draws'])
    # the other agents learned their own networks (different streams and initial weights)
    qa = ag.predict_on_batch(np.arange(env.n_states))
    assert float((qa[0] - qa[local]).abs().max()) > 1e-6 and float((qa[2] - qa[local]).abs().max()) > 1e-6


def _hybrid_env(n=2, seed=5):
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.policy import EpsilonGreedy
    stream = cb.BatchStream(n, seed=seed, device='cuda:0')
    env = Gridworld(make_world(WORLD), rng=stream)
    return stream, env, EpsilonGreedy(0.1, rng=stream), EpsilonGreedy(0.0, rng=stream)


def test_dyna_dqn_option_sweep_like_the_reference_unit_test():
    """unit_tests/test_dyna_dqn.py:36-87: every combination of test policy, action mask, DDQN and target-update rule
    runs through train() and test(); additionally the Q predictions stay finite and the memory has been filled."""
    from itertools import product
    from cobel_rl_b200.agent import DynaDQN
    from cobel_rl_b200.network import BatchedTorchNetwork
    for use_test_policy, mask_actions, ddqn, target_update in product([True, False], [True, False], [True, False], [0.001, 10]):
        stream, env, pol, pol_test = _hybrid_env()
        net = BatchedTorchNetwork([seeded(4, 10 + i) for i in range(2)], device='cuda:0')
        ag = DynaDQN(env.observation_space, env.action_space, pol, net, policy_test=pol_test if use_test_policy else None)
        ag.target_update, ag.mask_actions, ag.DDQN = target_update, mask_actions, ddqn
        ag.train(env, 2, 8, 8)
        ag.test(env, 1, 8)
        q = ag.predict_on_batch(np.arange(25))
        assert q.shape == (2, 25, 4) and bool(torch.isfinite(q).all())
        assert int((ag.M.terminals != 0).sum()) > 0 and ag.current_trial == 3


def test_dyna_dsr_option_sweep_like_the_reference_unit_test():
    """unit_tests/test_dyna_dsr.py:34-100: DR / follow-up state / terminality / target-update combinations."""
    from itertools import product
    from cobel_rl_b200.agent import DynaDSR
    from cobel_rl_b200.network import BatchedTorchNetwork
    for mask_actions, use_dr, follow_up, ignore_term, target_update in product([True, False], [True, False], [True, False],
                                                                             [True, False], [0.001, 10]):
        stream, env, pol, pol_test = _hybrid_env()
        sr = BatchedTorchNetwork([seeded(25, 20 + i) for i in range(2)], device='cuda:0')
        rw = BatchedTorchNetwork([seeded(1, 40 + i) for i in range(2)], device='cuda:0')
        ag = DynaDSR(env.observation_space, env.action_space, pol, sr, rw, policy_test=pol_test)
        ag.target_update, ag.mask_actions = target_update, mask_actions
        ag.use_DR, ag.use_follow_up_state, ag.ignore_terminality = use_dr, follow_up, ignore_term
        ag.train(env, 2, 6, 8)
        ag.test(env, 1, 6)
        q = ag.predict_on_batch(np.arange(25))
        assert q.shape == (2, 25, 4) and bool(torch.isfinite(q).all())
