"""GPU: the stand-alone methods of the class API (csrc/ops.cu and the replay phases of the fused kernels).

Each test drives the reference's train loop BY HAND -- env.reset / env.step, policy.select_action, M.store,
agent.update_q, agent.replay, one call at a time, exactly as agent/dyna_q.py:176-203, agent/q.py:186-221,
agent/sr.py:166-190, agent/pma.py:199-257 and agent/sfma.py:266-327 do -- on a single-agent stream and must land
on the same golden vector (generated from the unmodified reference) as the fused ``train()``."""
import numpy as np
import pytest
import torch

from oracle import cases, tabular as tb
from helpers import KEYS, assert_equal_records, load_golden, make_topology, make_world

pytestmark = pytest.mark.gpu


def _setup(name):
    import cobel_rl_b200 as cb
    kind, wname, case_agent, args = cases.CASES[name]
    a = dict(args)
    stream = cb.BatchStream(None, seed=cases.SEED, device='cuda:0', agent_id_base=case_agent)
    return kind, wname, a, stream


def _policy(spec, stream):
    from cobel_rl_b200 import policy as P
    kind, par = spec
    return {'eps': P.EpsilonGreedy, 'xeps': P.ExclusiveEpsilonGreedy, 'softmax': P.Softmax}[kind](par, rng=stream)


class _Trace:
    def __init__(self):
        self.s, self.a, self.s2, self.r, self.steps, self.rew = [], [], [], [], [], []

    def arrays(self):
        return {'states': np.array(self.s, dtype=np.int32), 'actions': np.array(self.a, dtype=np.int32),
                'next_states': np.array(self.s2, dtype=np.int32), 'rewards': np.array(self.r, dtype=np.float64),
                'trial_steps': np.array(self.steps, dtype=np.int32), 'trial_reward': np.array(self.rew, dtype=np.float64)}


TRAJ = ['states', 'actions', 'next_states', 'rewards', 'trial_steps', 'trial_reward']


@pytest.mark.parametrize('name', ['dynaq_open5_eps', 'dynaq_walls5_mask', 'dynaq_open5_softmax', 'dynaq_slip5'])
def test_dynaq_hand_written_loop_matches_golden(name):
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import DynaQ
    from cobel_rl_b200.memory import DynaQMemory
    kind, wname, a, stream = _setup(name)
    env = Gridworld(make_world(wname), rng=stream)
    pol = _policy(a.get('policy', ('eps', 0.1)), stream)
    mem = DynaQMemory(env.n_states, 4, a.get('mem_lr', 0.9), rng=stream)
    ag = DynaQ(env.observation_space, env.action_space, pol, None, a.get('lr', 0.99), a.get('gamma', 0.99), mem)
    mask = torch.as_tensor(tb.valid_move_mask(env._succ.cpu().numpy()), device='cuda:0') if a.get('valid_mask') else None
    tr = _Trace()
    for _ in range(a['trials']):
        state, _info = env.reset()
        total, step = 0.0, 0
        for step in range(a['steps']):
            action = pol.select_action(ag.Q[state], mask[state] if a.get('mask_actions') else None)
            nxt, reward, end, _trunc, _info = env.step(action)
            exp = {'state': state, 'action': action, 'reward': reward, 'next_state': nxt, 'terminal': 1 - int(end)}
            mem.store(exp)
            out = ag.update_q(exp)
            assert 'td' in out
            tr.s.append(state); tr.a.append(action); tr.s2.append(nxt); tr.r.append(reward)
            state = nxt
            ag.replay(a['batch'])
            total += reward
            if end:
                break
        tr.steps.append(step); tr.rew.append(total)
    torch.cuda.synchronize()
    got = tr.arrays()
    got.update(Q=ag.Q.cpu().numpy(), Mr=mem.rewards.cpu().numpy(), Ms=mem.states.cpu().numpy(), Mt=mem.terminals.cpu().numpy(),
               draws=int(stream.draw_count[0]))
    assert_equal_records(got, load_golden(name), TRAJ + ['Q', 'Mr', 'Ms', 'Mt', 'draws'], what=name)


def test_qagent_hand_written_loop_on_topology_matches_golden():
    from cobel_rl_b200.interface import Topology
    from cobel_rl_b200.agent import QAgent
    name = 'q_track'
    kind, wname, a, stream = _setup(name)
    env = Topology(*make_topology(wname), rng=stream)
    pol = _policy(a.get('policy', ('eps', 0.1)), stream)
    ag = QAgent(env.observation_space, env.action_space, pol, None, a.get('lr', 0.9), a.get('gamma', 0.8), rng=stream)
    ag.bind_interface(env)
    key = env._obs_key
    tr = _Trace()
    for _ in range(a['trials']):
        obs, _info = env.reset()
        assert np.asarray(obs.cpu()).shape == (6,)          # a pose (interface/topology.py:174-193)
        node = int(env._current[0])
        total, step = 0.0, 0
        for step in range(a['steps']):
            action = pol.select_action(ag.Q[int(key[node])])
            obs, reward, end, trunc, _info = env.step(action)
            assert end == trunc
            assert torch.equal(torch.as_tensor(obs), env.get_observation())
            nxt = int(env._current[0])
            exp = {'state': node, 'action': action, 'reward': reward, 'next_state': nxt, 'terminal': 1 - int(end)}
            ag.append(exp)
            ag.update_q(exp)
            tr.s.append(node); tr.a.append(action); tr.s2.append(nxt); tr.r.append(reward)
            node = nxt
            ag.replay(a['batch'])
            total += reward
            if end:
                break
        tr.steps.append(step); tr.rew.append(total)
    torch.cuda.synchronize()
    got = tr.arrays()
    got.update(Q=ag.Q.cpu().numpy(), draws=int(stream.draw_count[0]))
    assert_equal_records(got, load_golden(name), TRAJ + ['Q', 'draws'], what=name)
    assert int(ag._log_len[0]) == len(tr.s)


@pytest.mark.parametrize('name', ['sr_open5', 'sr_walls5_mask'])
def test_sr_hand_written_loop_matches_golden(name):
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import SR
    kind, wname, a, stream = _setup(name)
    env = Gridworld(make_world(wname), rng=stream)
    pol = _policy(a.get('policy', ('eps', 0.1)), stream)
    ag = SR(env.observation_space, env.action_space, pol, None, a.get('lr', 0.1), a.get('gamma', 0.99))
    mask = torch.as_tensor(tb.valid_move_mask(env._succ.cpu().numpy()), device='cuda:0') if a.get('valid_mask') else None
    tr = _Trace()
    for _ in range(a['trials']):
        state, _info = env.reset()
        total, step = 0.0, 0
        for step in range(a['steps']):
            action = pol.select_action(ag.retrieve_q(state), mask[state] if a.get('mask_actions') else None)
            nxt, reward, end, _trunc, _info = env.step(action)
            ag.update({'state': state, 'action': action, 'reward': reward, 'next_state': nxt, 'terminal': 1 - int(end)})
            tr.s.append(state); tr.a.append(action); tr.s2.append(nxt); tr.r.append(reward)
            state = nxt
            total += reward
            if end:
                break
        tr.steps.append(step); tr.rew.append(total)
    torch.cuda.synchronize()
    got = tr.arrays()
    got.update(SR=ag.SR.cpu().numpy(), rew=ag.rewards.cpu().numpy(), model=ag.model.cpu().numpy(), draws=int(stream.draw_count[0]))
    assert_equal_records(got, load_golden(name), TRAJ + ['SR', 'rew', 'model', 'draws'], what=name)


@pytest.mark.parametrize('name', ['pma_walls5', 'pma_walls5_timeout'])
def test_pma_hand_written_loop_matches_golden(name):
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import PMA
    from cobel_rl_b200.memory import PMAMemory
    kind, wname, a, stream = _setup(name)
    world = make_world(wname)
    env = Gridworld(world, rng=stream)
    pol = _policy(a.get('policy', ('eps', 0.1)), stream)
    mem = PMAMemory(world['sas'], _policy(('eps', 0.1), stream), 0.9, 0.9, 0.9, a.get('gamma_q', 0.9), rng=stream)
    ag = PMA(env.observation_space, env.action_space, pol, mem, None, 0.9, 0.99)
    ag.mask_actions = a.get('mask_actions', False)
    if a.get('valid_mask'):
        ag.action_mask = tb.valid_move_mask(env._succ.cpu().numpy())
    mask = ag.action_mask if ag.mask_actions else None
    tr = _Trace()
    replays, lens = [], []

    def replay(cur, update_sr=False):
        upd, q = mem.replay(ag, mask, a['batch'], cur, update_sr=update_sr)
        assert q is ag.Q or torch.equal(q, ag.Q)
        u = upd.cpu().numpy()
        u = u[u >= 0]
        replays.extend(u.tolist()); lens.append(len(u))

    for _ in range(a['trials']):
        state, _info = env.reset()
        replay(state)
        total, step, last = 0.0, 0, None
        for step in range(a['steps']):
            action = pol.select_action(ag.Q[state], mask[state] if mask is not None else None)
            nxt, reward, end, _trunc, _info = env.step(action)
            exp = {'state': state, 'action': action, 'reward': reward, 'next_state': nxt, 'terminal': 1 - int(end)}
            ag.update_q([exp])          # the reference updates Q before it stores (agent/pma.py:233-234)
            mem.store(exp)
            tr.s.append(state); tr.a.append(action); tr.s2.append(nxt); tr.r.append(reward)
            state = nxt
            total += reward
            if end:
                last = nxt
                break
        tr.steps.append(step); tr.rew.append(total)
        replay(last, update_sr=True)    # M.update_sr(), then replay from the terminal state (None: stationary need)
    torch.cuda.synchronize()
    got = tr.arrays()
    got.update(Q=ag.Q.cpu().numpy(), Mr=mem.rewards.cpu().numpy(), Ms=mem.states.cpu().numpy(), Mt=mem.terminals.cpu().numpy(),
               T=mem.T.cpu().numpy(), SR=mem.SR.cpu().numpy(), replay=np.array(replays, dtype=np.int32),
               replay_len=np.array(lens, dtype=np.int32), draws=int(stream.draw_count[0]))
    assert_equal_records(got, load_golden(name), TRAJ + ['replay', 'replay_len', 'Q', 'Mr', 'Ms', 'Mt', 'T', 'SR', 'draws'],
                         rtol={'SR': 1e-12}, what=name)


@pytest.mark.parametrize('name', ['sfma_walls5_default', 'sfma_walls5_reverse', 'sfma_walls5_timeout', 'sfma_walls5_random'])
def test_sfma_hand_written_loop_matches_golden(name):
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import SFMA
    from cobel_rl_b200.memory import SFMAMemory
    kind, wname, a, stream = _setup(name)
    env = Gridworld(make_world(wname), rng=stream)
    pol = _policy(a.get('policy', ('eps', 0.1)), stream)

    class _Metric:
        D = load_golden(name)['D']
    mem = SFMAMemory(_Metric(), env.n_states, 4, learning_rate=0.9, rng=stream)
    mem.mode = a.get('mode', 'default')
    ag = SFMA(env.observation_space, env.action_space, pol, mem, None, 0.99, 0.99, rng=stream)
    ag.mask_actions = a.get('mask_actions', False)
    ag.random = a.get('random_replay', False)
    if a.get('valid_mask'):
        ag.action_mask = tb.valid_move_mask(env._succ.cpu().numpy())
    mask = ag.action_mask if ag.mask_actions else None
    tr = _Trace()
    replays, lens = [], []
    S = env.n_states
    for _ in range(a['trials']):
        state, _info = env.reset()
        total, step, last = 0.0, 0, None
        for step in range(a['steps']):
            action = pol.select_action(ag.Q[state], mask[state] if mask is not None else None)
            nxt, reward, end, _trunc, _info = env.step(action)
            exp = {'state': state, 'action': action, 'reward': reward, 'next_state': nxt, 'terminal': 1 - int(end)}
            mem.store(exp)
            ag.update_q(exp)
            tr.s.append(state); tr.a.append(action); tr.s2.append(nxt); tr.r.append(reward)
            state = nxt
            total += reward
            if end:
                last = nxt
                break
        tr.steps.append(step); tr.rew.append(total)
        batch = ag.replay(a['batch'], last)
        idx = [e['action'] * S + e['state'] for e in batch if e['state'] >= 0]
        replays.extend(idx); lens.append(len(idx))
        mem._T.zero_()                   # self.M.T.fill(0), agent/sfma.py:324
    torch.cuda.synchronize()
    got = tr.arrays()
    got.update(Q=ag.Q.cpu().numpy(), Mr=mem.rewards.cpu().numpy(), Ms=mem.states.cpu().numpy(), Mt=mem.terminals.cpu().numpy(),
               C=mem.C.cpu().numpy(), T=mem.T.cpu().numpy(), I=mem.I.cpu().numpy(), replay=np.array(replays, dtype=np.int32),
               replay_len=np.array(lens, dtype=np.int32), draws=int(stream.draw_count[0]))
    assert_equal_records(got, load_golden(name), KEYS['sfma'], what=name)


def test_env_known_answers_through_the_product():
    """unit_tests/test_gridworld.py:12-41 and unit_tests/test_topology.py:60-109 through Gridworld.step /
    Topology.step on the GPU."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld, Topology
    from cobel_rl_b200.misc import gridworld_tools, topology_tools
    stream = cb.BatchStream(None, seed=1, device='cuda:0')
    env = Gridworld(gridworld_tools.make_gridworld(5, 5, [0], np.array([[0, 10.]]), starting_states=[24]), rng=stream)
    assert env.current_state == 24
    states, rewards, ends = [], [], []
    for act in [0, 0, 0, 0, 0, 1, 1, 1, 1]:
        s, r, end, trunc, info = env.step(act)
        states.append(s); rewards.append(r); ends.append(end)
        assert trunc is False and info == {}
    assert states == [23, 22, 21, 20, 20, 15, 10, 5, 0]
    assert rewards == [0] * 8 + [10.] and ends == [False] * 8 + [True]
    with pytest.raises(IndexError):
        env.step(4)
    nodes, starting = topology_tools.t_maze(4, 3, 1)
    topo = Topology(nodes, starting, rng=stream)
    assert topo.current_node == '10'
    visited, rewards, ends = [], [], []
    for act in [1, 1, 1, 1, 1, 2, 2, 2]:
        obs, r, end, trunc, _ = topo.step(act)
        visited.append(topo.current_node); rewards.append(r); ends.append(end)
        assert np.allclose(np.asarray(obs.cpu()), np.asarray(nodes[topo.current_node]['pose'], dtype=np.float64))
    assert visited == ['9', '8', '7', '3', '3', '4', '5', '6']
    assert rewards == [0.] * 7 + [1.] and ends == [False] * 7 + [True]
    # a batch of agents steps together
    st3 = cb.BatchStream(3, seed=2, device='cuda:0')
    env3 = Gridworld(gridworld_tools.make_gridworld(5, 5, [0], np.array([[0, 10.]]), starting_states=[24]), rng=st3)
    s, r, end, _, _ = env3.step(torch.tensor([0, 1, 0]))
    assert s.tolist() == [23, 19, 23] and r.tolist() == [0.0, 0.0, 0.0] and end.tolist() == [False] * 3


def test_pma_memory_gain_need_and_probs_vs_oracle():
    """PMAMemory.compute_gain_batch / compute_need / action_probs_batch (memory/pma.py:333-450) against the oracle."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import PMA
    from cobel_rl_b200.memory import PMAMemory
    from cobel_rl_b200.policy import EpsilonGreedy
    from oracle.philox import LazyStream
    world = make_world('walls5')
    W = tb.compile_gridworld(world)
    stream = cb.BatchStream(2, seed=321, device='cuda:0')
    env = Gridworld(world, rng=stream)
    mem = PMAMemory(world['sas'], EpsilonGreedy(0.1, rng=stream), 0.9, 0.9, 0.9, 0.99, rng=stream)
    ag = PMA(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream), mem)
    ag.mask_actions = True
    ag.action_mask = tb.valid_move_mask(W['succ'])
    ag.train(env, 3, 20, 8)
    torch.cuda.synchronize()
    for i in range(2):
        rng = tb.Draws(LazyStream(321, i), 1)
        st = tb.pma_init(tb.t0_from_succ(W['succ']), 25, 4)
        st['action_mask'] = tb.valid_move_mask(W['succ'])
        tb.pma_train(W, st, rng, 3, 20, 8, gamma_q=0.99, mask_actions=True)
        assert np.array_equal(ag.Q[i].cpu().numpy(), st['Q'])
        gain = mem.compute_gain_batch(ag, ag.action_mask)[i].cpu().numpy()
        assert np.array_equal(gain, tb.pma_gain_batch(st, st['Q'], st['action_mask'], ('eps', 0.1), 0.9, 0.99, 1e-6))
        need = mem.compute_need(torch.tensor([12, 7]))[i].cpu().numpy()
        np.testing.assert_allclose(need, np.tile(st['SR'][[12, 7][i]], 4), rtol=0, atol=1e-12 * np.abs(st['SR']).max())
        stat = mem.compute_need(None)[i].cpu().numpy()
        np.testing.assert_allclose(stat, tb.pma_need(st, None), rtol=0, atol=1e-12)
        probs = mem.action_probs_batch(ag.Q, ag.action_mask)[i].cpu().numpy()
        want = tb._probs_rows(('eps', 0.1), st['Q'], st['action_mask'])
        assert np.array_equal(probs, want / np.sum(want, axis=1).reshape(-1, 1))


def test_sfma_memory_replay_and_random_batch():
    """SFMAMemory.replay (memory level: no Q update) and retrieve_random_batch (memory/sfma.py:238-416) vs the oracle."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import SFMA
    from cobel_rl_b200.memory import SFMAMemory
    from cobel_rl_b200.policy import EpsilonGreedy
    from oracle.philox import LazyStream
    world = make_world('walls5')
    W = tb.compile_gridworld(world)
    D = load_golden('sfma_walls5_default')['D']

    class _Metric:
        pass
    _Metric.D = D
    stream = cb.BatchStream(2, seed=99, device='cuda:0')
    env = Gridworld(world, rng=stream)
    mem = SFMAMemory(_Metric(), 25, 4, rng=stream)
    ag = SFMA(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream), mem, rng=stream)
    ag.train(env, 4, 30, 16)
    q_before = ag.Q.clone()
    batch = mem.replay(12, torch.tensor([12, 3]))
    rnd = mem.retrieve_random_batch(6, np.ones((25, 4), dtype=bool))
    torch.cuda.synchronize()
    assert torch.equal(ag.Q, q_before), 'the memory-level calls must not touch Q'
    for i in range(2):
        rng = tb.Draws(LazyStream(99, i), 1)
        st = tb.sfma_init(25, 4)
        tb.sfma_train(W, st, D, rng, 4, 30, 16)
        idx = tb.sfma_memory_replay(st, D, rng, 12, [12, 3][i])
        got = [int(e['action'][i]) * 25 + int(e['state'][i]) for e in batch if int(e['state'][i]) >= 0]
        assert got == list(idx)
        for e, j in zip(batch, idx):
            assert float(e['reward'][i]) == st['Mr'][j % 25, j // 25] and int(e['next_state'][i]) == st['Ms'][j % 25, j // 25]
        cdf = np.cumsum(np.ones(100) / np.sum(np.ones(100)))
        cdf /= cdf[-1]
        want = [int(cdf.searchsorted(rng.next(), side='right')) for _ in range(6)]
        assert [int(e['action'][i]) * 25 + int(e['state'][i]) for e in rnd] == want
        assert int(stream.draw_count[i]) == rng.k


def test_pma_memory_update_sr_stand_alone():
    """PMAMemory.update_sr() (memory/pma.py:413-415) as its own call: SR = inv(I - gamma T) from the current T by the
    library's Gauss-Jordan kernel, nothing else touched (stream, Q, the experience tables)."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import PMA
    from cobel_rl_b200.memory import PMAMemory
    from cobel_rl_b200.policy import EpsilonGreedy
    world = make_world('walls5')
    stream = cb.BatchStream(7, seed=99, device='cuda:0')
    env = Gridworld(world, rng=stream)
    mem = PMAMemory(world['sas'], EpsilonGreedy(0.1, rng=stream), 0.9, 0.9, np.linspace(0.5, 0.95, 7), 0.9, rng=stream)
    ag = PMA(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream), mem, None, 0.9, 0.99)
    ag.train(env, 2, 30, 8)
    # a T the SR does not belong to: perturb, renormalise the rows
    T = mem.T + 0.01 * torch.rand_like(mem.T)
    mem.T.copy_(T / T.sum(dim=2, keepdim=True))
    before = {k: v.clone() for k, v in dict(Q=ag.Q, Mr=mem.rewards, Ms=mem.states, Mt=mem.terminals, dc=stream.draw_count).items()}
    mem.update_sr()
    torch.cuda.synchronize()
    S = env.n_states
    for i, g in enumerate(np.linspace(0.5, 0.95, 7)):
        want = np.linalg.inv(np.eye(S) - g * mem.T[i].cpu().numpy())
        assert np.abs(mem.SR[i].cpu().numpy() - want).max() < 1e-13 * np.abs(want).max()
    for k, v in dict(Q=ag.Q, Mr=mem.rewards, Ms=mem.states, Mt=mem.terminals, dc=stream.draw_count).items():
        assert torch.equal(v, before[k]), k
