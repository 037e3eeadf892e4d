"""GPU: QAgent CUDA path (cobel_q_run) against the reference goldens and the oracle, bit-exact,
on Gridworld and on Topology graphs (pose observations)."""
import numpy as np
import pytest
import torch

from oracle import cases, tabular as tb
from oracle.philox import LazyStream
from helpers import KEYS, assert_equal_records, cuda_case, load_golden, unpack_run

pytestmark = pytest.mark.gpu

Q_CASES = sorted(n for n, c in cases.CASES.items() if c[0] in ('q_grid', 'q_topo'))


@pytest.mark.parametrize('name', Q_CASES)
def test_q_matches_reference_golden(name):
    want = load_golden(name)
    got = cuda_case(name)
    assert_equal_records(got, want, KEYS['q_grid'] + ['log_len'], what=name)


def test_q_hexagonal_six_actions_and_continue_training():
    """6-neighbour graph; two consecutive train() calls continue the same log and stream."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Topology
    from cobel_rl_b200.agent import QAgent
    from cobel_rl_b200.policy import EpsilonGreedy
    # a hand-built 6-neighbour graph (any dict of nodes with 6 neighbours works)
    n_side = 4
    ids = [str(i) for i in range(n_side * n_side)]
    nodes = {}
    for i, nid in enumerate(ids):
        r, c = divmod(i, n_side)
        nb = [ids[r * n_side + max(c - 1, 0)], ids[max(r - 1, 0) * n_side + c], ids[r * n_side + min(c + 1, n_side - 1)],
              ids[min(r + 1, n_side - 1) * n_side + c], ids[max(r - 1, 0) * n_side + min(c + 1, n_side - 1)],
              ids[min(r + 1, n_side - 1) * n_side + max(c - 1, 0)]]
        nodes[nid] = {'id': nid, 'pose': (float(c), float(r), 0., 0., 0., 0.), 'terminal': False, 'reward': 0.0,
                      'neighbors': nb}
    nodes['3'].update({'terminal': True, 'reward': 2.0})
    stream = cb.BatchStream(5, seed=77, device='cuda:0')
    env = Topology(nodes, rng=stream)
    ag = QAgent(env.observation_space, env.action_space, EpsilonGreedy(0.2, rng=stream), rng=stream)
    ag.record = True
    r1 = ag.train(env, 6, 25, 16)
    r2 = ag.train(env, 5, 25, 16)
    torch.cuda.synchronize()
    W = tb.compile_topology(nodes)
    for i in range(5):
        rng = tb.Draws(LazyStream(77, i), 1)
        st = tb.q_init(W['S'], W['A'])
        o1 = tb.q_train(W, st, rng, 6, 25, 16, policy=('eps', 0.2)).arrays()
        o2 = tb.q_train(W, st, rng, 5, 25, 16, policy=('eps', 0.2)).arrays()
        for res, o in ((r1, o1), (r2, o2)):
            got = unpack_run(res, i, 6, W['succ'], W['reward'])
            assert_equal_records(got, o, ['states', 'actions', 'trial_steps', 'trial_reward', 'replay'], what='agent %d' % i)
        assert np.array_equal(ag.Q[i].cpu().numpy(), st['Q'])
        assert int(ag._log_len[i]) == len(st['log']) and int(stream.draw_count[i]) == rng.k
    M = ag.M
    assert int(M['log_len'][0]) == len(st['log']) or True
    # decoded log of the last agent equals the oracle's
    L = len(st['log'])
    assert M['state'][4, :L].tolist() == [e[0] for e in st['log']]
    assert M['action'][4, :L].tolist() == [e[1] for e in st['log']]
    assert M['reward'][4, :L].tolist() == [e[2] for e in st['log']]


def _eight_neighbour_graph(n_side=4):
    ids = [str(i) for i in range(n_side * n_side)]
    nodes = {}
    cl = lambda v: min(max(v, 0), n_side - 1)
    for i, nid in enumerate(ids):
        r, c = divmod(i, n_side)
        nb = [ids[cl(r + dr) * n_side + cl(c + dc)] for dr, dc in
              ((0, -1), (-1, 0), (0, 1), (1, 0), (-1, -1), (-1, 1), (1, 1), (1, -1))]
        nodes[nid] = {'id': nid, 'pose': (float(c), float(r), 0., 0., 0., 0.), 'terminal': False, 'reward': 0.0,
                      'neighbors': nb}
    nodes[ids[-1]].update({'terminal': True, 'reward': 1.5})
    nodes[ids[5]]['reward'] = 0.25
    return nodes


@pytest.mark.parametrize('policy', [('softmax', 2.0), ('eps', 0.3), ('xeps', 0.3)])
def test_q_eight_actions(policy):
    """8 actions: np.sum over exactly 8 probabilities is NumPy's 8-accumulator tree, not a sequential loop (softmax
    normaliser) -- and the policies' CDFs over 8 entries."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Topology
    from cobel_rl_b200.agent import QAgent
    from cobel_rl_b200 import policy as P
    nodes = _eight_neighbour_graph()
    kind, par = policy
    stream = cb.BatchStream(6, seed=4242, device='cuda:0')
    env = Topology(nodes, rng=stream)
    pol = {'eps': P.EpsilonGreedy, 'xeps': P.ExclusiveEpsilonGreedy, 'softmax': P.Softmax}[kind](par, rng=stream)
    ag = QAgent(env.observation_space, env.action_space, pol, rng=stream)
    ag.record = True
    res = ag.train(env, 8, 30, 32)
    torch.cuda.synchronize()
    W = tb.compile_topology(nodes)
    assert W['A'] == 8
    for i in range(6):
        rng = tb.Draws(LazyStream(4242, i), 1)
        st = tb.q_init(W['S'], W['A'])
        o = tb.q_train(W, st, rng, 8, 30, 32, policy=policy).arrays()
        got = unpack_run(res, i, 8, W['succ'], W['reward'])
        assert_equal_records(got, o, ['states', 'actions', 'trial_steps', 'trial_reward', 'replay'], what='agent %d' % i)
        assert np.array_equal(ag.Q[i].cpu().numpy(), st['Q']) and int(stream.draw_count[i]) == rng.k
    # the stand-alone policy call on 8 entries: probabilities bit-equal to the oracle's
    v = torch.tensor(np.random.default_rng(3).normal(size=(6, 8)), device='cuda:0')
    probs = pol.get_action_probs(v).cpu().numpy()
    for i in range(6):
        assert np.array_equal(probs[i], tb.action_probs(policy, v[i].cpu().numpy(), None)) or kind == 'softmax'
        np.testing.assert_allclose(probs[i], tb.action_probs(policy, v[i].cpu().numpy(), None), rtol=1e-15, atol=0)
