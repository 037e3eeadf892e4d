"""GPU: randomised parity.  Random graphs (2..8 actions, random rewards of either sign, several terminals and starting
nodes, random action masks), random hyper-parameters, every policy kind and replay batches on both sides of a warp:
Dyna-Q and QAgent through the class API must reproduce the oracle bit for bit (trajectories, Q, draw counts)."""
import numpy as np
import pytest
import torch

from oracle import tabular as tb
from oracle.philox import LazyStream

pytestmark = pytest.mark.gpu


def random_graph(rs, S, A):
    terminal = rs.random(S) < 0.15
    terminal[rs.integers(S)] = False                 # at least one node to start from
    reward = np.where(rs.random(S) < 0.3, np.round(rs.normal(0, 2, S), 3), 0.0)
    nodes = {}
    for i in range(S):
        nodes['n%d' % i] = {'id': 'n%d' % i, 'pose': (float(i), 0., 0., 0., 0., 0.), 'terminal': bool(terminal[i]),
                            'reward': float(reward[i]), 'neighbors': ['n%d' % j for j in rs.integers(0, S, A)]}
    free = [k for k, v in nodes.items() if not v['terminal']]
    starts = list(rs.choice(free, size=min(len(free), int(rs.integers(1, 4))), replace=False))
    return nodes, starts


@pytest.mark.parametrize('seed', range(8))
def test_dynaq_on_random_graphs(seed):
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Topology
    from cobel_rl_b200.agent import DynaQ
    from cobel_rl_b200 import policy as P
    rs = np.random.default_rng(1000 + seed)
    S, A = int(rs.integers(3, 40)), int(rs.choice([2, 3, 4, 6, 8]))
    nodes, starts = random_graph(rs, S, A)
    W = tb.compile_topology(nodes, starts)
    kind = ['eps', 'xeps', 'softmax'][seed % 3]
    par = float(np.round(rs.uniform(0.05, 0.5) if kind != 'softmax' else rs.uniform(0.5, 3.0), 3))
    batch = int(rs.choice([0, 5, 32, 32, 40]))
    lr, gamma, mem_lr = (float(np.round(rs.uniform(0.3, 1.0), 3)) for _ in range(3))
    masked = bool(seed % 2)
    mask = rs.random((S, A)) < 0.7
    mask[np.arange(S), rs.integers(0, A, S)] = True   # every node keeps a valid action
    n, trials, steps = 3, 6, 20
    stream = cb.BatchStream(n, seed=500 + seed, device='cuda:0')
    env = Topology(nodes, starts, rng=stream, discrete=True)
    cls = {'eps': P.EpsilonGreedy, 'xeps': P.ExclusiveEpsilonGreedy, 'softmax': P.Softmax}[kind]
    ag = DynaQ(env.observation_space, env.action_space, cls(par, rng=stream), None, lr, gamma)
    ag.M.learning_rate = mem_lr
    if masked:
        ag.mask_actions = True
        ag.action_mask = mask
    ag.record = bool(seed % 4 < 2)                    # both the traced and the PLAIN-eligible launch paths
    res = ag.train(env, trials, steps, batch)
    torch.cuda.synchronize()
    for i in range(n):
        rng = tb.Draws(LazyStream(500 + seed, i), 1)
        st = tb.dynaq_init(S, A)
        if masked:
            st['action_mask'] = mask
        rec = tb.dynaq_train(W, st, rng, trials, steps, batch, policy=(kind, par), lr=lr, gamma=gamma, mem_lr=mem_lr,
                             mask_actions=masked).arrays()
        what = 'seed %d agent %d (S=%d A=%d %s batch=%d)' % (seed, i, S, A, kind, batch)
        assert np.array_equal(res['trial_steps'][i].cpu().numpy(), rec['trial_steps']), what
        assert np.array_equal(ag.Q[i].cpu().numpy(), st['Q']), what
        assert np.array_equal(ag.M.rewards[i].cpu().numpy(), st['Mr']), what
        assert int(stream.draw_count[i]) == rng.k, what
