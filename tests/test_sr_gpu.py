"""GPU: SR-agent CUDA path (cobel_sr_run) against the reference goldens and the oracle, bit-exact
(the row dots follow NumPy's pairwise summation tree)."""
import numpy as np
import pytest
import torch

from oracle import cases, tabular as tb
from oracle.philox import LazyStream
from helpers import KEYS, assert_equal_records, cuda_case, load_golden, unpack_run

pytestmark = pytest.mark.gpu

SR_CASES = sorted(n for n, c in cases.CASES.items() if c[0] == 'sr')


@pytest.mark.parametrize('name', SR_CASES)
def test_sr_matches_reference_golden(name):
    want = load_golden(name)
    got = cuda_case(name)
    keys = [k for k in KEYS['sr'] if k not in ('replay', 'replay_len')]
    assert_equal_records(got, want, keys, what=name)


@pytest.mark.parametrize('hw', [(3, 2), (9, 15), (20, 20), (23, 31)])
def test_sr_sizes_vs_oracle(hw):
    """State counts below 8, in 8..128, and above 128 (two- and three-level pairwise trees)."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import SR
    from cobel_rl_b200.policy import EpsilonGreedy
    from cobel_rl_b200.misc.gridworld_tools import make_gridworld
    h, w = hw
    S = h * w
    world = make_gridworld(h, w, terminals=[0, S - 1], rewards=np.array([[0, 1.0], [S - 1, -0.5], [S // 2, 0.25]]))
    stream = cb.BatchStream(3, seed=2024, device='cuda:0')
    env = Gridworld(world, rng=stream)
    ag = SR(env.observation_space, env.action_space, EpsilonGreedy(0.3, rng=stream), None, [0.1, 0.5, 0.9], 0.95)
    ag.record = True
    trials, steps = 6, 60
    res = ag.train(env, trials, steps)
    rt = ag.test(env, 2, 30)
    torch.cuda.synchronize()
    W = tb.compile_gridworld(world)
    for i, lr in enumerate([0.1, 0.5, 0.9]):
        rng = tb.Draws(LazyStream(2024, i), 1)
        st = tb.sr_init(S, 4)
        rec = tb.sr_train(W, st, rng, trials, steps, policy=('eps', 0.3), lr=lr, gamma=0.95).arrays()
        rec2 = tb.sr_train(W, st, rng, 2, 30, policy=('eps', 0.3), learn=False).arrays()
        got = unpack_run(res, i, 4, W['succ'], W['reward'])
        got.update(SR=ag.SR[i].cpu().numpy(), rew=ag.rewards[i].cpu().numpy(), model=ag.model[i].cpu().numpy())
        rec.update(SR=st['SR'], rew=st['rew'], model=st['model'])
        assert_equal_records(got, rec, ['states', 'actions', 'trial_steps', 'trial_reward', 'SR', 'rew', 'model'],
                             what='%dx%d agent %d' % (h, w, i))
        got2 = unpack_run(rt, i, 4, W['succ'], W['reward'])
        assert_equal_records(got2, rec2, ['states', 'actions', 'trial_steps'], what='test() agent %d' % i)
        assert int(stream.draw_count[i]) == rng.k


@pytest.mark.parametrize('hw,rewards', [((50, 50), 'goal'), ((40, 40), 'many'), ((100, 100), 'goal')])
def test_sr_compact_matches_dense_oracle(hw, rewards):
    """Visited-set compaction (config C5 shape): trajectories, draw counts and the implied dense SR /
    rewards / model are bit-equal to the dense oracle; 'many' puts non-zero rewards on a quarter of the
    states so that several columns enter the sparse pairwise sum."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import SR
    from cobel_rl_b200.policy import EpsilonGreedy
    from cobel_rl_b200.misc.gridworld_tools import make_gridworld
    h, w = hw
    S = h * w
    if rewards == 'goal':
        rw = np.array([[0, 1.0]])
        starts = [1, w, w + 1, 2 * w + 2]                      # near the goal so that it is found
    else:
        ids = np.arange(S)[::4]
        rw = np.stack([ids, 0.1 + 0.01 * (ids % 17)], axis=1)
        starts = [S // 2 + w // 2]
    world = make_gridworld(h, w, terminals=[0], rewards=rw, starting_states=starts, dense_sas=False)
    n, trials, steps = 3, 5, 40
    stream = cb.BatchStream(n, seed=777, device='cuda:0')
    env = Gridworld(world, rng=stream)
    ag = SR(env.observation_space, env.action_space, EpsilonGreedy(0.2, rng=stream), None, 0.2, 0.95,
            compact=True, max_visited=256)
    ag.record = True
    res = ag.train(env, trials, steps)
    torch.cuda.synchronize()
    W = {'S': S, 'A': 4, 'succ': world['succ'], 'reward': world['rewards'].astype(np.float64),
         'terminal': world['terminals'].astype(np.uint8), 'starts': world['starting_states'].astype(np.int32)}
    for i in range(n):
        rng = tb.Draws(LazyStream(777, i), 1)
        st = tb.sr_init(S, 4)
        rec = tb.sr_train(W, st, rng, trials, steps, policy=('eps', 0.2), lr=0.2, gamma=0.95).arrays()
        got = unpack_run(res, i, 4, W['succ'], W['reward'])
        assert_equal_records(got, rec, ['states', 'actions', 'trial_steps', 'trial_reward'], what='agent %d' % i)
        assert int(stream.draw_count[i]) == rng.k
        assert np.array_equal(ag.dense_sr(i).cpu().numpy(), st['SR']), 'SR of agent %d' % i
        assert np.array_equal(ag.dense_rewards(i).cpu().numpy(), st['rew'])
        v = int(ag.n_visited[i])
        vis = ag.visited[i, :v].cpu().numpy()
        assert len(set(vis.tolist())) == v and set(vis.tolist()) == set(rec['states'].tolist()) | set(rec['next_states'].tolist())
        model_dense = np.tile(np.arange(S).reshape(S, 1), 4)
        model_dense[vis] = vis[ag.model[i, :v].cpu().numpy()]
        assert np.array_equal(model_dense, st['model'])


def test_sr_compact_overflow_is_reported():
    import cobel_rl_b200 as cb
    from cobel_rl_b200 import _lib
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import SR
    from cobel_rl_b200.policy import EpsilonGreedy
    from cobel_rl_b200.misc.gridworld_tools import make_open_field
    stream = cb.BatchStream(2, seed=1, device='cuda:0')
    env = Gridworld(make_open_field(30, 30, 0, 1, dense_sas=False), rng=stream)
    ag = SR(env.observation_space, env.action_space, EpsilonGreedy(1.0, rng=stream), compact=True, max_visited=8)
    with pytest.raises(_lib.CobelError):
        ag.train(env, 3, 200)
