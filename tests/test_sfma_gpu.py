"""GPU: SFMA CUDA path (cobel_sfma_run) against the reference goldens and the oracle.
Integer sequences (trajectory, reactivated experiences, draw counts) and the fp64 tables must be
bit-equal; a draw within 1e-12 of a CDF bin edge would raise COBEL_FLAG_CDF_NEAR_TIE (asserted
absent), because exp() and the CDF prefix sums are the only quantities not bit-identical to NumPy's."""
import numpy as np
import pytest
import torch

from oracle import cases, tabular as tb
from oracle.philox import LazyStream
from helpers import KEYS, assert_equal_records, cuda_case, load_golden, make_world, unpack_run

pytestmark = pytest.mark.gpu

SFMA_CASES = sorted(n for n, c in cases.CASES.items() if c[0] == 'sfma')


@pytest.mark.parametrize('name', SFMA_CASES)
def test_sfma_matches_reference_golden(name):
    want = load_golden(name)
    got = cuda_case(name)
    assert got['flags'] & 2 == 0, 'a CDF draw fell within 1e-12 of a bin edge'
    assert_equal_records(got, want, KEYS['sfma'] + (['modes', 'td'] if 'modes' in want else []), what=name)


@pytest.mark.parametrize('mode', ['default', 'reverse', 'forward', 'blend_forward', 'blend_reverse', 'interpolate', 'sweeping'])
def test_sfma_modes_batch_vs_oracle(mode):
    """All replay modes, 20x20-like larger table (12x12), several agents, start_replay and 2 replays per trial."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import SFMA
    from cobel_rl_b200.memory import SFMAMemory
    from cobel_rl_b200.memory.utils.metrics import DR
    from cobel_rl_b200.policy import EpsilonGreedy
    from cobel_rl_b200.misc.gridworld_tools import make_gridworld
    walls = [(5, 6), (6, 5), (17, 18), (18, 17), (29, 30), (30, 29), (70, 82), (82, 70)]
    world = make_gridworld(12, 12, terminals=[11], rewards=np.array([[11, 5.0]]), starting_states=[132, 77],
                           invalid_transitions=walls)
    metric = DR(12, 12, world['sas'], 0.9, walls)
    n, trials, steps, batch = 4, 5, 70, 24
    stream = cb.BatchStream(n, seed=31337, device='cuda:0')
    env = Gridworld(world, rng=stream)
    mem = SFMAMemory(metric, 144, 4, rng=stream)
    mem.mode = mode
    ag = SFMA(env.observation_space, env.action_space, EpsilonGreedy(0.15, rng=stream), mem, rng=stream)
    ag.mask_actions = True
    W = tb.compile_gridworld(world)
    ag.action_mask = tb.valid_move_mask(W['succ'])
    ag.start_replay = True
    ag.nb_replays = 2
    ag.record = True
    res = ag.train(env, trials, steps, batch)
    torch.cuda.synchronize()
    assert int((res['flags'] & 2).sum()) == 0
    for i in range(n):
        rng = tb.Draws(LazyStream(31337, i), 1)
        st = tb.sfma_init(144, 4)
        st['action_mask'] = tb.valid_move_mask(W['succ'])
        rec = tb.sfma_train(W, st, metric.D, rng, trials, steps, batch, policy=('eps', 0.15), mode=mode,
                            mask_actions=True, start_replay=True, nb_replays=2).arrays()
        got = unpack_run(res, i, 4, W['succ'], W['reward'])
        got.update(Q=ag.Q[i].cpu().numpy(), C=mem.C[i].cpu().numpy(), I=mem.I[i].cpu().numpy(), draws=int(stream.draw_count[i]))
        rec.update(Q=st['Q'], C=st['C'], I=st['I'], draws=rng.k)
        assert_equal_records(got, rec, ['states', 'actions', 'trial_steps', 'replay', 'replay_len', 'Q', 'C', 'I', 'draws'],
                             what='%s agent %d' % (mode, i))


def test_metrics_match_reference_fixture():
    """DR computed by the product's metrics module equals the D stored with the golden (same LAPACK)."""
    from cobel_rl_b200.memory.utils.metrics import DR
    world = make_world('walls5')
    D = DR(5, 5, world['sas'], 0.9, world['invalid_transitions']).D
    np.testing.assert_allclose(D, load_golden('sfma_walls5_default')['D'], rtol=1e-12, atol=1e-15)


def test_sfma_no_replay_and_test_mode():
    """no_replay=True: strengths / recency keep evolving but nothing is replayed (M.T is then never
    zeroed); test(): acts with `policy`, learns nothing."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import SFMA
    from cobel_rl_b200.memory import SFMAMemory
    from cobel_rl_b200.memory.utils.metrics import DR
    from cobel_rl_b200.policy import EpsilonGreedy
    world = make_world('walls5')
    metric = DR(5, 5, world['sas'], 0.9, world['invalid_transitions'])
    stream = cb.BatchStream(3, seed=55, device='cuda:0')
    env = Gridworld(world, rng=stream)
    mem = SFMAMemory(metric, 25, 4, rng=stream)
    ag = SFMA(env.observation_space, env.action_space, EpsilonGreedy(0.2, rng=stream), mem, rng=stream)
    ag.record = True
    res = ag.train(env, 6, 15, 8, no_replay=True)
    rt = ag.test(env, 3, 15)
    torch.cuda.synchronize()
    W = tb.compile_gridworld(world)
    for i in range(3):
        rng = tb.Draws(LazyStream(55, i), 1)
        st = tb.sfma_init(25, 4)
        rec = tb.sfma_train(W, st, metric.D, rng, 6, 15, 8, policy=('eps', 0.2), no_replay=True).arrays()
        rec2 = tb.tabular_test(W, st['Q'], rng, 3, 15, policy=('eps', 0.2)).arrays()
        got = unpack_run(res, i, 4, W['succ'], W['reward'])
        got.update(Q=ag.Q[i].cpu().numpy(), C=mem.C[i].cpu().numpy(), T=mem.T[i].cpu().numpy())
        rec.update(Q=st['Q'], C=st['C'], T=st['T'])
        assert_equal_records(got, rec, ['states', 'actions', 'trial_steps', 'Q', 'C', 'T'], what='agent %d' % i)
        assert float(np.abs(st['T']).sum()) > 0
        got2 = unpack_run(rt, i, 4, W['succ'], W['reward'])
        assert_equal_records(got2, rec2, ['states', 'actions', 'trial_steps'], what='test agent %d' % i)
        assert int(stream.draw_count[i]) == rng.k


@pytest.mark.parametrize('flags,mode', [
    (dict(reward_mod_local=True, reward_modulation=2.5), 'default'),
    (dict(reward_mod=True, reward_modulation=0.5), 'reverse'),
    (dict(state_mod=True), 'default'),
    (dict(C_normalize=True), 'blend_reverse'),
    (dict(D_normalize=True), 'blend_forward'),
    (dict(D_normalize=True, C_normalize=True), 'interpolate'),
    (dict(R_normalize=False, beta=2.0), 'default'),
    (dict(reward_mod_local=True, reward_mod=True, state_mod=True, C_normalize=True, D_normalize=True), 'reverse'),
])
def test_sfma_modulation_flags_vs_oracle(flags, mode):
    """SFMAMemory strength-modulation / normalisation switches (memory/sfma.py:216-236, 283-288, 319-320) on the GPU
    against the oracle (pinned to the reference for the same switches by tests/test_oracle_vs_reference.py)."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import SFMA
    from cobel_rl_b200.memory import SFMAMemory
    from cobel_rl_b200.memory.utils.metrics import DR
    from cobel_rl_b200.policy import EpsilonGreedy
    world = make_world('walls5')
    W = tb.compile_gridworld(world)
    # a world with a negative reward as well: modulated strengths can go negative
    world['rewards'][7] = -0.5
    W['reward'][7] = -0.5
    n, trials, steps, batch = 3, 8, 30, 16
    stream = cb.BatchStream(n, seed=909, device='cuda:0')
    env = Gridworld(world, rng=stream)
    metric = DR(5, 5, world['sas'], 0.9, world['invalid_transitions'])
    mem = SFMAMemory(metric, 25, 4, rng=stream)
    mem.mode = mode
    for k, v in flags.items():
        assert hasattr(mem, k), k
        setattr(mem, k, v)
    ag = SFMA(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream), mem, rng=stream)
    ag.mask_actions = True
    ag.record = True
    res = ag.train(env, trials, steps, batch)
    torch.cuda.synchronize()
    assert int((res['flags'] & 2).sum()) == 0
    store = {k: flags[k] for k in ('reward_mod_local', 'reward_mod', 'state_mod', 'reward_modulation') if k in flags}
    names = {'C_normalize': 'c_normalize', 'D_normalize': 'd_normalize', 'R_normalize': 'r_normalize', 'beta': 'beta'}
    rk = {names[k]: flags[k] for k in names if k in flags}
    D = np.asarray(metric.D)
    for i in range(n):
        rng = tb.Draws(LazyStream(909, i), 1)
        st = tb.sfma_init(25, 4)
        rec = tb.sfma_train(W, st, D, rng, trials, steps, batch, mode=mode, mask_actions=True, replay_kwargs=rk, **store).arrays()
        got = unpack_run(res, i, 4, W['succ'], W['reward'])
        got.update(Q=ag.Q[i].cpu().numpy(), C=mem._C[i].cpu().numpy(), I=mem._I[i].cpu().numpy(), draws=int(stream.draw_count[i]))
        rec.update(Q=st['Q'], C=st['C'], I=st['I'], draws=rng.k)
        assert_equal_records(got, rec, ['states', 'actions', 'trial_steps', 'replay', 'replay_len', 'Q', 'C', 'I', 'draws'],
                             what='agent %d %s' % (i, flags))


def test_sfma_large_state_space_split_path():
    """50x50 (2500 states, 10^4 experiences): the tables do not fit in shared memory -- the split path steps on the tables
    in HBM (one thread per agent) and its replay kernel stages only the list of experienced (s, a); the reference has no
    limit (memory/sfma.py:162-172).  Euclidean similarity (the DR metric would need the dense 200 MB sas)."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import SFMA
    from cobel_rl_b200.memory import SFMAMemory
    from cobel_rl_b200.memory.utils.metrics import Euclidean
    from cobel_rl_b200.misc.gridworld_tools import make_open_field
    from cobel_rl_b200.policy import EpsilonGreedy
    world = make_open_field(50, 50, 0, 1, dense_sas=False)
    metric = Euclidean(50, 50)
    n, trials, steps, batch = 2, 3, 60, 16
    stream = cb.BatchStream(n, seed=808, device='cuda:0')
    env = Gridworld(world, rng=stream)
    mem = SFMAMemory(metric, 2500, 4, rng=stream)
    ag = SFMA(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream), mem, rng=stream)
    ag.record = True
    res = ag.train(env, trials, steps, batch)
    torch.cuda.synchronize()
    assert int(res['flags'].sum()) == 0
    W = {'S': 2500, 'A': 4, 'succ': world['succ'], 'reward': world['rewards'].astype(np.float64),
         'terminal': world['terminals'].astype(np.uint8), 'starts': world['starting_states'].astype(np.int32)}
    for i in range(n):
        rng = tb.Draws(LazyStream(808, i), 1)
        st = tb.sfma_init(2500, 4)
        rec = tb.sfma_train(W, st, metric.D, rng, trials, steps, batch).arrays()
        got = unpack_run(res, i, 4, W['succ'], W['reward'])
        got.update(Q=ag.Q[i].cpu().numpy(), C=mem.C[i].cpu().numpy(), I=mem.I[i].cpu().numpy(), draws=int(stream.draw_count[i]))
        rec.update(Q=st['Q'], C=st['C'], I=st['I'], draws=rng.k)
        assert_equal_records(got, rec, ['states', 'actions', 'trial_steps', 'replay', 'replay_len', 'Q', 'C', 'I', 'draws'],
                             what='agent %d' % i)
