"""GPU: PMA CUDA path (cobel_pma_run) against the reference goldens and the oracle.

Integer sequences (trajectory, performed updates, draw counts) and Q / M tables must be
bit-equal.  SR and the stationary `need` vector come from an in-kernel Gauss-Jordan / GTH
factorisation instead of LAPACK: tolerance 1e-12 relative to the matrix scale (BASELINE.json
north_star: "within 1e-12 relative"; see helpers.assert_equal_records), and the kernel's top-2
utility gap certificate must stay far above that tolerance."""
import numpy as np
import pytest
import torch

from oracle import cases, tabular as tb
from oracle.philox import LazyStream
from helpers import KEYS, assert_equal_records, cuda_case, load_golden, make_world, unpack_run

pytestmark = pytest.mark.gpu

PMA_CASES = sorted(n for n, c in cases.CASES.items() if c[0] == 'pma')
RTOL = {'SR': 1e-12}


@pytest.mark.parametrize('name', PMA_CASES)
def test_pma_matches_reference_golden(name):
    want = load_golden(name)
    got = cuda_case(name)
    assert got['flags'] == 0
    assert got['min_gap'] > 1e-9, 'utilities too close for a 1e-12-accurate need vector'
    assert_equal_records(got, want, KEYS['pma'], rtol=RTOL, what=name)


@pytest.mark.parametrize('dense', [False, True])
def test_pma_batch_sweep_vs_oracle(dense):
    """Several agents with per-agent hyper-parameters on the 10x10 walled world (config C3 shape),
    including timed-out trials (stationary need).  dense=False: banded update_sr inside the main kernel
    (one launch for all trials); dense=True: the dense Gauss-Jordan / GTH kernel after every trial."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import PMA
    from cobel_rl_b200.memory import PMAMemory
    from cobel_rl_b200.policy import EpsilonGreedy
    world = make_world('walls10')
    n, trials, steps, batch = 4, 2, 40, 24
    gq = [0.99, 0.95, 0.9, 0.99]
    lrq = [0.9, 0.8, 0.9, 0.5]
    stream = cb.BatchStream(n, seed=4242, device='cuda:0')
    env = Gridworld(world, rng=stream)
    mem = PMAMemory(world['sas'], EpsilonGreedy(0.1, rng=stream), 0.9, lrq, 0.9, gq, rng=stream)
    ag = PMA(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream), mem)
    ag.mask_actions = True
    W = tb.compile_gridworld(world)
    ag.action_mask = tb.valid_move_mask(W['succ'])
    ag.record = True
    if dense:
        mem.sr_band_max = -1
    assert mem.sr_band(env.transition_band)[0] == (-1 if dense else 10)
    res = ag.train(env, trials, steps, batch)
    torch.cuda.synchronize()
    assert int(res['flags'].sum()) == 0
    for i in range(n):
        rng = tb.Draws(LazyStream(4242, i), 1)
        st = tb.pma_init(tb.t0_from_succ(W['succ']), 100, 4)
        st['action_mask'] = tb.valid_move_mask(W['succ'])
        rec = tb.pma_train(W, st, rng, trials, steps, batch, lr_q=lrq[i], gamma_q=gq[i], mask_actions=True).arrays()
        got = unpack_run(res, i, 4, W['succ'], W['reward'])
        got.update(Q=ag.Q[i].cpu().numpy(), T=mem.T[i].cpu().numpy(), SR=mem.SR[i].cpu().numpy(), draws=int(stream.draw_count[i]))
        rec.update(Q=st['Q'], T=st['T'], SR=st['SR'], draws=rng.k)
        assert_equal_records(got, rec, ['states', 'actions', 'trial_steps', 'replay', 'replay_len', 'Q', 'T', 'SR', 'draws'],
                             rtol=RTOL, what='agent %d' % i)
    assert float(mem.min_gap.min()) > 1e-9


def test_pma_no_replay_test_mode_and_softmax_memory_policy():
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import PMA
    from cobel_rl_b200.memory import PMAMemory
    from cobel_rl_b200.policy import EpsilonGreedy, ExclusiveEpsilonGreedy
    world = make_world('walls5')
    W = tb.compile_gridworld(world)
    # (a) no_replay + test()
    stream = cb.BatchStream(2, seed=66, device='cuda:0')
    env = Gridworld(world, rng=stream)
    mem = PMAMemory(world['sas'], EpsilonGreedy(0.1, rng=stream), rng=stream)
    ag = PMA(env.observation_space, env.action_space, EpsilonGreedy(0.3, rng=stream), mem)
    ag.record = True
    res = ag.train(env, 5, 12, 16, no_replay=True)
    rt = ag.test(env, 2, 12)
    torch.cuda.synchronize()
    for i in range(2):
        rng = tb.Draws(LazyStream(66, i), 1)
        st = tb.pma_init(tb.t0_from_succ(W['succ']), 25, 4)
        rec = tb.pma_train(W, st, rng, 5, 12, 16, policy=('eps', 0.3), no_replay=True).arrays()
        rec2 = tb.tabular_test(W, st['Q'], rng, 2, 12, policy=('eps', 0.3)).arrays()
        got = unpack_run(res, i, 4, W['succ'], W['reward'])
        got.update(Q=ag.Q[i].cpu().numpy(), T=mem.T[i].cpu().numpy(), SR=mem.SR[i].cpu().numpy())
        rec.update(Q=st['Q'], T=st['T'], SR=st['SR'])
        assert_equal_records(got, rec, ['states', 'actions', 'trial_steps', 'Q', 'T', 'SR'], what='agent %d' % i)
        assert_equal_records(unpack_run(rt, i, 4, W['succ'], W['reward']), rec2, ['states', 'actions', 'trial_steps'])
        assert int(stream.draw_count[i]) == rng.k
    # (b) exclusive epsilon-greedy as memory policy, min_gain_mode != 'original'
    stream = cb.BatchStream(2, seed=67, device='cuda:0')
    env = Gridworld(world, rng=stream)
    mem = PMAMemory(world['sas'], ExclusiveEpsilonGreedy(0.2, rng=stream), gamma_q=0.95, rng=stream)
    mem.min_gain_mode = ''
    ag = PMA(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream), mem)
    ag.record = True
    res = ag.train(env, 4, 20, 12)
    torch.cuda.synchronize()
    for i in range(2):
        rng = tb.Draws(LazyStream(67, i), 1)
        st = tb.pma_init(tb.t0_from_succ(W['succ']), 25, 4)
        rec = tb.pma_train(W, st, rng, 4, 20, 12, mem_policy=('xeps', 0.2), gamma_q=0.95,
                           replay_kwargs={'original': False}).arrays()
        got = unpack_run(res, i, 4, W['succ'], W['reward'])
        got['Q'] = ag.Q[i].cpu().numpy(); rec['Q'] = st['Q']
        assert_equal_records(got, rec, ['states', 'actions', 'trial_steps', 'replay', 'replay_len', 'Q'], what='xeps agent %d' % i)


def test_pma_band_violation_is_reported():
    """A caller that promises a narrower band than T has gets COBEL_FLAG_BAND_VIOLATION, not a wrong answer."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200 import _lib
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import PMA
    from cobel_rl_b200.memory import PMAMemory
    from cobel_rl_b200.policy import EpsilonGreedy
    world = make_world('walls10')
    stream = cb.BatchStream(2, seed=5, device='cuda:0')
    env = Gridworld(world, rng=stream)
    mem = PMAMemory(world['sas'], EpsilonGreedy(0.1, rng=stream), rng=stream)
    ag = PMA(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream), mem)
    assert env.transition_band == 10
    real = mem.sr_band
    mem.sr_band = lambda wb: (3, real(wb)[1], False)     # lie: the world's band is 10 (and T is not vouched for)
    with pytest.raises(_lib.CobelError, match='band'):
        ag.train(env, 1, 20, 4)
    # a T written through the public view is re-measured and checked by the library again
    mem.sr_band = real
    assert mem.sr_band(10)[2] is True
    mem.T[0, 0, 99] = 0.5                                 # far outside the band of 10: the dense update_sr takes over
    assert mem.sr_band(10)[2] is False and mem.sr_band(10)[0] == -1


@pytest.mark.parametrize('opts', [dict(equal_need=True), dict(equal_gain=True), dict(ignore_barriers=False),
                                  dict(equal_need=True, equal_gain=True, ignore_barriers=False), dict(allow_loops=True),
                                  dict(allow_loops=True, equal_need=True)])
def test_pma_replay_switches_vs_oracle(opts):
    """PMAMemory.equal_need / equal_gain / ignore_barriers (memory/pma.py:238-249) on the GPU vs the oracle (which
    tests/test_oracle_vs_reference.py pins against the reference for the same switches)."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import PMA
    from cobel_rl_b200.memory import PMAMemory
    from cobel_rl_b200.policy import EpsilonGreedy
    world = make_world('walls5')
    W = tb.compile_gridworld(world)
    n, trials, steps, batch = 3, 3, 15, 10
    stream = cb.BatchStream(n, seed=1212, device='cuda:0')
    env = Gridworld(world, rng=stream)
    mem = PMAMemory(world['sas'], EpsilonGreedy(0.1, rng=stream), 0.9, 0.9, 0.9, 0.99, rng=stream)
    for k, v in opts.items():
        setattr(mem, k, v)
    ag = PMA(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream), mem)
    ag.mask_actions = True
    ag.record = True
    res = ag.train(env, trials, steps, batch)
    torch.cuda.synchronize()
    assert int(res['flags'].sum()) == 0
    for i in range(n):
        rng = tb.Draws(LazyStream(1212, i), 1)
        st = tb.pma_init(tb.t0_from_succ(W['succ']), 25, 4)
        rec = tb.pma_train(W, st, rng, trials, steps, batch, gamma_q=0.99, mask_actions=True, replay_kwargs=dict(opts)).arrays()
        got = unpack_run(res, i, 4, W['succ'], W['reward'])
        got.update(Q=ag.Q[i].cpu().numpy(), T=mem.T[i].cpu().numpy(), SR=mem.SR[i].cpu().numpy(), draws=int(stream.draw_count[i]))
        rec.update(Q=st['Q'], T=st['T'], SR=st['SR'], draws=rng.k)
        assert_equal_records(got, rec, ['states', 'actions', 'trial_steps', 'replay', 'replay_len', 'Q', 'T', 'SR', 'draws'],
                             rtol=RTOL, what='agent %d %s' % (i, opts))


@pytest.mark.parametrize('shape', [(6, 7), (4, 9), (7, 5)])
def test_pma_on_other_grid_shapes_vs_oracle(shape):
    """PMA on walled grids of other shapes (bandwidths 7, 9 and 5; state counts 42, 36, 35): the banded update_sr
    against the oracle's dense LAPACK inverse.  (Worlds with unreachable cells, e.g. the maze templates, have no unique
    stationary distribution: a timed-out trial there raises COBEL_FLAG_SINGULAR, and the reference's `eig` need is
    not defined either.)"""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import PMA
    from cobel_rl_b200.memory import PMAMemory
    from cobel_rl_b200.misc import gridworld_tools as mg
    from cobel_rl_b200.policy import EpsilonGreedy
    h, w = shape
    goal = w - 1
    walls = [(1, 2), (2, 1), (w + 1, w + 2), (w + 2, w + 1), (2 * w + 3, 3 * w + 3), (3 * w + 3, 2 * w + 3)]
    world = mg.make_gridworld(h, w, terminals=[goal], rewards=np.array([[goal, 5.0]]), goals=[goal],
                              starting_states=[h * w - w], invalid_transitions=walls)
    W = tb.compile_gridworld(world)
    S = world['states']
    n, trials, steps, batch = 3, 3, 30, 12
    stream = cb.BatchStream(n, seed=77, device='cuda:0')
    env = Gridworld(world, rng=stream)
    mem = PMAMemory(world['sas'], EpsilonGreedy(0.1, rng=stream), 0.9, 0.9, 0.9, 0.99, rng=stream)
    ag = PMA(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream), mem)
    ag.record = True
    assert mem.sr_band(env.transition_band)[0] == w
    res = ag.train(env, trials, steps, batch)
    torch.cuda.synchronize()
    assert int(res['flags'].sum()) == 0
    for i in range(n):
        rng = tb.Draws(LazyStream(77, i), 1)
        st = tb.pma_init(tb.t0_from_succ(W['succ']), S, 4)
        rec = tb.pma_train(W, st, rng, trials, steps, batch, gamma_q=0.99).arrays()
        got = unpack_run(res, i, 4, W['succ'], W['reward'])
        got.update(Q=ag.Q[i].cpu().numpy(), T=mem.T[i].cpu().numpy(), SR=mem.SR[i].cpu().numpy(), draws=int(stream.draw_count[i]))
        rec.update(Q=st['Q'], T=st['T'], SR=st['SR'], draws=rng.k)
        assert_equal_records(got, rec, ['states', 'actions', 'trial_steps', 'replay', 'replay_len', 'Q', 'T', 'SR', 'draws'],
                             rtol=RTOL, what='%s agent %d' % (shape, i))


def test_pma_20x20_large_state_space():
    """PMA on a 20x20 walled gridworld (400 states, 1600 one-step backups, band 20): pma_main_kernel<A, false, BIG> with
    two chunk maxima per lane; the reference has no limit on the state space (memory/pma.py:139-141).  Integer
    sequences and Q bit-equal to the oracle, SR against its dense LAPACK inverse."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import PMA
    from cobel_rl_b200.memory import PMAMemory
    from cobel_rl_b200.misc import gridworld_tools as mg
    from cobel_rl_b200.policy import EpsilonGreedy
    walls = [(5, 6), (6, 5), (25, 26), (26, 25), (45, 46), (46, 45), (208, 228), (228, 208), (209, 229), (229, 209)]
    world = mg.make_gridworld(20, 20, terminals=[19], rewards=np.array([[19, 10.0]]), goals=[19], starting_states=[210],
                              invalid_transitions=walls)
    W = tb.compile_gridworld(world)
    n, trials, steps, batch = 2, 2, 70, 16
    stream = cb.BatchStream(n, seed=2020, device='cuda:0')
    env = Gridworld(world, rng=stream)
    mem = PMAMemory(world['sas'], EpsilonGreedy(0.1, rng=stream), 0.9, 0.9, 0.9, 0.99, rng=stream)
    ag = PMA(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream), mem)
    ag.mask_actions = True
    ag.action_mask = tb.valid_move_mask(W['succ'])
    ag.record = True
    assert mem.sr_band(env.transition_band)[0] == 20
    res = ag.train(env, trials, steps, batch)
    torch.cuda.synchronize()
    assert int(res['flags'].sum()) == 0
    for i in range(n):
        rng = tb.Draws(LazyStream(2020, i), 1)
        st = tb.pma_init(tb.t0_from_succ(W['succ']), 400, 4)
        st['action_mask'] = tb.valid_move_mask(W['succ'])
        rec = tb.pma_train(W, st, rng, trials, steps, batch, gamma_q=0.99, mask_actions=True).arrays()
        got = unpack_run(res, i, 4, W['succ'], W['reward'])
        got.update(Q=ag.Q[i].cpu().numpy(), T=mem.T[i].cpu().numpy(), SR=mem.SR[i].cpu().numpy(), draws=int(stream.draw_count[i]))
        rec.update(Q=st['Q'], T=st['T'], SR=st['SR'], draws=rng.k)
        assert_equal_records(got, rec, ['states', 'actions', 'trial_steps', 'replay', 'replay_len', 'Q', 'T', 'SR', 'draws'],
                             rtol=RTOL, what='agent %d' % i)
