"""GPU: the BASELINE.json configurations at (or near) full size.  The oracle can only afford a few
agents at these sizes, so every run is checked through size-independent identities on ALL agents
(step / replay / draw bookkeeping) and bit-exactly against the oracle on a sample of agents."""
import numpy as np
import pytest
import torch

from oracle import tabular as tb
from oracle.philox import LazyStream
from helpers import make_world

pytestmark = pytest.mark.gpu
SEED = 0x5EED


def test_c2_dynaq_4096_agents_full_horizon():
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import DynaQ
    from cobel_rl_b200.policy import EpsilonGreedy
    n, trials, steps, batch = 4096, 500, 50, 32
    world = make_world('open5')
    stream = cb.BatchStream(n, seed=SEED, device='cuda:0')
    env = Gridworld(world, rng=stream)
    ag = DynaQ(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream))
    res = ag.train(env, trials, steps, batch)
    torch.cuda.synchronize()
    ts = res['trial_steps'].long()
    assert torch.equal(res['n_steps'], (ts + 1).sum(dim=1))
    assert torch.equal(res['n_replay'], res['n_steps'] * batch)
    # draws: 1 (env constructor) + 1 per reset + 1 per step + batch per step
    assert torch.equal(stream.draw_count, 1 + trials + res['n_steps'] * (1 + batch))
    assert int(ts.min()) >= 0 and int(ts.max()) <= steps - 1
    # learning happened everywhere: late trials are short
    assert float(ts[:, -50:].double().mean()) < 8.0 < float(ts[:, :5].double().mean())
    W = tb.compile_gridworld(world)
    for i in (0, 1, 2047, 4095):
        rng = tb.Draws(LazyStream(SEED, i), 1)
        st = tb.dynaq_init(25, 4)
        rec = tb.dynaq_train(W, st, rng, trials, steps, batch).arrays()
        assert np.array_equal(rec['trial_steps'], res['trial_steps'][i].cpu().numpy()), 'agent %d' % i
        assert np.array_equal(st['Q'], ag.Q[i].cpu().numpy()) and np.array_equal(st['Mr'], ag.M.rewards[i].cpu().numpy())
        assert rng.k == int(stream.draw_count[i])


def test_c3_pma_16384_agents():
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import PMA
    from cobel_rl_b200.memory import PMAMemory
    from cobel_rl_b200.policy import EpsilonGreedy
    n, trials, steps, batch = 16384, 2, 100, 32
    world = make_world('walls10')
    stream = cb.BatchStream(n, seed=SEED, device='cuda:0')
    env = Gridworld(world, rng=stream)
    mem = PMAMemory(world['sas'], EpsilonGreedy(0.1, rng=stream), 0.9, 0.9, 0.9, 0.99, rng=stream)
    ag = PMA(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream), mem, None, 0.9, 0.99)
    ag.mask_actions = True
    res = ag.train(env, trials, steps, batch)
    torch.cuda.synchronize()
    assert int(res['flags'].sum()) == 0
    assert torch.equal(res['n_replay'], torch.full_like(res['n_replay'], 2 * batch * trials))
    assert torch.equal(res['n_steps'], (res['trial_steps'].long() + 1).sum(dim=1))
    # certificate: the closest call between two distinct utilities over all 2M selections (5.9e-10 here) is
    # still orders of magnitude above the accuracy of the need vector (~1e-15 for the entries that can win)
    assert float(mem.min_gap.min()) > 1e-11
    # T stays row-stochastic, SR = inv(I - 0.9 T): check the defining identity on a sample
    T, SR = mem.T[:64], mem.SR[:64]
    assert float((T.sum(dim=2) - 1).abs().max()) < 1e-12
    eye = torch.eye(100, dtype=torch.float64, device=T.device)
    assert float((torch.matmul(eye - 0.9 * T, SR) - eye).abs().max()) < 1e-12
    W = tb.compile_gridworld(world)
    for i in (0, 16383):
        rng = tb.Draws(LazyStream(SEED, i), 1)
        st = tb.pma_init(tb.t0_from_succ(W['succ']), 100, 4)
        rec = tb.pma_train(W, st, rng, trials, steps, batch, gamma_q=0.99, mask_actions=True).arrays()
        assert np.array_equal(rec['trial_steps'], res['trial_steps'][i].cpu().numpy())
        assert np.array_equal(st['Q'], ag.Q[i].cpu().numpy()), 'agent %d' % i
        assert rng.k == int(stream.draw_count[i])


def test_c4_sfma_20x20_and_track():
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import SFMA
    from cobel_rl_b200.memory import SFMAMemory
    from cobel_rl_b200.memory.utils.metrics import DR
    from cobel_rl_b200.policy import EpsilonGreedy
    from cobel_rl_b200.misc.gridworld_tools import make_open_field
    for world, hw, mode, n in ((make_open_field(20, 20, 0, 1), (20, 20), 'default', 8192),
                               (make_world('track10x2'), (2, 10), 'reverse', 8192)):
        S = hw[0] * hw[1]
        metric = DR(hw[1], hw[0], world['sas'], 0.9, world['invalid_transitions'])
        stream = cb.BatchStream(n, seed=SEED, device='cuda:0')
        env = Gridworld(world, rng=stream)
        mem = SFMAMemory(metric, S, 4, rng=stream)
        mem.mode = mode
        ag = SFMA(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream), mem, rng=stream)
        ag.mask_actions = True
        trials, steps, batch = 3, 150, 32
        res = ag.train(env, trials, steps, batch)
        torch.cuda.synchronize()
        assert int((res['flags'] & ~2).sum()) == 0
        assert int((res['flags'] & 2).sum()) == 0, 'a CDF draw fell within 1e-12 of a bin edge'
        assert torch.equal(res['n_steps'], (res['trial_steps'].long() + 1).sum(dim=1))
        assert int(res['n_replay'].max()) <= trials * batch
        # strengths count the stored experiences: sum(C) == number of steps (decay_strength = 1)
        assert torch.equal(mem.C.sum(dim=1), res['n_steps'].double())
        assert float(mem.T.abs().max()) == 0.0 and float(mem.I.max()) <= 1.0
        W = tb.compile_gridworld(world)
        for i in (0, n - 1):
            rng = tb.Draws(LazyStream(SEED, i), 1)
            st = tb.sfma_init(S, 4)
            rec = tb.sfma_train(W, st, metric.D, rng, trials, steps, batch, mode=mode, mask_actions=True).arrays()
            assert np.array_equal(rec['trial_steps'], res['trial_steps'][i].cpu().numpy())
            assert np.array_equal(st['Q'], ag.Q[i].cpu().numpy()) and np.array_equal(st['C'], mem.C[i].cpu().numpy())
            assert rng.k == int(stream.draw_count[i])


def test_c5_sr_100x100_compact_262144_agents():
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import SR
    from cobel_rl_b200.policy import EpsilonGreedy
    from cobel_rl_b200.misc.gridworld_tools import make_open_field
    n, trials, steps = 262144, 2, 48
    world = make_open_field(100, 100, 0, 1, dense_sas=False)
    stream = cb.BatchStream(n, seed=SEED, device='cuda:0')
    env = Gridworld(world, rng=stream)
    ag = SR(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream), None, 0.1, 0.99, compact=True,
            max_visited=100)
    res = ag.train(env, trials, steps)
    torch.cuda.synchronize()
    assert int(res['flags'].sum()) == 0
    assert torch.equal(res['n_steps'], (res['trial_steps'].long() + 1).sum(dim=1))
    assert torch.equal(stream.draw_count, 1 + trials + res['n_steps'])
    v = ag.n_visited.long()
    assert int(v.min()) >= 2 and bool((v <= res['n_steps'] + trials).all())
    # every SR row is a discounted occupancy: non-negative, diagonal >= its initial weight decay
    blk = ag.SR_compact[:256]
    assert float(blk.min()) >= 0.0
    W = {'S': 10000, 'A': 4, 'succ': world['succ'], 'reward': world['rewards'].astype(np.float64),
         'terminal': world['terminals'].astype(np.uint8), 'starts': world['starting_states'].astype(np.int32)}
    for i in (0, n - 1):
        rng = tb.Draws(LazyStream(SEED, i), 1)
        st = tb.sr_init(10000, 4)
        rec = tb.sr_train(W, st, rng, trials, steps).arrays()
        assert np.array_equal(rec['trial_steps'], res['trial_steps'][i].cpu().numpy())
        assert np.array_equal(ag.dense_sr(i).cpu().numpy(), st['SR']), 'agent %d' % i
