"""GPU: host-side class API behaviour (reference-shaped single-agent mode, callbacks, chunked
launches, interactive env/policy/memory calls) and edge cases (ragged agent counts, replay
batches larger / smaller than a warp, zero trials)."""
import numpy as np
import pytest
import torch

from oracle import tabular as tb
from oracle.philox import LazyStream
from helpers import assert_equal_records, make_world, unpack_run

pytestmark = pytest.mark.gpu


def _dynaq(n, seed=21, world='open5', **kw):
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import DynaQ
    from cobel_rl_b200.policy import EpsilonGreedy
    stream = cb.BatchStream(n, seed=seed, device='cuda:0')
    env = Gridworld(make_world(world), rng=stream)
    ag = DynaQ(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream), **kw)
    return stream, env, ag


def test_single_agent_mode_has_reference_shapes():
    stream, env, ag = _dynaq(None)
    assert isinstance(env.current_state, int)
    s, info = env.reset()
    assert isinstance(s, int) and info == {}
    s2, r, end, trunc, info = env.step(0)
    assert isinstance(s2, int) and isinstance(r, float) and isinstance(end, bool) and trunc is False
    assert env.get_position().shape == (2,)
    ag.train(env, 5, 20, 32)
    assert tuple(ag.Q.shape) == (25, 4) and tuple(ag.M.rewards.shape) == (25, 4)
    assert ag.predict_on_batch([0, 3, 7]).shape == (3, 4)
    assert ag.current_trial == 5


@pytest.mark.parametrize('n,batch', [(1, 32), (5, 7), (9, 64), (3, 0), (4, 33)])
def test_ragged_agent_counts_and_batch_sizes(n, batch):
    stream, env, ag = _dynaq(n, seed=5)
    ag.record = True
    res = ag.train(env, 6, 25, batch)
    torch.cuda.synchronize()
    W = tb.compile_gridworld(make_world('open5'))
    for i in range(n):
        rng = tb.Draws(LazyStream(5, i), 1)
        st = tb.dynaq_init(25, 4)
        rec = tb.dynaq_train(W, st, rng, 6, 25, batch).arrays()
        got = unpack_run(res, i, 4, W['succ'], W['reward'])
        got['Q'] = ag.Q[i].cpu().numpy(); rec['Q'] = st['Q']
        assert_equal_records(got, rec, ['states', 'actions', 'trial_steps', 'replay', 'replay_len', 'Q'], what='agent %d' % i)
        assert int(stream.draw_count[i]) == rng.k


def test_zero_trials_is_a_noop():
    stream, env, ag = _dynaq(3)
    before = stream.draw_count.clone()
    res = ag.train(env, 0, 10, 32)
    assert res['trial_steps'].shape == (3, 0) and torch.equal(stream.draw_count, before)
    assert float(ag.Q.abs().sum()) == 0.0


def test_callbacks_chunked_launches_and_stop():
    seen = []

    def on_trial_end(logs):
        seen.append((logs['trial'], logs['trial_session'], logs['steps'].clone(), logs['trial_reward'].clone()))
        if logs['trial'] == 5:
            logs['agent'].stop = True

    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import DynaQ
    from cobel_rl_b200.policy import EpsilonGreedy
    stream = cb.BatchStream(4, seed=3, device='cuda:0')
    env = Gridworld(make_world('open5'), rng=stream)
    ag = DynaQ(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream),
               custom_callbacks={'on_trial_end': [on_trial_end]})
    ag.trials_per_launch = 3
    res = ag.train(env, 12, 20, 8)
    # stop was raised while processing trial 5 -> the session ends after the launch that contains it (trials 3..5)
    assert [t for t, _, _, _ in seen] == list(range(6)) and ag.current_trial == 6
    assert res['trial_steps'].shape == (4, 6)
    assert torch.equal(torch.stack([s for _, _, s, _ in seen], dim=1), res['trial_steps'])
    # chunked launches give the same result as one launch
    stream2 = cb.BatchStream(4, seed=3, device='cuda:0')
    env2 = Gridworld(make_world('open5'), rng=stream2)
    ag2 = DynaQ(env2.observation_space, env2.action_space, EpsilonGreedy(0.1, rng=stream2))
    res2 = ag2.train(env2, 6, 20, 8)
    assert torch.equal(res2['trial_steps'], res['trial_steps']) and torch.equal(ag2.Q, ag.Q)


def test_interactive_policy_and_memory_calls():
    """Policy.get_action_probs / select_action and DynaQMemory.store / retrieve_batch outside train()."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.memory import DynaQMemory
    from cobel_rl_b200.policy import EpsilonGreedy, ExclusiveEpsilonGreedy, Softmax
    stream = cb.BatchStream(3, seed=8, device='cuda:0')
    v = torch.tensor([[0.0, 1.0, 1.0, -2.0], [0.5, 0.5, 0.5, 0.5], [3.0, 0.0, 1.0, 2.0]], dtype=torch.float64, device='cuda:0')
    mask = torch.tensor([[1, 1, 1, 1], [1, 0, 1, 1], [0, 1, 1, 1]], dtype=torch.bool, device='cuda:0')
    for pol, spec in ((EpsilonGreedy(0.2, rng=stream), ('eps', 0.2)), (ExclusiveEpsilonGreedy(0.2, rng=stream), ('xeps', 0.2)),
                      (Softmax(1.5, rng=stream), ('softmax', 1.5))):
        p = pol.get_action_probs(v, mask).cpu().numpy()
        for i in range(3):
            np.testing.assert_allclose(p[i], tb.action_probs(spec, v[i].cpu().numpy(), mask[i].cpu().numpy()), rtol=1e-15, atol=0)
        k0 = stream.draw_count.clone()
        a = pol.select_action(v, mask)
        assert a.shape == (3,) and torch.equal(stream.draw_count, k0 + 1)
        assert bool(mask[torch.arange(3), a.cpu()].all())
    mem = DynaQMemory(25, 4, 0.9, rng=stream)
    mem.store({'state': torch.tensor([1, 2, 3]), 'action': 1, 'reward': 2.0, 'next_state': torch.tensor([6, 7, 8]), 'terminal': 1})
    got = mem.retrieve(torch.tensor([1, 2, 3]), 1)
    assert got['reward'].tolist() == [1.8, 1.8, 1.8] and got['next_state'].tolist() == [6, 7, 8]
    b = mem.retrieve_batch(5)          # a list of 5 Experience dicts, fields [N] tensors (memory/dyna_q.py:122-157)
    assert len(b) == 5 and b[0]['state'].shape == (3,) and max(int(e['state'].max()) for e in b) < 25


def test_topology_discrete_mode_runs_tabular_agents():
    """Topology(discrete=True) exposes node indices so that DynaQ can run on a graph
    (10x2 linear track of config C4); equals the same structure as a gridworld."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Topology
    from cobel_rl_b200.agent import DynaQ
    from cobel_rl_b200.policy import EpsilonGreedy
    from cobel_rl_b200.misc.topology_tools import linear_track
    nodes, starts = linear_track(10, 2, 1.0, 1.0, 'right')
    stream = cb.BatchStream(2, seed=12, device='cuda:0')
    env = Topology(nodes, starts, rng=stream, discrete=True)
    ag = DynaQ(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream))
    ag.record = True
    res = ag.train(env, 8, 40, 16)
    torch.cuda.synchronize()
    W = tb.compile_topology(nodes, starts)
    for i in range(2):
        rng = tb.Draws(LazyStream(12, i), 1)
        st = tb.dynaq_init(20, 4)
        rec = tb.dynaq_train(W, st, rng, 8, 40, 16).arrays()
        got = unpack_run(res, i, 4, W['succ'], W['reward'])
        got['Q'] = ag.Q[i].cpu().numpy(); rec['Q'] = st['Q']
        assert_equal_records(got, rec, ['states', 'actions', 'trial_steps', 'replay', 'Q'], what='agent %d' % i)


def test_grid_search_over_batched_dynaq_and_monitors(tmp_path):
    """every combination x run of a grid search is one agent of a single launch;
    monitors are fed from the batched per-trial logs."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import DynaQ
    from cobel_rl_b200.policy import EpsilonGreedy
    from cobel_rl_b200.monitor import EscapeLatencyMonitor
    from cobel_rl_b200.optimizer import GridSearchOptimizer
    world = make_world('open5')
    launches = []

    def simulation(task, params):
        n = len(params['_run'])
        stream = cb.BatchStream(n, seed=task['seed'], device='cuda:0')
        env = Gridworld(world, rng=stream)
        mon = EscapeLatencyMonitor(task['trials'], 50, n_agents=n)
        ag = DynaQ(env.observation_space, env.action_space, EpsilonGreedy(params['epsilon'], rng=stream), None,
                   params['lr'], 0.99, custom_callbacks={'on_trial_end': [mon.update]})
        res = ag.train(env, task['trials'], 50, 32)
        launches.append(n)
        assert np.array_equal(mon.get_trace(), res['trial_steps'].cpu().numpy())
        return mon.get_trace()                              # [n, trials] escape latencies

    def loss(sim, data):                                     # mean latency over the last trials vs a target
        return float(sum((np.mean([run[-5:] for run in sim[t]]) - data[t]) ** 2 for t in sim))

    opt = GridSearchOptimizer(str(tmp_path) + '/', {'epsilon': [0.05, 0.3, 0.8], 'lr': [0.2, 0.99]}, nb_runs=16)
    fit = opt.fit(simulation, {'a': {'seed': 1, 'trials': 30}}, {'a': 3.0}, loss)
    assert launches == [6 * 16] and len(fit) == 6
    # strongly exploring agents (epsilon 0.8) are the worst fit to a short-latency target
    assert max(fit, key=fit.get)[0] == 0.8


def test_replay_callbacks_pma_and_sfma():
    """on_replay_end receives the performed / reactivated experiences of every replay call
    (agent/pma.py:112-135, agent/sfma.py:139-187) as padded batched tensors."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import PMA, SFMA
    from cobel_rl_b200.memory import PMAMemory, SFMAMemory
    from cobel_rl_b200.memory.utils.metrics import DR
    from cobel_rl_b200.policy import EpsilonGreedy
    world = make_world('walls5')
    for kind in ('pma', 'sfma'):
        got = []
        stream = cb.BatchStream(3, seed=91, device='cuda:0')
        env = Gridworld(world, rng=stream)
        cbs = {'on_replay_end': [lambda logs: got.append((logs['trial'], logs['replay']))]}
        if kind == 'pma':
            mem = PMAMemory(world['sas'], EpsilonGreedy(0.1, rng=stream), rng=stream)
            ag = PMA(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream), mem, custom_callbacks=cbs)
            per_trial = 2
        else:
            mem = SFMAMemory(DR(5, 5, world['sas'], 0.9, world['invalid_transitions']), 25, 4, rng=stream)
            ag = SFMA(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream), mem, custom_callbacks=cbs,
                      rng=stream)
            ag.start_replay = True
            per_trial = 2
        ag.record = True
        res = ag.train(env, 4, 20, 8)
        assert [t for t, _ in got] == [t for t in range(4) for _ in range(per_trial)]
        flat = torch.cat([r['index'][0][r['index'][0] >= 0] for _, r in got])
        n0 = int(res['n_replay'][0])
        assert torch.equal(flat.int(), res['replay_idx'][0, :n0])
        r = got[-1][1]
        ok = r['index'] >= 0
        assert torch.equal((r['action'] * 25 + r['state'])[ok], r['index'][ok])


def test_non_current_device():
    """Agents bound to a BatchStream on cuda:1 run there even while cuda:0 is torch's current device,
    and give the same results as on cuda:0 (needs 2 GPUs)."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import DynaQ
    from cobel_rl_b200.policy import EpsilonGreedy
    out = []
    for dev in ('cuda:0', 'cuda:1'):
        stream = cb.BatchStream(8, seed=2, device=dev)
        env = Gridworld(make_world('open5'), rng=stream)
        ag = DynaQ(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream))
        res = ag.train(env, 10, 30, 32)
        torch.cuda.synchronize(dev)
        assert ag.Q.device == torch.device(dev)
        out.append((ag.Q.cpu(), res['trial_steps'].cpu()))
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])


def test_unsupported_shapes_and_options_fail_loudly():
    """Nothing silently diverges from the reference: unsupported shapes / options raise."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld, Topology
    from cobel_rl_b200.agent import DynaQ, PMA, SFMA
    from cobel_rl_b200.memory import PMAMemory, SFMAMemory
    from cobel_rl_b200.memory.utils.metrics import Euclidean
    from cobel_rl_b200.policy import EpsilonGreedy
    from cobel_rl_b200.misc.gridworld_tools import make_open_field
    stream = cb.BatchStream(2, seed=1, device='cuda:0')
    # 5 neighbours per node: no kernel instantiation for A = 5
    nodes = {str(i): {'id': str(i), 'pose': (float(i), 0., 0., 0., 0., 0.), 'terminal': i == 3, 'reward': float(i == 3),
                      'neighbors': [str((i + d) % 4) for d in range(5)]} for i in range(4)}
    env5 = Topology(nodes, rng=stream, discrete=True)
    ag = DynaQ(env5.observation_space, env5.action_space, EpsilonGreedy(0.1, rng=stream))
    with pytest.raises(NotImplementedError):
        ag.train(env5, 2, 5, 4)
    # PMA beyond 160 states runs with the banded update_sr only: the dense S x S eliminations are register-tiled for
    # S <= 160 (13x13 = 169 states trains fine on the banded path, see test_pma_gpu.py for 20x20)
    world = make_open_field(13, 13, 0, 1)
    stream = cb.BatchStream(2, seed=1, device='cuda:0')
    env = Gridworld(world, rng=stream)
    mem = PMAMemory(world['sas'], EpsilonGreedy(0.1, rng=stream), rng=stream)
    pma = PMA(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream), mem)
    pma.train(env, 1, 5, 4)
    mem.sr_band_max = -1                  # force the dense path
    with pytest.raises(NotImplementedError):
        pma.train(env, 1, 5, 4)
    # and beyond 512 states / 2048 one-step backups not at all
    world = make_open_field(23, 23, 0, 1)
    stream = cb.BatchStream(2, seed=1, device='cuda:0')
    env = Gridworld(world, rng=stream)
    mem = PMAMemory(world['sas'], EpsilonGreedy(0.1, rng=stream), rng=stream)
    pma = PMA(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream), mem)
    with pytest.raises(NotImplementedError):
        pma.train(env, 1, 5, 4)
    # SFMA: options of the reference that are not implemented
    world = make_world('open5')
    stream = cb.BatchStream(2, seed=1, device='cuda:0')
    env = Gridworld(world, rng=stream)
    smem = SFMAMemory(Euclidean(5, 5), 25, 4, rng=stream)
    sfma = SFMA(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream), smem, rng=stream)
    smem.error_mod = True                   # reads experience['td'] before it exists: broken in the reference too
    with pytest.raises(NotImplementedError):
        sfma.train(env, 1, 5, 4)
    smem.error_mod = False
    sfma.train(env, 2, 10, 4)              # and the supported configuration runs (Euclidean metric)
    # environment / agent mismatch and a CPU device are rejected
    other = Gridworld(make_open_field(4, 4, 0, 1), rng=cb.BatchStream(2, seed=1, device='cuda:0'))
    with pytest.raises(AssertionError):
        sfma.train(other, 1, 5, 4)
    from cobel_rl_b200 import _lib
    cpu_stream = cb.BatchStream(2, device='cpu')
    with pytest.raises(_lib.CobelError):
        cpu_stream.next(1)


def test_maze_template_worlds_run_bit_exact():
    """The T-maze templates (misc/gridworld_tools.py:237-507) through Dyna-Q, PLAIN and recorded kernels, vs the oracle."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import DynaQ
    from cobel_rl_b200.policy import EpsilonGreedy
    from cobel_rl_b200.misc.gridworld_tools import make_double_t_maze, make_t_maze
    for world in (make_t_maze(3, 2), make_double_t_maze(2, 1, 'left-right')):
        W = tb.compile_gridworld(world)
        S = world['states']
        for record in (True, False):
            stream = cb.BatchStream(3, seed=77, device='cuda:0')
            env = Gridworld(world, rng=stream)
            ag = DynaQ(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream))
            ag.record = record
            res = ag.train(env, 8, 30, 32)
            torch.cuda.synchronize()
            for i in range(3):
                rng = tb.Draws(LazyStream(77, i), 1)
                st = tb.dynaq_init(S, 4)
                rec = tb.dynaq_train(W, st, rng, 8, 30, 32).arrays()
                assert np.array_equal(ag.Q[i].cpu().numpy(), st['Q']) and int(stream.draw_count[i]) == rng.k
                assert np.array_equal(res['trial_steps'][i].cpu().numpy(), rec['trial_steps'])


def test_trajectory_monitor_and_occupancy_map_on_device():
    """TrajectoryMonitor (monitor/behavior.py:304-385) rebuilt from the recorded step buffers, and the occupancy map
    (analysis/behavior_spatial.py:9-73) computed on the GPU from it, against the oracle's trajectory."""
    from cobel_rl_b200.monitor import TrajectoryMonitor
    from cobel_rl_b200.analysis import get_occupancy_map
    stream, env, ag = _dynaq(3, seed=31)
    ag.record = True
    res = ag.train(env, 5, 12, 8)
    torch.cuda.synchronize()
    mon = TrajectoryMonitor(5, env)
    pos = mon.from_result(res)                                   # [N, trials, max_steps, 2] on the device, NaN padded
    assert pos.is_cuda and pos.shape[:2] == (3, 5) and pos.shape[3] == 2
    world = make_world('open5')
    W = tb.compile_gridworld(world)
    for i in range(3):
        rng = tb.Draws(LazyStream(31, i), 1)
        st = tb.dynaq_init(25, 4)
        rec = tb.dynaq_train(W, st, rng, 5, 12, 8).arrays()
        lists = mon.from_result(res, agent=i)                    # the reference's list (trials) of lists (steps)
        flat = np.array([p for trial in lists for p in trial])
        assert [len(t) for t in lists] == list(rec['trial_steps'] + 1)
        assert np.array_equal(flat, world['coordinates'][rec['next_states']])
    occ = get_occupancy_map(pos, 5.0, 5.0, 1.0)
    assert occ.is_cuda and occ.shape == (5, 5)
    valid = pos[~torch.isnan(pos).any(dim=-1)].cpu().numpy()
    want = np.histogram2d(valid[:, 0], valid[:, 1], bins=(5, 5), range=[[0, 5], [0, 5]])[0]
    assert np.array_equal(occ.cpu().numpy(), want) and occ.sum().item() == float(res['n_steps'].sum().item())


def test_recorded_run_in_chunked_launches_equals_one_launch():
    """``record`` together with ``trials_per_launch``: the per-launch trace buffers are compacted per agent when the
    launches are merged, so the merged run decodes exactly like a single launch (and ``trial_session`` counts over the
    whole session)."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import DynaQ
    from cobel_rl_b200.monitor import TrajectoryMonitor
    from cobel_rl_b200.policy import EpsilonGreedy
    world = make_world('open5')

    def run(chunk):
        stream = cb.BatchStream(5, seed=77, device='cuda:0')
        env = Gridworld(world, rng=stream)
        seen = []
        ag = DynaQ(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream),
                   custom_callbacks={'on_trial_end': [lambda logs: seen.append((logs['trial'], logs['trial_session']))]})
        ag.record = True
        ag.trials_per_launch = chunk
        res = ag.train(env, 7, 15, 4)
        torch.cuda.synchronize()
        return res, seen
    one, seen_one = run(None)
    many, seen_many = run(3)
    assert seen_one == seen_many == [(t, t) for t in range(7)]
    for k in ('n_steps', 'n_replay', 'trial_steps'):
        assert torch.equal(one[k], many[k]), k
    ns, nr = one['n_steps'], one['n_replay']
    for i in range(5):
        assert torch.equal(one['step_sa'][i, :ns[i]], many['step_sa'][i, :ns[i]])
        assert torch.equal(one['step_next'][i, :ns[i]], many['step_next'][i, :ns[i]])
        assert torch.equal(one['replay_idx'][i, :nr[i]], many['replay_idx'][i, :nr[i]])
        calls = int((one['replay_len'][i] >= 0).sum())
        assert torch.equal(one['replay_len'][i, :calls], many['replay_len'][i, :calls])
        assert bool((many['step_sa'][i, ns[i]:] == -1).all())


def test_zero_trial_sessions_are_no_ops_for_every_agent():
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import QAgent, SR
    from cobel_rl_b200.policy import EpsilonGreedy
    stream = cb.BatchStream(2, seed=1, device='cuda:0')
    env = Gridworld(make_world('open5'), rng=stream)
    k0 = stream.draw_count.clone()
    for ag in (QAgent(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream), rng=stream),
               SR(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream))):
        res = ag.train(env, 0, 10) if isinstance(ag, SR) else ag.train(env, 0, 10, 8)
        assert res['trial_steps'].shape == (2, 0) and int(res['n_steps'].sum()) == 0
    assert torch.equal(stream.draw_count, k0)
