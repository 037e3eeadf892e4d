"""GPU: Dyna-Q CUDA path (cobel_dynaq_run through the Python class API) against the golden
vectors generated from the reference and against the oracle, bit-exact."""
import numpy as np
import pytest
import torch

from oracle import cases, tabular as tb
from oracle.philox import LazyStream, uniforms
from helpers import KEYS, assert_equal_records, cuda_case, load_golden, make_world, unpack_run

pytestmark = pytest.mark.gpu

DYNAQ_CASES = sorted(n for n, c in cases.CASES.items() if c[0] == 'dynaq')


def test_stream_contract():
    """Device Philox stream == host definition (oracle/philox.py), random access and sequential."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200 import _lib
    out = torch.empty((5, 33), dtype=torch.float64, device='cuda:0')
    _lib.check(_lib.lib().cobel_draw_uniforms(cases.SEED, 7, 5, 3, 33, out.data_ptr(), None))
    torch.cuda.synchronize()
    for i in range(5):
        assert np.array_equal(out[i].cpu().numpy(), uniforms(cases.SEED, 7 + i, 33, first=3))
    st = cb.BatchStream(4, seed=123456789012345, device='cuda:0', agent_id_base=2**33)
    a = st.next(3).cpu().numpy()
    b = st.next(4).cpu().numpy()
    for i in range(4):
        assert np.array_equal(np.concatenate([a[i], b[i]]), uniforms(123456789012345, 2**33 + i, 7))
    assert st.draw_count.tolist() == [7] * 4


@pytest.mark.parametrize('name', DYNAQ_CASES)
def test_dynaq_matches_reference_golden(name):
    want = load_golden(name)
    got = cuda_case(name)
    keys = KEYS['dynaq'] + ['test_states', 'test_actions', 'test_trial_steps', 'test_trial_reward', 'draws_after_test']
    assert_equal_records(got, want, keys, what=name)


def test_dynaq_sweep_batch_matches_oracle():
    """256 agents with per-agent hyper-parameters (a sweep); a sample is replayed by the oracle."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import DynaQ
    from cobel_rl_b200.memory import DynaQMemory
    from cobel_rl_b200.policy import EpsilonGreedy
    n, trials, steps, batch = 256, 12, 30, 32
    world = make_world('walls5')
    stream = cb.BatchStream(n, seed=99, device='cuda:0')
    env = Gridworld(world, rng=stream)
    eps = np.linspace(0.0, 0.5, n)
    lr = np.linspace(0.5, 1.0, n)
    gamma = np.linspace(0.8, 0.99, n)
    mlr = np.linspace(0.1, 0.9, n)
    mem = DynaQMemory(25, 4, mlr, rng=stream)
    ag = DynaQ(env.observation_space, env.action_space, EpsilonGreedy(eps, rng=stream), None, lr, gamma, mem)
    ag.record = True
    res = ag.train(env, trials, steps, batch)
    torch.cuda.synchronize()
    W = tb.compile_gridworld(world)
    for i in (0, 1, 31, 32, 100, 255):
        rng = tb.Draws(LazyStream(99, i), 1)
        st = tb.dynaq_init(25, 4)
        rec = tb.dynaq_train(W, st, rng, trials, steps, batch, policy=('eps', eps[i]), lr=lr[i], gamma=gamma[i],
                             mem_lr=mlr[i]).arrays()
        got = unpack_run(res, i, 4, W['succ'], W['reward'])
        got.update(Q=ag.Q[i].cpu().numpy(), Mr=mem.rewards[i].cpu().numpy(), draws=int(stream.draw_count[i]))
        rec.update(Q=st['Q'], Mr=st['Mr'], draws=rng.k)
        assert_equal_records(got, rec, ['states', 'actions', 'trial_steps', 'trial_reward', 'replay', 'Q', 'Mr', 'draws'],
                             what='agent %d' % i)


def test_dynaq_results_independent_of_sharding():
    """Agent g's result depends only on (seed, g): a 40-agent shard at base 100 equals the same
    agents inside a 200-agent run."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import DynaQ
    from cobel_rl_b200.policy import EpsilonGreedy
    world = make_world('open5')

    def run(n, base):
        stream = cb.BatchStream(n, seed=5, device='cuda:0', agent_id_base=base)
        env = Gridworld(world, rng=stream)
        ag = DynaQ(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream))
        res = ag.train(env, 20, 50, 32)
        return ag.Q.clone(), res['trial_steps'].clone(), stream.draw_count.clone()
    q1, s1, d1 = run(200, 0)
    q2, s2, d2 = run(40, 100)
    assert torch.equal(q1[100:140], q2) and torch.equal(s1[100:140], s2) and torch.equal(d1[100:140], d2)


def _succ_world_tables(world):
    """Oracle tables of a world built without its dense ``sas`` (large state spaces)."""
    return {'S': world['states'], 'A': 4, 'succ': world['succ'], 'reward': world['rewards'].astype(np.float64),
            'terminal': world['terminals'].astype(np.uint8), 'starts': world['starting_states'].astype(np.int32)}


@pytest.mark.parametrize('shape', [(12, 12), (50, 50)])
def test_dynaq_large_state_spaces(shape):
    """12x12: the largest tables still staged in shared memory.  50x50 (2500 states, 180 KB of tables per agent): Q and
    the memory stay in HBM / L2 (dynaq_warp_kernel<A, false, HBM>), only the replay's dependency masks are on chip
    -- the reference has no limit on the state space (agent/dyna_q.py:128)."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import DynaQ
    from cobel_rl_b200.policy import EpsilonGreedy
    from cobel_rl_b200.misc.gridworld_tools import make_open_field
    h, w = shape
    world = make_open_field(h, w, 0, 1, dense_sas=False)
    stream = cb.BatchStream(3, seed=11, device='cuda:0')
    env = Gridworld(world, rng=stream)
    ag = DynaQ(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream))
    ag.record = True
    res = ag.train(env, 4, 60, 16)
    rt = ag.test(env, 2, 40)
    torch.cuda.synchronize()
    W = _succ_world_tables(world)
    for i in range(3):
        rng = tb.Draws(LazyStream(11, i), 1)
        st = tb.dynaq_init(h * w, 4)
        rec = tb.dynaq_train(W, st, rng, 4, 60, 16).arrays()
        rec2 = tb.tabular_test(W, st['Q'], rng, 2, 40, policy=('eps', 0.1)).arrays()
        got = unpack_run(res, i, 4, W['succ'], W['reward'])
        got.update(Q=ag.Q[i].cpu().numpy(), Mr=ag.M.rewards[i].cpu().numpy(), Ms=ag.M.states[i].cpu().numpy(),
                   Mt=ag.M.terminals[i].cpu().numpy())
        rec.update(Q=st['Q'], Mr=st['Mr'], Ms=st['Ms'], Mt=st['Mt'])
        assert_equal_records(got, rec, ['states', 'actions', 'trial_steps', 'replay', 'Q', 'Mr', 'Ms', 'Mt'], what='agent %d' % i)
        assert_equal_records(unpack_run(rt, i, 4, W['succ'], W['reward']), rec2, ['states', 'actions', 'trial_steps'])
        assert int(stream.draw_count[i]) == rng.k


def test_qagent_large_state_space_hbm_path():
    """QAgent on a 60x60 gridworld (3600 observation keys): the Q table stays in HBM / L2 (q_warp_kernel<A, false, HBM>)."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import QAgent
    from cobel_rl_b200.policy import EpsilonGreedy
    from cobel_rl_b200.misc.gridworld_tools import make_open_field
    world = make_open_field(60, 60, 0, 1, dense_sas=False)
    stream = cb.BatchStream(2, seed=13, device='cuda:0')
    env = Gridworld(world, rng=stream)
    ag = QAgent(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream), None, 0.9, 0.8, rng=stream)
    ag.record = True
    res = ag.train(env, 3, 50, 16)
    torch.cuda.synchronize()
    W = _succ_world_tables(world)
    for i in range(2):
        rng = tb.Draws(LazyStream(13, i), 1)
        st = tb.q_init(3600, 4)
        rec = tb.q_train(W, st, rng, 3, 50, 16).arrays()
        got = unpack_run(res, i, 4, W['succ'], W['reward'])
        got['Q'] = ag.Q[i].cpu().numpy(); rec['Q'] = st['Q']
        assert_equal_records(got, rec, ['states', 'actions', 'trial_steps', 'replay', 'Q'], what='agent %d' % i)
        assert int(stream.draw_count[i]) == rng.k


def test_dynaq_user_stream():
    """Arbitrary pre-drawn uniforms instead of Philox ("load stream from HBM" mode)."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import DynaQ
    from cobel_rl_b200.policy import EpsilonGreedy
    world = make_world('open5')
    u = np.random.default_rng(3).random((2, 20000))
    stream = cb.BatchStream(2, device='cuda:0', user_stream=u)
    env = Gridworld(world, rng=stream)
    ag = DynaQ(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream))
    ag.record = True
    res = ag.train(env, 10, 30, 32)
    torch.cuda.synchronize()
    W = tb.compile_gridworld(world)
    for i in range(2):
        rng = tb.Draws(u[i], 1)
        st = tb.dynaq_init(25, 4)
        rec = tb.dynaq_train(W, st, rng, 10, 30, 32).arrays()
        got = unpack_run(res, i, 4, W['succ'], W['reward'])
        got['Q'] = ag.Q[i].cpu().numpy(); rec['Q'] = st['Q']
        assert_equal_records(got, rec, ['states', 'actions', 'trial_steps', 'replay', 'Q'], what='agent %d' % i)


def test_bad_arguments_raise():
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import DynaQ
    from cobel_rl_b200.policy import EpsilonGreedy
    from cobel_rl_b200.spaces import Box
    stream = cb.BatchStream(2, device='cuda:0')
    env = Gridworld(make_world('open5'), rng=stream)
    with pytest.raises(AssertionError):
        DynaQ(Box([0.], [1.]), env.action_space, EpsilonGreedy(0.1, rng=stream))
    with pytest.raises(AssertionError):
        EpsilonGreedy(1.5)
    ag = DynaQ(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream))
    with pytest.raises(AssertionError):
        ag.train(env, 3, 0, 32)          # steps must be positive


def test_integration_md_binding_stub_reproduces_the_golden(monkeypatch):
    """The reference-side ctypes binding shown in INTEGRATION.md (section 2), executed as written: its
    ``dynaq_train_b200`` drives ``cobel_dynaq_run`` for copies of a reference-shaped agent object and agent 0 must land
    on the golden vector that the unmodified reference produced for the same stream."""
    import os
    import types
    import torch
    import __graft_entry__ as g
    from oracle import cases
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, 'INTEGRATION.md')).read()
    block = text.split('```python')[2].split('```')[0]
    monkeypatch.setenv('COBEL_B200_LIB', g.LIB)
    ns = {}
    exec(compile(block, 'INTEGRATION.md', 'exec'), ns)
    kind, wname, agent_id, args = cases.CASES['dynaq_open5_eps']
    assert agent_id == 0
    world = make_world(wname)
    S = world['states']
    # what the reference's DynaQ / DynaQMemory hold after construction (agent/dyna_q.py:106-138, memory/dyna_q.py:62-75)
    mem = types.SimpleNamespace(rewards=np.zeros((S, 4)), states=np.tile(np.arange(S).reshape(S, 1), 4).astype(int),
                                terminals=np.zeros((S, 4)).astype(int), learning_rate=0.9)
    agent = types.SimpleNamespace(Q=np.zeros((S, 4)), M=mem, learning_rate=0.99, gamma=0.99,
                                  policy=types.SimpleNamespace(epsilon=args['policy'][1]))
    Q, (Mr, Ms, Mt), tsteps, trew = ns['dynaq_train_b200'](agent, world, args['trials'], args['steps'], args['batch'],
                                                           3, cases.SEED)
    torch.cuda.synchronize()
    gold = load_golden('dynaq_open5_eps')
    assert np.array_equal(Q[0].cpu().numpy(), gold['Q']) and np.array_equal(Mr[0].cpu().numpy(), gold['Mr'])
    assert np.array_equal(Ms[0].cpu().numpy(), gold['Ms']) and np.array_equal(Mt[0].cpu().numpy(), gold['Mt'])
    assert np.array_equal(tsteps[0].cpu().numpy(), gold['trial_steps'])
    assert np.array_equal(trew[0].cpu().numpy(), gold['trial_reward'])
    assert not np.array_equal(Q[1].cpu().numpy(), gold['Q'])         # the other copies follow their own streams


def test_pair_kernel_per_agent_hyper_parameters_and_walls():
    """The two-agents-per-warp kernel (selected from two waves of agents on) with a parameter sweep on the agent axis
    (epsilon, learning rates, gamma differ per agent: per-half CDF tables and constants) on a walled world with a
    fixed start state, against the oracle on agents spread over the range; 9001 agents = a half-empty last warp."""
    import cobel_rl_b200 as cb
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.agent import DynaQ
    from cobel_rl_b200.memory import DynaQMemory
    from cobel_rl_b200.policy import EpsilonGreedy
    n, trials, steps, batch = 9001, 30, 40, 32
    g = np.random.default_rng(5)
    eps, lr, gm, mlr = g.uniform(0.02, 0.5, n), g.uniform(0.3, 1.0, n), g.uniform(0.5, 0.99, n), g.uniform(0.2, 1.0, n)
    world = make_world('walls5')
    stream = cb.BatchStream(n, seed=31337, device='cuda:0')
    env = Gridworld(world, rng=stream)
    mem = DynaQMemory(env.n_states, 4, mlr, rng=stream)
    ag = DynaQ(env.observation_space, env.action_space, EpsilonGreedy(eps, rng=stream), None, lr, gm, mem)
    res = ag.train(env, trials, steps, batch)
    torch.cuda.synchronize()
    assert torch.equal(res['n_replay'], res['n_steps'] * batch)
    W = tb.compile_gridworld(world)
    for i in (0, 1, 2, 4499, 4500, 8998, 8999, 9000):
        rng = tb.Draws(LazyStream(31337, i), 1)
        st = tb.dynaq_init(25, 4)
        rec = tb.dynaq_train(W, st, rng, trials, steps, batch, policy=('eps', float(eps[i])), lr=float(lr[i]),
                             gamma=float(gm[i]), mem_lr=float(mlr[i])).arrays()
        assert np.array_equal(rec['trial_steps'], res['trial_steps'][i].cpu().numpy()), 'agent %d' % i
        assert np.array_equal(rec['trial_reward'], res['trial_reward'][i].cpu().numpy())
        assert np.array_equal(st['Q'], ag.Q[i].cpu().numpy()) and np.array_equal(st['Mr'], mem.rewards[i].cpu().numpy())
        assert np.array_equal(st['Ms'], mem.states[i].cpu().numpy()) and np.array_equal(st['Mt'], mem.terminals[i].cpu().numpy())
        assert rng.k == int(stream.draw_count[i])
