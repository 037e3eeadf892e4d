"""CPU: the reference's own known-answer tests for the environments, restated on the
compiled tables (unit_tests/test_gridworld.py:12-41, unit_tests/test_topology.py:60-109),
and the product's world builders against the golden trajectories' transition structure."""
import numpy as np

from oracle import tabular as tb
from cobel_rl_b200.misc import gridworld_tools, topology_tools


def test_gridworld_known_answer():
    world = gridworld_tools.make_gridworld(5, 5, [0], np.array([[0, 10.]]), starting_states=[24])
    assert world['states'] == 25
    W = tb.compile_gridworld(world)
    assert W['A'] == 4 and list(W['starts']) == [24]
    s = 24
    states, rewards, terminals = [], [], []
    for a in [0, 0, 0, 0, 0, 1, 1, 1, 1]:
        s, r, end = tb.env_step(W, s, a)
        states.append(s); rewards.append(r); terminals.append(bool(end))
    assert states == [23, 22, 21, 20, 20, 15, 10, 5, 0]
    assert rewards == [0] * 8 + [10.]
    assert terminals == [False] * 8 + [True]


def test_topology_known_answer():
    nodes, starting = topology_tools.t_maze(4, 3, 1)
    W = tb.compile_topology(nodes, starting)
    assert W['A'] == 4
    assert [W['ids'][i] for i in W['starts']] == ['10']
    s = int(W['starts'][0])
    visited, rewards, terminals = [], [], []
    for a in [1, 1, 1, 1, 1, 2, 2, 2]:
        s, r, end = tb.env_step(W, s, a)
        visited.append(W['ids'][s]); rewards.append(r); terminals.append(bool(end))
    assert visited == ['9', '8', '7', '3', '3', '4', '5', '6']
    assert rewards == [0.] * 7 + [1.]
    assert terminals == [False] * 7 + [True]


def test_succ_matches_dense_sas():
    for world in (gridworld_tools.make_open_field(6, 4, 3, 1),
                  gridworld_tools.make_gridworld(5, 5, invalid_transitions=[(1, 2), (2, 1), (7, 12)], invalid_states=[18])):
        assert np.array_equal(np.argmax(world['sas'], axis=2), world['succ'])
        assert np.array_equal(world['sas'].sum(axis=2), np.ones((world['states'], 4)))
