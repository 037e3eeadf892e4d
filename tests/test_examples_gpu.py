"""GPU: the ported reference demos run end to end (batched and single-agent) and the agents learn."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'examples'))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('kind,trials', [('dyna_q', 60), ('sr', 80), ('pma', 12), ('sfma', 40)])
def test_gridworld_demos(kind, trials):
    import demo_gridworld
    agent, trace = demo_gridworld.simulation(kind, 64, trials)
    assert trace.shape == (64, trials)
    assert trace[:, -5:].mean() < trace[:, :5].mean(), 'escape latency should drop with training'
    agent1, trace1 = demo_gridworld.simulation(kind, None, 5)          # single agent: reference shapes
    assert trace1.shape == (5,) and agent1.Q.dim() == 2 if kind != 'sr' else agent1.SR.dim() == 2
    # agent 0 of the batch is the single-agent run with the same seed
    assert np.array_equal(trace[0, :5], trace1)


@pytest.mark.parametrize('template', ['linear_track', 'grid', 'hexagonal', 't_maze'])
def test_topology_demo(template):
    import demo_topology
    agent, trace = demo_topology.simulation(template, 32, 120)
    assert trace.shape == (32, 120) and trace[:, -10:].mean() < trace[:, :10].mean()
    assert agent.Q.shape[0] == 32


def test_dyna_dqn_demo_learns():
    """examples/demo_dyna_dqn.py (demo/gridworld/demo_dyna_dqn.py for N agents): the batch of Dyna-DQN agents learns."""
    import demo_dyna_dqn
    agent, trace = demo_dyna_dqn.simulation(32, 40)
    assert trace.shape == (32, 40)
    assert trace[:, -10:].mean() < 0.7 * trace[:, :10].mean(), 'escape latency should drop with training'
    q = agent.predict_on_batch(np.arange(25))
    assert q.shape == (32, 25, 4) and bool(q.isfinite().all())
