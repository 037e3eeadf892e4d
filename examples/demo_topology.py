"""The reference's topology demo (demo/topology/demo.py:40-103) on the B200 path: a tabular
Q-learning agent on one of the template graphs, pose observations, no replay (batch 0).

    python examples/demo_topology.py hexagonal --agents 4096
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import cobel_rl_b200 as cb  # noqa: E402
from cobel_rl_b200.agent import QAgent  # noqa: E402
from cobel_rl_b200.interface import Topology  # noqa: E402
from cobel_rl_b200.misc.topology_tools import grid, hexagonal, linear_track, t_maze  # noqa: E402
from cobel_rl_b200.monitor import EscapeLatencyMonitor  # noqa: E402
from cobel_rl_b200.policy import EpsilonGreedy  # noqa: E402

TEMPLATES = {'linear_track': lambda: linear_track(10, 2, 1.0, 20, 'right'), 'grid': lambda: grid(5),
             'hexagonal': lambda: hexagonal(10), 't_maze': lambda: t_maze(6, 3, 2)}


def simulation(template, n_agents=None, trials=500, steps=50, batch_size=0, seed=0x5EED, device='cuda:0'):
    rng = cb.BatchStream(n_agents, seed=seed, device=device)
    nodes, starting_nodes = TEMPLATES[template]()
    env = Topology(nodes, starting_nodes, rng=rng)
    el = EscapeLatencyMonitor(trials, steps, n_agents)
    agent = QAgent(env.observation_space, env.action_space, EpsilonGreedy(0.1), EpsilonGreedy(0.0),
                   custom_callbacks={'on_trial_end': [el.update]})
    agent.train(env, trials, steps, batch_size)
    return agent, el.get_trace()


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('template', choices=sorted(TEMPLATES))
    ap.add_argument('--agents', type=int, default=1024)
    ap.add_argument('--trials', type=int, default=500)
    args = ap.parse_args()
    _, trace = simulation(args.template, args.agents, args.trials)
    trace = np.atleast_2d(trace)
    print('%s: %d agents x %d trials; mean escape latency first 10 trials %.1f, last 10 trials %.1f steps'
          % (args.template, trace.shape[0], trace.shape[1], trace[:, :10].mean() + 1, trace[:, -10:].mean() + 1))
