"""The reference's gridworld demos (demo/gridworld/demo_{dyna_q,sr,pma,sfma}.py) on the B200 path:
same worlds, agents and hyper-parameters, N independent agents per launch, no widgets.

    python examples/demo_gridworld.py dyna_q --agents 4096
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import cobel_rl_b200 as cb  # noqa: E402
from cobel_rl_b200.agent import DynaQ, PMA, SFMA, SR  # noqa: E402
from cobel_rl_b200.interface import Gridworld  # noqa: E402
from cobel_rl_b200.memory import PMAMemory, SFMAMemory  # noqa: E402
from cobel_rl_b200.memory.utils.metrics import DR  # noqa: E402
from cobel_rl_b200.misc.gridworld_tools import make_gridworld, make_open_field  # noqa: E402
from cobel_rl_b200.monitor import EscapeLatencyMonitor  # noqa: E402
from cobel_rl_b200.policy import EpsilonGreedy  # noqa: E402

WALLS = [(3, 4), (4, 3), (8, 9), (9, 8), (13, 14), (14, 13), (18, 19), (19, 18)]


def walled_world():
    """demo/gridworld/demo_pma.py:35-60, demo_sfma.py:36-60."""
    world = make_gridworld(5, 5, terminals=[4], rewards=np.array([[4, 10]]), goals=[4], invalid_transitions=WALLS)
    world['starting_states'] = np.array([12])
    return world


def simulation(kind, n_agents=None, trials=None, steps=50, seed=0x5EED, device='cuda:0'):
    """One simulation run of `kind` for n_agents independent agents (None: a single agent with the
    reference's tensor shapes).  Returns (agent, escape-latency trace [N, trials])."""
    rng = cb.BatchStream(n_agents, seed=seed, device=device)
    if kind == 'dyna_q':        # demo/gridworld/demo_dyna_q.py:36-56
        trials = trials or 500
        env = Gridworld(make_open_field(5, 5, 0, 1), rng=rng)
        el = EscapeLatencyMonitor(trials, steps, n_agents)
        agent = DynaQ(env.observation_space, env.action_space, EpsilonGreedy(0.1), EpsilonGreedy(0.0),
                      custom_callbacks={'on_trial_end': [el.update]})
        agent.train(env, trials, steps, 32)
    elif kind == 'sr':          # demo/gridworld/demo_sr.py:36-56
        trials = trials or 500
        env = Gridworld(make_open_field(5, 5, 0, 1), rng=rng)
        el = EscapeLatencyMonitor(trials, steps, n_agents)
        agent = SR(env.observation_space, env.action_space, EpsilonGreedy(0.1), EpsilonGreedy(0.0),
                   custom_callbacks={'on_trial_end': [el.update]})
        agent.train(env, trials, steps)
    elif kind == 'pma':         # demo/gridworld/demo_pma.py:29-72
        trials = trials or 250
        env = Gridworld(walled_world(), rng=rng)
        el = EscapeLatencyMonitor(trials, steps, n_agents)
        memory = PMAMemory(env.world['sas'], EpsilonGreedy(), gamma_q=0.99)
        agent = PMA(env.observation_space, env.action_space, EpsilonGreedy(), memory,
                    custom_callbacks={'on_trial_end': [el.update]})
        agent.mask_actions = True
        agent.train(env, trials, steps, 32)
    elif kind == 'sfma':        # demo/gridworld/demo_sfma.py:30-80
        trials = trials or 250
        world = walled_world()
        env = Gridworld(world, rng=rng)
        el = EscapeLatencyMonitor(trials, steps, n_agents)
        memory = SFMAMemory(DR(5, 5, world['sas'], 0.9, WALLS), 25, 4)
        agent = SFMA(env.observation_space, env.action_space, EpsilonGreedy(), memory,
                     custom_callbacks={'on_trial_end': [el.update]})
        agent.mask_actions = True
        agent.M.mode = 'reverse'
        agent.train(env, trials, steps, 32)
    else:
        raise ValueError(kind)
    return agent, el.get_trace()


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('kind', choices=['dyna_q', 'sr', 'pma', 'sfma'])
    ap.add_argument('--agents', type=int, default=1024)
    ap.add_argument('--trials', type=int, default=None)
    args = ap.parse_args()
    _, trace = simulation(args.kind, args.agents, args.trials)
    trace = np.atleast_2d(trace)
    print('%s: %d agents x %d trials; mean escape latency first 10 trials %.1f, last 10 trials %.1f steps'
          % (args.kind, trace.shape[0], trace.shape[1], trace[:, :10].mean() + 1, trace[:, -10:].mean() + 1))
