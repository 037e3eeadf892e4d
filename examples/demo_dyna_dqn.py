"""The reference's Dyna-DQN demo (demo/gridworld/demo_dyna_dqn.py) on the batched path: N independent agents, each with
its own copy of the network, advance in lock-step; the table side of every step (environment, policy draw, memory
store, replay batch) is one CUDA launch for all agents, the networks are evaluated together (torch.func.vmap).

    python examples/demo_dyna_dqn.py --agents 64 --trials 60
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import cobel_rl_b200 as cb  # noqa: E402
from cobel_rl_b200.agent import DynaDQN  # noqa: E402
from cobel_rl_b200.interface import Gridworld  # noqa: E402
from cobel_rl_b200.misc.gridworld_tools import make_open_field  # noqa: E402
from cobel_rl_b200.network import BatchedTorchNetwork  # noqa: E402
from cobel_rl_b200.policy import EpsilonGreedy  # noqa: E402


class Model(torch.nn.Module):       # demo/gridworld/demo_dyna_dqn.py:38-55
    def __init__(self, input_size, output_size):
        super().__init__()
        self.layer_dense_1 = torch.nn.Linear(input_size, 64)
        self.layer_dense_2 = torch.nn.Linear(64, 64)
        self.layer_output = torch.nn.Linear(64, output_size)
        self.double()

    def forward(self, x):
        x = torch.relu(self.layer_dense_1(x))
        x = torch.relu(self.layer_dense_2(x))
        return self.layer_output(x)


def simulation(n_agents=64, trials=60, steps=50, seed=0x5EED, device='cuda:0'):
    """Returns (agent, escape-latency trace [N, trials]) -- demo/gridworld/demo_dyna_dqn.py:58-95 for N agents."""
    rng = cb.BatchStream(n_agents, seed=seed, device=device)
    env = Gridworld(make_open_field(5, 5, 0, 1), rng=rng)
    torch.manual_seed(seed)
    network = BatchedTorchNetwork([Model(25, 4) for _ in range(n_agents)], device=device)
    trace = []
    agent = DynaDQN(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=rng), network, gamma=0.8,
                    policy_test=EpsilonGreedy(0.0, rng=rng),
                    custom_callbacks={'on_trial_end': [lambda logs: trace.append(logs['steps'].clone())]})
    agent.train(env, trials, steps, 32)
    return agent, torch.stack(trace, dim=1).cpu().numpy()


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--agents', type=int, default=64)
    ap.add_argument('--trials', type=int, default=60)
    args = ap.parse_args()
    agent, trace = simulation(args.agents, args.trials)
    np.set_printoptions(precision=2, floatmode='fixed')
    print('mean escape latency, first / last 10 trials: %.1f / %.1f' % (trace[:, :10].mean(), trace[:, -10:].mean()))
    print('Q-function of agent 0:')
    print(agent.predict_on_batch(np.arange(25))[0].cpu().numpy())
