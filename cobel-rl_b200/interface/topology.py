"""Topology-graph environment (reference: interface/topology.py:29-204).

Nodes are mapped to their insertion index; ``neighbors`` lists become the
successor table.  Observations are node poses (6 floats, topology.py:174-193);
``discrete=True`` additionally exposes node indices through a Discrete
observation space so that the tabular agents that require one (DynaQ, SR, PMA,
SFMA: agent/dyna_q.py:117-122) can run on graphs.
Simulator-rendered observations are out of scope.
"""
import numpy as np
import torch

from .interface import Interface
from ..spaces import Box, Discrete


class Topology(Interface):
    def __init__(self, nodes, starting_nodes=None, simulator=None, widget=None, rng=None, discrete=False):
        super().__init__(widget, rng)
        assert simulator is None, 'simulator observations are out of scope of the B200 path'
        self.nodes = nodes
        self.node_ids = list(nodes.keys())
        self._index = {n: i for i, n in enumerate(self.node_ids)}
        if starting_nodes is None:      # topology.py:100-107
            starting_nodes = [n for n, node in nodes.items() if not node['terminal']]
        self.starting_nodes = list(starting_nodes)
        n_act = {len(node['neighbors']) for node in nodes.values()}
        assert len(n_act) == 1, 'all nodes must have the same number of neighbours'
        succ = np.array([[self._index[m] for m in nodes[n]['neighbors']] for n in self.node_ids], dtype=np.int32)
        self._set_tables(succ, [float(nodes[n]['reward']) for n in self.node_ids],
                         [1 if nodes[n]['terminal'] else 0 for n in self.node_ids],
                         [self._index[n] for n in self.starting_nodes])
        self._pose = torch.as_tensor(np.array([nodes[n]['pose'] for n in self.node_ids], dtype=np.float64)
                                     ).to(self.rng.device)
        # observation key of a node = index of the first node with the same pose
        # (QAgent keys its Q rows by the pose tuple, agent/q.py:155)
        first = {}
        self._obs_key = np.array([first.setdefault(tuple(np.asarray(nodes[n]['pose']).flatten()), i)
                                  for i, n in enumerate(self.node_ids)], dtype=np.int32)
        self.discrete = discrete
        self.action_space = Discrete(succ.shape[1])
        if discrete:
            self.observation_space = Discrete(len(self.node_ids))
        else:
            self.observation_space = Box(low=np.array([-np.inf] * 3 + [0.0] * 3),
                                         high=np.array([np.inf] * 3 + [360.0] * 3), dtype=np.float64)
        # topology.py:109 -- one draw in the constructor
        self._launch_reset()

    @property
    def current_node(self):
        if self.rng.single:
            return self.node_ids[int(self._current[0])]
        return [self.node_ids[int(i)] for i in self._current.tolist()]

    def get_observation(self):
        """topology.py:174-193: the pose of the current node (or its index with ``discrete=True``)."""
        cur = self._current.long()
        return self._out(cur) if self.discrete else self._out(self._pose[cur].clone())

    _observation = get_observation

    def step(self, action):
        """topology.py:126-157: returns ``end_trial`` as both terminated and truncated."""
        _, reward, end = self._launch_step(action)
        end = self._out(end)
        return self.get_observation(), self._out(reward), end, end, {}

    def reset(self):
        """topology.py:159-172."""
        self._launch_reset()
        return self.get_observation(), {}

    def get_position(self):
        return self._out(self._pose[self._current.long()].clone())
