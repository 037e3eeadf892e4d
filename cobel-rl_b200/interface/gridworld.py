"""Gridworld environment (reference: interface/gridworld.py:33-156).

The reference steps one agent through a dense ``sas[S,4,S]`` row (arg-max scan).
Here the world is compiled once to ``succ[S,4]`` / ``reward[S]`` / ``terminal[S]``
/ ``starts[K]`` device tables that the fused agent kernels read; ``step``/``reset``
are kept as batched table lookups for interactive use.
"""
from typing import Any, TypedDict

import numpy as np
import torch

from .interface import Interface
from ..spaces import Discrete


class WorldDict(TypedDict, total=False):   # interface/gridworld.py:17-30
    width: int
    height: int
    states: int
    rewards: Any
    terminals: Any
    sas: Any
    succ: Any          # extension: successor table; `sas` may be None for very large worlds
    starting_states: Any
    invalid_transitions: list
    invalid_states: list
    wind: Any
    goals: list
    coordinates: Any
    deterministic: bool


def successor_table(world):
    """``argmax(sas[s, a, :])`` for every (s, a) (gridworld.py:116-117), or the
    builder-provided ``succ`` when the dense tensor was not materialised."""
    if world.get('succ') is not None:
        return np.asarray(world['succ'], dtype=np.int32)
    return np.argmax(np.asarray(world['sas']), axis=2).astype(np.int32)


class Gridworld(Interface):
    def __init__(self, world, widget=None, rng=None):
        super().__init__(widget, rng)
        self.world = world
        self.deterministic = bool(world.get('deterministic', True))
        assert self.deterministic or world.get('sas') is not None, 'non-deterministic worlds need the dense sas'

        self.observation_space = Discrete(world['states'])
        self.action_space = Discrete(4)
        self._set_tables(successor_table(world), world['rewards'], world['terminals'], world['starting_states'],
                         None if self.deterministic else world['sas'])
        self._coordinates = torch.as_tensor(np.asarray(world['coordinates']), dtype=torch.float64).to(self.rng.device)
        self._current = torch.zeros(self.rng.n_agents, dtype=torch.int32, device=self.rng.device)
        self.reset()      # gridworld.py:89 -- consumes one draw per agent, like the reference

    @property
    def current_state(self):
        return self._out(self._current.long())

    @property
    def current_coordinates(self):
        return self._out(self._coordinates[self._current.long()])

    def step(self, action):
        """gridworld.py:92-129 for all agents: (state, reward, end_trial, False, {})."""
        cur, reward, end = self._launch_step(action)
        return self._out(cur.long()), self._out(reward), self._out(end), False, {}

    def reset(self):
        """gridworld.py:131-145: uniform draw over the starting states."""
        return self._out(self._launch_reset().long()), {}

    def get_position(self):
        return self._out(self._coordinates[self._current.long()].clone())
