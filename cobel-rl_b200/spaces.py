"""Minimal observation / action spaces (stand-ins for ``gymnasium.spaces``, which the
reference uses at interface/gridworld.py:85-86 and interface/topology.py:110-122 and
which is not a dependency here)."""
import numpy as np


class Space:
    pass


class Discrete(Space):
    def __init__(self, n):
        self.n = np.int64(n)

    def __repr__(self):
        return 'Discrete(%d)' % int(self.n)


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float64):
        self.low, self.high, self.dtype = np.asarray(low), np.asarray(high), dtype
        self.shape = tuple(np.shape(low) if shape is None else shape)

    def __repr__(self):
        return 'Box(shape=%s)' % (self.shape,)
