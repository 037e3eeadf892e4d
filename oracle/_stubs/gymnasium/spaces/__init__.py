"""Minimal `gymnasium.spaces` stand-in (Discrete, Box, Dict, Tuple)."""
import numpy as np


class Space:
    pass


class Discrete(Space):
    def __init__(self, n, start=0):
        self.n = np.int64(n)
        self.start = start


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float64):
        self.low, self.high, self.dtype = low, high, dtype
        if shape is None:
            shape = np.shape(low)
        self.shape = tuple(shape)


class Dict(Space):
    def __init__(self, spaces=None, **kwargs):
        self.spaces = dict(spaces or {}, **kwargs)

    def items(self):
        return self.spaces.items()


class Tuple(Space):
    def __init__(self, spaces):
        self.spaces = list(spaces)
