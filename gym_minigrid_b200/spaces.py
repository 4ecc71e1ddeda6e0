"""Minimal stand-ins for gym.spaces.{Box, Discrete, Dict} (the image has no gym).  Same attribute
names as the reference uses (minigrid.py:792-807, dynamicobstacles.py:32)."""
import numpy as np


class Space:
    pass


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low, self.high = low, high
        self.shape = tuple(shape) if shape is not None else np.shape(low)
        self.dtype = np.dtype(dtype)

    def __repr__(self):
        return "Box(%s, %s, %s, %s)" % (self.low, self.high, self.shape, self.dtype)


class Discrete(Space):
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def sample(self):
        return int(np.random.randint(self.n))

    def __repr__(self):
        return "Discrete(%d)" % self.n


class Dict(Space):
    def __init__(self, spaces):
        self.spaces = dict(spaces)

    def __getitem__(self, k):
        return self.spaces[k]

    def __repr__(self):
        return "Dict(%s)" % ", ".join("%s:%r" % kv for kv in self.spaces.items())
