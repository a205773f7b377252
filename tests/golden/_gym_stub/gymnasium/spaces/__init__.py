import numpy as np


class Space:
    pass


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        self.dtype = np.dtype(dtype)
        low = np.asarray(low)
        high = np.asarray(high)
        if shape is None:
            shape = low.shape
        self._shape = tuple(shape)
        self.low = np.broadcast_to(low, self._shape).astype(self.dtype)
        self.high = np.broadcast_to(high, self._shape).astype(self.dtype)
        self._rng = np.random.default_rng(seed)

    @property
    def shape(self):
        return self._shape

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self._shape and bool(np.all(x >= self.low) and np.all(x <= self.high))


class Discrete(Space):
    def __init__(self, n):
        self.n = n
        self.shape = ()
