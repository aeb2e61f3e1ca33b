"""Minimal stand-in for ``gym.spaces.Box`` (gym is not a dependency of this package).

The reference builds ``gym.spaces.Box(low, high)`` in ``SnakeGymEnv.py:60-79``; callers only read
``.shape`` (``ppo/train.py:73-74``, ``ars/train.py:33-36``), ``.low`` and ``.high``.
"""
import numpy as np


class Box:
    def __init__(self, low, high, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        if self.low.shape != self.high.shape:
            raise ValueError("low and high must have the same shape")
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def sample(self, rng=None):
        rng = rng or np.random.default_rng()
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return rng.uniform(lo, hi).astype(self.dtype)

    def __repr__(self):
        return "Box(%s)" % (self.shape,)
