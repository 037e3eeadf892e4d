"""Stream-fed stand-in for ``numpy.random.Generator`` (test infrastructure).

Duck-types the three Generator methods the reference's hot path calls
(``random``, ``integers``, ``choice``) and serves them from a pre-drawn uniform
stream, so the unmodified reference and the restatement/CUDA path consume the
same numbers in the same order.

``Generator.choice(a, p=p)`` is NumPy's own inverse-CDF algorithm:
``searchsorted(cumsum(p) / cumsum(p)[-1], u, side='right')``;
``choice(a)`` / ``integers(n)`` map one uniform to ``min(floor(u*n), n-1)``.
"""
import numpy as np


class StreamRNG:
    def __init__(self, u):
        self.u = u          # indexable stream of float64 in [0, 1)
        self.k = 0          # number of draws consumed so far

    def _draw(self, n=None):
        if n is None:
            v = self.u[self.k]
            self.k += 1
            return float(v)
        v = np.asarray(self.u[self.k:self.k + n], dtype=np.float64)
        self.k += n
        return v

    def random(self, size=None):
        return self._draw(size)

    def integers(self, low, high=None, size=None):
        if high is None:
            low, high = 0, low
        n = int(high) - int(low)
        u = self._draw(size)
        if size is None:
            return min(int(low) + int(np.floor(u * n)), int(high) - 1)
        return np.minimum((int(low) + np.floor(u * n)).astype(np.int64), int(high) - 1)

    def choice(self, a, size=None, p=None):
        a = np.arange(a) if np.ndim(a) == 0 else np.asarray(a)
        if p is None:
            return a[self.integers(0, len(a), size)]
        cdf = np.cumsum(p)
        cdf /= cdf[-1]
        return a[cdf.searchsorted(self._draw(size), side='right')]
