"""Host definition of the per-agent uniform random stream (Philox4x32-10).

Test infrastructure (see oracle/__init__.py).  The CUDA kernels implement the
same generator on the device (cobel-rl_b200/csrc/rng.cuh); a GPU test checks
that both sides produce identical doubles.

Stream contract
-----------------------------------------------------
The k-th uniform consumed by agent ``i`` of a run with seed ``seed`` is

    block  = Philox4x32-10(counter = (k>>1 lo32, k>>1 hi32, i lo32, i hi32),
                           key     = (seed lo32, seed hi32))
    (a, b) = block[0:2] if k is even else block[2:4]
    u      = ((a >> 5) * 2**26 + (b >> 6)) * 2**-53          in [0, 1)

i.e. the classic 53-bit construction.  All components of one agent (environment,
policies, memory, agent) consume this single stream in program order, exactly
like one shared ``numpy.random.Generator`` passed as ``rng=`` to every reference
object.
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32_10(ctr, key):
    """Vectorised Philox4x32-10.

    ctr: (..., 4) uint32-valued array, key: (..., 2).  Returns (..., 4) uint32.
    """
    c = [np.asarray(ctr[..., j], dtype=np.uint64) for j in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint64)
    k1 = np.asarray(key[..., 1], dtype=np.uint64)
    for r in range(10):
        p0 = _M0 * c[0]
        p1 = _M1 * c[2]
        hi0, lo0 = p0 >> _S32, p0 & _MASK
        hi1, lo1 = p1 >> _S32, p1 & _MASK
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(_W0)) & _MASK
        k1 = (k1 + np.uint64(_W1)) & _MASK
    return np.stack(c, axis=-1).astype(np.uint32)


def uniforms(seed, agent, n_draws, first=0):
    """Return draws ``first .. first+n_draws-1`` of agent ``agent`` as float64."""
    k = np.arange(first, first + n_draws, dtype=np.uint64)
    blk = k >> np.uint64(1)
    ctr = np.empty((n_draws, 4), dtype=np.uint64)
    ctr[:, 0] = blk & _MASK
    ctr[:, 1] = blk >> _S32
    ctr[:, 2] = np.uint64(agent) & _MASK
    ctr[:, 3] = np.uint64(agent) >> _S32
    key = np.empty((n_draws, 2), dtype=np.uint64)
    key[:, 0] = np.uint64(seed) & _MASK
    key[:, 1] = np.uint64(seed) >> _S32
    out = philox4x32_10(ctr, key).astype(np.uint64)
    odd = (k & np.uint64(1)).astype(bool)
    a = np.where(odd, out[:, 2], out[:, 0])
    b = np.where(odd, out[:, 3], out[:, 1])
    return ((a >> np.uint64(5)).astype(np.float64) * 67108864.0
            + (b >> np.uint64(6)).astype(np.float64)) * (1.0 / 9007199254740992.0)


class LazyStream:
    """Array-like view of one agent's stream that generates blocks on demand."""

    def __init__(self, seed, agent, chunk=1 << 14):
        self.seed, self.agent, self.chunk = seed, agent, chunk
        self._buf = np.empty(0)

    def _grow(self, n):
        while self._buf.shape[0] < n:
            more = uniforms(self.seed, self.agent, self.chunk, self._buf.shape[0])
            self._buf = np.concatenate([self._buf, more])

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            self._grow(idx.stop)
        else:
            self._grow(idx + 1)
        return self._buf[idx]
