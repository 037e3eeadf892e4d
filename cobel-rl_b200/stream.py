"""The per-agent random stream object that takes the place of ``numpy.random.Generator``.

In the reference one ``Generator`` can be passed as ``rng=`` to the environment, the
policies, the memory and the agent (interface/gridworld.py:83, policy/policy.py:29-30,
memory/dyna_q.py:69, agent/q.py:138); here one ``BatchStream`` is passed to the same
constructors.  It fixes the number of agents N, the device, the Philox seed and owns the
per-agent draw counters (include/cobel_b200.h: CobelStream).
"""
import torch

from . import _lib


class BatchStream:
    """N independent uniform streams (Philox4x32-10, see oracle/philox.py for the host definition).

    Parameters
    ----------
    n_agents : int or None
        Number of agents.  ``None`` = a single agent whose tensors are exposed with
        the reference's shapes (no leading agent axis).
    seed : int
        64-bit Philox key shared by all agents of a run.
    device : torch.device or str
        CUDA device that holds every per-agent tensor.
    agent_id_base : int
        Global id of local agent 0 (multi-GPU shards use contiguous ranges, so results
        do not depend on the number of GPUs).
    user_stream : torch.Tensor or None
        Optional ``[N, L]`` float64 tensor of pre-drawn uniforms used instead of Philox.
    """

    def __init__(self, n_agents=None, seed=0x5EED, device='cuda', agent_id_base=0, user_stream=None):
        self.single = n_agents is None
        self.n_agents = 1 if n_agents is None else int(n_agents)
        assert self.n_agents > 0
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.device = torch.device(device)
        self.agent_id_base = int(agent_id_base)
        self.draw_count = torch.zeros(self.n_agents, dtype=torch.int64, device=self.device)
        self.user_stream = None
        if user_stream is not None:
            us = torch.as_tensor(user_stream, dtype=torch.float64)
            if us.dim() == 1:
                us = us.unsqueeze(0)
            assert us.shape[0] == self.n_agents
            self.user_stream = us.to(self.device).contiguous()

    def c_struct(self):
        return _lib.Stream(self.seed, self.agent_id_base, self.draw_count.data_ptr(),
                           _lib.ptr(self.user_stream),
                           0 if self.user_stream is None else self.user_stream.shape[1])

    def next(self, n_draws=1):
        """Consume the next ``n_draws`` uniforms of every agent -> ``[N, n_draws]`` tensor."""
        out = torch.empty((self.n_agents, n_draws), dtype=torch.float64, device=self.device)
        s = self.c_struct()
        _lib.call('cobel_stream_next', self.device, s, self.n_agents, n_draws, out.data_ptr(), cuda_stream(self.device))
        return out

    def integers(self, n):
        """One ``Generator.integers(n)`` per agent: ``min(floor(u*n), n-1)``."""
        u = self.next(1)[:, 0]
        return torch.clamp((u * n).floor().to(torch.int64), max=n - 1)

    def param(self, x, name='parameter'):
        """Broadcast a scalar / sequence / tensor hyper-parameter to a ``[N]`` fp64 device tensor."""
        if isinstance(x, (int, float)):                 # the common case: one value for all agents -- built once per value
            cache = self.__dict__.setdefault('_param_cache', {})
            t = cache.get(float(x))
            if t is None:
                if len(cache) >= 64:
                    cache.clear()
                t = cache[float(x)] = torch.full((self.n_agents,), float(x), dtype=torch.float64, device=self.device)
            return t
        t = torch.as_tensor(x, dtype=torch.float64).to(self.device).reshape(-1)
        if t.numel() == 1:
            t = t.expand(self.n_agents)
        assert t.numel() == self.n_agents, '%s must be a scalar or have one entry per agent' % name
        return t.contiguous()


def cuda_stream(device):
    """Raw ``cudaStream_t`` of torch's current stream on ``device``."""
    if device.type != 'cuda':
        raise _lib.CobelError('cobel_rl_b200 runs on CUDA devices only (got %s); there is no CPU fallback' % device)
    return torch.cuda.current_stream(device).cuda_stream
