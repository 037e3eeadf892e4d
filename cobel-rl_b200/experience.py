"""Experience dictionaries and their batched device form.

The reference passes experiences around as dicts (``Experience`` TypedDicts: memory/dyna_q.py:8-14,
agent/q.py:17-23, agent/sr.py:16-22, memory/pma.py:11-17, memory/sfma.py:12-18).  Here a field holds one value
per agent (a scalar, or an ``[N]`` tensor), and a list of B dicts travels to the kernels as a
``CobelExperiences`` structure of ``[N, B]`` arrays (include/cobel_b200.h).
"""
from typing import Any, TypedDict

import torch

from . import _lib


class Experience(TypedDict, total=False):
    state: Any
    action: Any
    reward: Any
    next_state: Any
    terminal: Any          # holds 1 - end_trial, like the reference's field of that name
    td: Any                # added by update_q


FIELDS = (('state', torch.int32), ('action', torch.int32), ('reward', torch.float64),
          ('next_state', torch.int32), ('terminal', torch.int32))


class ExperienceBatch:
    """``[N, B]`` device arrays of B experiences per agent."""

    def __init__(self, stream, batch, with_td=False):
        self.stream, self.batch = stream, int(batch)
        n, dev = stream.n_agents, stream.device
        for name, dt in FIELDS:
            setattr(self, name, torch.zeros((n, max(self.batch, 1)), dtype=dt, device=dev))
        self.td = torch.zeros((n, max(self.batch, 1)), dtype=torch.float64, device=dev) if with_td else None

    @classmethod
    def from_dicts(cls, stream, experiences, with_td=False):
        """One Experience dict or a list of them; every field a scalar or an ``[N]`` tensor / sequence."""
        if isinstance(experiences, dict):
            experiences = [experiences]
        out = cls(stream, len(experiences), with_td)
        n, dev = stream.n_agents, stream.device
        for b, e in enumerate(experiences):
            for name, dt in FIELDS:
                v = e[name]
                if name == 'terminal' and isinstance(v, bool):
                    v = int(v)
                t = torch.as_tensor(v, device=dev).reshape(-1).to(dt)
                assert t.numel() in (1, n), 'experience field %r must be a scalar or have one entry per agent' % name
                getattr(out, name)[:, b] = t
        return out

    def c_struct(self):
        return _lib.Experiences(self.batch, 0, self.state.data_ptr(), self.action.data_ptr(), self.reward.data_ptr(),
                                self.next_state.data_ptr(), self.terminal.data_ptr(), _lib.ptr(self.td))

    def to_dicts(self, count=None):
        """List of B Experience dicts (``[N]`` tensors, Python scalars for a single-agent stream)."""
        single = self.stream.single
        out = []
        for b in range(self.batch if count is None else count):
            d = {}
            for name, _ in FIELDS + ((('td', None),) if self.td is not None else ()):
                col = getattr(self, name)[:, b]
                d[name] = col[0].item() if single else col
            out.append(d)
        return out
