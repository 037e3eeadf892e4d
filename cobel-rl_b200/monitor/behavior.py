"""Escape-latency and reward monitors (reference: monitor/behavior.py:16-250, numeric part only).

The reference monitors are callbacks that read ``logs['trial']`` / ``logs['steps']`` /
``logs['trial_reward']`` once per trial (behavior.py:73-97, 170-199) and optionally plot with
pyqtgraph.  These do the same on the batched ``logs`` (values are ``[N]`` tensors, or scalars
for a single-agent stream) and keep one trace row per agent; register ``monitor.update`` under
``custom_callbacks['on_trial_end']``, or fill them after the fact with ``from_result``.
"""
import numpy as np
import torch


class _TrialTrace:
    key = ''

    def __init__(self, trials, n_agents=None):
        self.trials = int(trials)
        self.single = n_agents is None
        self.trace = np.zeros((1 if n_agents is None else int(n_agents), self.trials), dtype=np.float64)

    def _store(self, trial, value):
        v = value.detach().cpu().numpy() if isinstance(value, torch.Tensor) else np.asarray(value)
        self.trace[:, trial] = v.reshape(-1)

    def update(self, logs):
        """Callback for ``on_trial_end``."""
        self._store(int(logs['trial']), logs[self.key])
        return logs

    def from_result(self, res, first_trial=0):
        """Fill the trace from a ``RunResult`` (all trials of one ``train()`` / ``test()`` call)."""
        data = res['trial_steps' if self.key == 'steps' else 'trial_reward'].detach().cpu().numpy()
        self.trace[:, first_trial:first_trial + data.shape[1]] = data
        return self

    def get_trace(self):
        """behavior.py:99-108 / 201-210: the recorded trace (``[trials]``, or ``[N, trials]`` for a batch)."""
        return np.copy(self.trace[0] if self.single else self.trace)


class EscapeLatencyMonitor(_TrialTrace):
    """Steps needed per trial (reference: monitor/behavior.py:16-108; ``logs['steps']`` is the index of
    the last step, behavior.py:83)."""
    key = 'steps'

    def __init__(self, trials, max_steps, n_agents=None):
        super().__init__(trials, n_agents)
        self.max_steps = max_steps


class RewardMonitor(_TrialTrace):
    """Cumulative reward per trial (reference: monitor/behavior.py:111-210)."""
    key = 'trial_reward'

    def __init__(self, trials, reward_range=(0.0, 1.0), n_agents=None):
        super().__init__(trials, n_agents)
        self.reward_range = reward_range


class ResponseMonitor(_TrialTrace):
    """Per-trial response and its cumulative sum (reference: monitor/behavior.py:212-301): the response is
    ``logs['response']`` if present, else ``int(trial_reward > 0)``; ``CRC`` is the running sum."""
    key = 'trial_reward'

    def __init__(self, trials, n_agents=None):
        super().__init__(trials, n_agents)
        self.trace[:] = np.nan                       # the reference starts from NaN-filled traces
        self.CRC = np.full_like(self.trace, np.nan)

    def _store(self, trial, value):
        v = value.detach().cpu().numpy() if isinstance(value, torch.Tensor) else np.asarray(value)
        self.trace[:, trial] = v.reshape(-1)
        self.CRC[:, trial] = np.sum(self.trace[:, :trial + 1], axis=1)

    def update(self, logs):
        if 'response' in logs:
            self._store(int(logs['trial']), logs['response'])
        else:
            r = logs['trial_reward']
            r = r.detach().cpu().numpy() if isinstance(r, torch.Tensor) else np.asarray(r)
            self._store(int(logs['trial']), (r > 0).astype(int))
        return logs

    def from_result(self, res, first_trial=0):
        data = (res['trial_reward'].detach().cpu().numpy() > 0).astype(np.float64)
        for t in range(data.shape[1]):
            self._store(first_trial + t, data[:, t])
        return self

    def get_cumulative(self):
        return np.copy(self.CRC[0] if self.single else self.CRC)


class QMonitor:
    """Q-values of a fixed batch of observations after every trial (reference: monitor/behavior.py:388-464):
    ``update`` appends ``logs['agent'].predict_on_batch(observations)`` -- ``[len(observations), A]`` for a
    single-agent stream, ``[N, len(observations), A]`` for a batch."""

    def __init__(self, trials, observations):
        self.observations = observations
        self.q_trace = []

    def update(self, logs):
        q = logs['agent'].predict_on_batch(self.observations)
        self.q_trace.append(q.detach().cpu().numpy() if isinstance(q, torch.Tensor) else np.asarray(q))
        return logs

    def get_trace(self):
        return self.q_trace


class TrajectoryMonitor:
    """Positions visited in every trial (reference: monitor/behavior.py:304-385, which calls
    ``env.get_position()`` once per step).  On the B200 path the steps of a trial happen inside one launch, so
    the trace is rebuilt from the recorded step buffers of a run (``agent.record = True``): ``from_result`` gives,
    per trial, the ``[T, 2]`` positions after every step -- for one agent (``agent=i``) as the reference's
    list of lists, or for all agents as one NaN-padded ``[N, trials, max_steps, 2]`` tensor on the device
    (the input format of ``analysis.get_occupancy_map``)."""

    def __init__(self, trials, env):
        self.env = env
        self.trajectory_trace = []

    def _positions(self):
        if hasattr(self.env, '_coordinates'):
            return self.env._coordinates
        return self.env._pose[:, :2]

    def from_result(self, res, agent=None):
        assert 'step_next' in res or 'step_sa' in res, 'run the agent with agent.record = True'
        pos = self._positions()
        steps = res['trial_steps'].to(torch.int64) + 1                     # logs['steps'] is the last step's index
        if 'step_next' in res:
            nxt = res['step_next'].to(torch.int64)
        else:                                                              # deterministic world: s' = succ[s, a]
            sa = res['step_sa'].to(torch.int64)
            nxt = self.env._succ.reshape(-1)[sa].to(torch.int64)
        n, trials = steps.shape
        start = torch.cumsum(steps, dim=1) - steps
        tmax = int(steps.max().item()) if steps.numel() else 0
        t = torch.arange(tmax, device=steps.device).view(1, 1, tmax)
        valid = t < steps.unsqueeze(-1)
        idx = (start.unsqueeze(-1) + t).clamp(max=nxt.shape[1] - 1)
        states = torch.gather(nxt.unsqueeze(1).expand(n, trials, nxt.shape[1]), 2, idx)
        out = pos[states]
        out = torch.where(valid.unsqueeze(-1), out, torch.full_like(out, float('nan')))
        if agent is None:
            return out
        self.trajectory_trace = [[out[agent, tr, k].cpu().numpy() for k in range(int(steps[agent, tr]))]
                                 for tr in range(trials)]
        return self.trajectory_trace

    def get_trace(self):
        return self.trajectory_trace
