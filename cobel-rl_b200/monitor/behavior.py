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
