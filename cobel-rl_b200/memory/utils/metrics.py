"""State-similarity metrics used by SFMA (reference: memory/utils/metrics.py:10-268).

Computed once on the host with NumPy / LAPACK exactly like the reference (it is a read-only
``D[S,S]`` input of the replay kernel, uploaded by ``SFMAMemory``).
"""
import abc

import numpy as np


class Metric(abc.ABC):
    def __init__(self):
        self.D = None

    @abc.abstractmethod
    def update_transitions(self):
        ...


def _uniform_policy_transitions(sas):
    """State-to-state transition matrix under a uniform policy (``np.sum(sas, axis=1) / A``)."""
    sas = np.asarray(sas)
    return np.sum(sas, axis=1) / sas.shape[1]


class Euclidean(Metric):
    """``D = exp(-euclidean distance)`` between grid cells (metrics.py:28-59)."""

    def __init__(self, width, height):
        super().__init__()
        idx = np.arange(width * height)
        rc = np.stack(np.divmod(idx, width), axis=1).astype(float)
        self.D = np.exp(-np.sqrt(((rc[:, None, :] - rc[None, :, :]) ** 2).sum(axis=2)))

    def update_transitions(self):
        pass


class SR(Metric):
    """Successor representation of the uniform-policy random walk (metrics.py:62-111)."""

    def __init__(self, sas, gamma):
        super().__init__()
        self.sas, self.gamma = sas, gamma
        self.update_transitions()

    def update_transitions(self):
        T = _uniform_policy_transitions(self.sas)
        self.D = np.linalg.inv(np.eye(T.shape[0]) - self.gamma * T)


class DR(Metric):
    """Default representation: open-field SR ``D0`` corrected for the walls by a low-rank
    (Woodbury) update ``D = D0 - B`` restricted to the states with blocked moves
    (metrics.py:114-268)."""

    def __init__(self, width, height, sas, gamma, invalid_transitions, T_default=None):
        super().__init__()
        self.width, self.height, self.nb_states = width, height, width * height
        self.sas, self.gamma, self.invalid_transitions = sas, gamma, invalid_transitions
        if T_default is None:
            self.build_default_transition_matrix()
        else:
            self.T_default = T_default
        self.D0 = np.linalg.inv(np.eye(self.nb_states) - self.gamma * self.T_default)
        self.update_transitions()

    def update_transitions(self):
        self.T_new = _uniform_policy_transitions(self.sas)
        self.B = np.zeros(self.T_new.shape)
        if len(self.invalid_transitions) > 0:
            self.states = np.unique(np.array(self.invalid_transitions)[:, 0])
            L = np.eye(self.nb_states) - self.gamma * self.T_new
            L0 = np.eye(self.nb_states) - self.gamma * self.T_default
            delta = L[self.states] - L0[self.states]
            alpha = np.linalg.inv(np.eye(self.states.shape[0]) + np.matmul(delta, self.D0[:, self.states]))
            self.B = np.matmul(np.matmul(self.D0[:, self.states], alpha), np.matmul(delta, self.D0))
        self.D = self.D0 - self.B

    def build_default_transition_matrix(self):
        """Open-field random walk: each of the 4 moves with probability 1/4, clipped at the border."""
        S, W, H = self.nb_states, self.width, self.height
        idx = np.arange(S)
        row, col = idx // W, idx % W
        T = np.zeros((S, S))
        for dr, dc in ((0, -1), (-1, 0), (0, 1), (1, 0)):
            nxt = np.clip(row + dr, 0, H - 1) * W + np.clip(col + dc, 0, W - 1)
            np.add.at(T, (idx, nxt), 0.25)
        self.T_default = T
