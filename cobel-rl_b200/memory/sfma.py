"""SFMA memory (reference: memory/sfma.py:21-416).

Per-agent experience tables plus the replay-relevant structures: strengths ``C`` and recency
``T`` (flat index ``a*S + s``) and inhibition ``I``.  The replay itself
(priority = C * D * (1 - I), softmax sampling, inhibition) runs inside ``SFMA.train()``
(csrc/sfma.cu); this object owns the tensors and the tunables, with the reference's names.
"""
import torch

from .dyna_q import TableMemory

MODES = ('default', 'forward', 'reverse', 'blend_forward', 'blend_reverse', 'interpolate', 'sweeping')


class SFMAMemory(TableMemory):
    def __init__(self, metric, nb_states, nb_actions, decay_inhibition=0.9, decay_strength=1.0,
                 learning_rate=0.9, rng=None):
        self.metric = metric
        self.nb_states, self.nb_actions = int(nb_states), int(nb_actions)
        self.decay_inhibition = decay_inhibition
        self.decay_strength = decay_strength
        self.decay_recency = 0.9
        self.beta = 20
        self.C_step = 1.0
        self.I_step = 1.0
        self.R_threshold = 10.0 ** -6
        self.deterministic = False
        self.recency = False
        self.mode = 'default'
        self.blend = 0.1
        self.interpolation_fwd, self.interpolation_rev = 0.5, 0.5
        # options of the reference that the B200 path does not implement (must stay at their defaults)
        self.C_normalize = False
        self.D_normalize = False
        self.R_normalize = True
        self.reward_mod_local = self.error_mod_local = self.reward_mod = self.error_mod = False
        self.policy_mod = self.state_mod = False
        self.reward_modulation = 1.0
        self._D_dev = None
        super().__init__(nb_states, nb_actions, learning_rate, rng, init_self_loops=True)

    def _allocate_extra(self, stream):
        n, S, A, dev = stream.n_agents, self.nb_states, self.nb_actions, stream.device
        self._C = torch.zeros((n, S * A), dtype=torch.float64, device=dev)
        self._T = torch.zeros((n, S * A), dtype=torch.float64, device=dev)
        self._I = torch.zeros((n, S), dtype=torch.float64, device=dev)

    C = property(lambda self: self._view(self._C))
    T = property(lambda self: self._view(self._T))
    I = property(lambda self: self._view(self._I))      # noqa: E741

    def similarity(self):
        """``metric.D`` as a device tensor (re-uploaded when the metric object changed it)."""
        D = self.metric.D
        if self._D_dev is None or self._D_src is not D:
            self._D_dev = torch.as_tensor(D, dtype=torch.float64).contiguous().to(self._alloc_for.device)
            self._D_src = D
        assert self._D_dev.shape == (self.nb_states, self.nb_states)
        return self._D_dev

    def check_supported(self):
        # error_mod / error_mod_local read experience['td'], which the reference's SFMA.train() has not computed
        # when it calls store() (agent/sfma.py:289-291) -- they cannot run there either
        unsupported = [k for k in ('error_mod_local', 'error_mod') if getattr(self, k)]
        if unsupported:
            raise NotImplementedError('SFMAMemory options not implemented by the B200 path: %s' % unsupported)

    def mod_flags(self):
        """COBEL_SFMA_MOD_* / *_NORMALIZE bits (include/cobel_b200.h), memory/sfma.py:216-236, 283-288, 319-320."""
        return ((1 if self.reward_mod_local else 0) | (2 if self.reward_mod else 0) | (4 if self.state_mod else 0) |
                (8 if self.C_normalize else 0) | (16 if self.D_normalize else 0) | (0 if self.R_normalize else 32))

    def mode_id(self):
        # an unknown mode string behaves like 'default' in the reference (memory/sfma.py:289-306)
        return MODES.index(self.mode) if self.mode in MODES else 0
