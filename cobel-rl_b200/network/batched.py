"""N independent copies of a PyTorch module, evaluated and trained together (reference: network/network.py:12-150,
network/network_torch.py:31-452).

The reference's Dyna hybrids (agent/dyna_q.py:333-1150) own one ``TorchNetwork`` per agent.  On the batched path the
N agents of a ``BatchStream`` each need their own weights, so this class stacks the parameters of N copies of the
user's ``torch.nn.Module`` along a leading agent axis and evaluates them with ``torch.func.vmap`` -- same interface
(``predict_on_batch``, ``train_on_batch``, ``get_weights``, ``set_weights``, ``clone``), every batch carrying a
leading agent axis: ``[N, B, ...]``.  ``train_on_batch`` takes an optional ``active[N]`` mask: agents outside it keep
weights AND optimizer state (a finished trial must not advance its Adam moments).

Loss and optimizer follow ``TorchNetwork``'s defaults (network_torch.py:223-330): mean-squared error with 'mean'
reduction per agent and Adam(lr=1e-3, betas=(0.9, 0.999), eps=1e-8), written out with the operations of
``torch.optim.Adam`` (lerp / addcmul / addcdiv) per element, so that every agent's trajectory is that of its own
``torch.optim.Adam``; 'sgd' is the plain ``p -= lr * grad``.
"""
import copy

import torch
from torch.func import functional_call, stack_module_state, vmap


class BatchedTorchNetwork:
    def __init__(self, models, optimizer='adam', optimizer_params=None, loss='mse', device='cuda'):
        """``models``: a list of N ``torch.nn.Module`` with identical structure (one per agent)."""
        models = list(models)
        assert len(models) >= 1
        self.n_agents = len(models)
        self.device = torch.device(device)
        self._base = copy.deepcopy(models[0]).to('meta')
        params, buffers = stack_module_state([m.to(self.device) for m in models])
        self.params = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
        self.buffers = {k: v.detach().clone() for k, v in buffers.items()}
        assert loss in ('mse', 'mean_squared_error'), 'only the mean-squared error is built'
        assert optimizer in ('adam', 'Adam', 'sgd', 'SGD')
        self.optimizer = optimizer.lower()
        op = dict(optimizer_params or {})
        self.lr = op.pop('lr', 1e-3 if self.optimizer == 'adam' else 1e-2)
        self.betas = op.pop('betas', (0.9, 0.999))
        self.eps = op.pop('eps', 1e-8)
        assert not op, 'unsupported optimizer parameters %s' % sorted(op)
        self._m = {k: torch.zeros_like(v) for k, v in self.params.items()}
        self._v = {k: torch.zeros_like(v) for k, v in self.params.items()}
        self._t = torch.zeros(self.n_agents, dtype=torch.float64, device=self.device)
        self._call = vmap(lambda p, b, x: functional_call(self._base, (p, b), (x,)), in_dims=(0, 0, 0))

    # -- Network interface --------------------------------------------------------------------------------
    def predict_on_batch(self, batch):
        """network_torch.py:110-130 for every agent: ``[N, B, ...] -> [N, B, out]`` (device tensor)."""
        with torch.no_grad():
            return self._call(self.params, self.buffers, self._as_input(batch))

    def train_on_batch(self, batch, targets, active=None, sample_mask=None):
        """network_torch.py:132-167 for every agent in ``active`` (default: all).  ``sample_mask[N, B]`` restricts
        an agent's batch to the marked samples (the reference trains on sub-batches of varying size, e.g. the
        experiences of one action, agent/dyna_q.py:1101-1131); agents without a marked sample are skipped."""
        x, y = self._as_input(batch), self._as_input(targets)
        for p in self.params.values():
            p.grad = None
        pred = self._call(self.params, self.buffers, x)
        if y.dim() == 2:
            y = y.unsqueeze(-1)
        act = torch.ones(self.n_agents, dtype=torch.bool, device=self.device) if active is None else active.to(self.device).bool()
        if sample_mask is None:
            per_agent = ((pred - y) ** 2).reshape(self.n_agents, -1).mean(dim=1)  # MSELoss(reduction='mean') per agent
        else:
            w = sample_mask.to(self.device).to(pred.dtype)
            cnt = w.sum(dim=1)
            act = act & (cnt > 0)
            err = ((pred - y) ** 2).reshape(self.n_agents, w.shape[1], -1)
            per_agent = (err * w.unsqueeze(-1)).sum(dim=(1, 2)) / (cnt.clamp(min=1.0) * err.shape[2])
        per_agent.sum().backward()
        with torch.no_grad():
            self._t += act.to(self._t.dtype)
            t = self._t.clamp(min=1.0)
            for k, p in self.params.items():
                g = p.grad
                sel = act.reshape((-1,) + (1,) * (p.dim() - 1))
                if self.optimizer == 'sgd':
                    p.copy_(torch.where(sel, p - self.lr * g, p))
                    continue
                b1, b2 = self.betas
                m = torch.lerp(self._m[k], g, 1 - b1)
                v = (self._v[k] * b2).addcmul_(g, g, value=1 - b2)
                shape = (-1,) + (1,) * (p.dim() - 1)
                bc1 = (1 - b1 ** t).reshape(shape).to(p.dtype)
                bc2s = (1 - b2 ** t).sqrt().reshape(shape).to(p.dtype)
                denom = (v.sqrt() / bc2s).add_(self.eps)
                new = p - (self.lr / bc1) * (m / denom)
                p.copy_(torch.where(sel, new, p))
                self._m[k] = torch.where(sel, m, self._m[k])
                self._v[k] = torch.where(sel, v, self._v[k])
            for p in self.params.values():
                p.grad = None

    def get_weights(self):
        """network_torch.py:169-183: the state-dict entries in order, each with a leading agent axis."""
        return [v.detach().clone() for v in list(self.params.values()) + list(self.buffers.values())]

    def set_weights(self, weights, active=None):
        """network_torch.py:185-199; ``active``: only those agents take the new weights."""
        dst = list(self.params.values()) + list(self.buffers.values())
        assert len(weights) == len(dst)
        with torch.no_grad():
            for d, w in zip(dst, weights):
                w = torch.as_tensor(w, device=self.device, dtype=d.dtype)
                if active is None:
                    d.copy_(w)
                else:
                    d.copy_(torch.where(active.to(self.device).bool().reshape((-1,) + (1,) * (d.dim() - 1)), w, d))

    def clone(self):
        """network_torch.py:201-221: same weights, same optimizer state, independent storage."""
        other = copy.copy(self)
        other.params = {k: v.detach().clone().requires_grad_(True) for k, v in self.params.items()}
        other.buffers = {k: v.detach().clone() for k, v in self.buffers.items()}
        other._m = {k: v.clone() for k, v in self._m.items()}
        other._v = {k: v.clone() for k, v in self._v.items()}
        other._t = self._t.clone()
        other._call = vmap(lambda p, b, x: functional_call(other._base, (p, b), (x,)), in_dims=(0, 0, 0))
        return other

    # -- helpers -----------------------------------------------------------------------------------------------
    def _as_input(self, batch):
        ref = next(iter(self.params.values()))
        x = torch.as_tensor(batch, device=self.device)
        if x.is_floating_point() and x.dtype != ref.dtype:
            x = x.to(ref.dtype)
        assert x.shape[0] == self.n_agents, 'batches carry a leading agent axis'
        return x
