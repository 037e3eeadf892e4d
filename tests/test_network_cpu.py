"""CPU: ``network.BatchedTorchNetwork`` (N independent copies of a module under torch.func.vmap, the function
approximator of the Dyna hybrids) against N separate reference-style networks: ``torch.nn.MSELoss`` +
``torch.optim.Adam`` / ``SGD`` per copy (network/network_torch.py:110-167, 223-330)."""
import copy

import pytest
import torch

from cobel_rl_b200.network import BatchedTorchNetwork


class Model(torch.nn.Module):
    def __init__(self, n_in=5, n_out=3):
        super().__init__()
        self.hidden = torch.nn.Linear(n_in, 7)
        self.output = torch.nn.Linear(7, n_out)
        self.double()

    def forward(self, x):
        return self.output(torch.relu(self.hidden(x)))


def _setup(n=3, seed=0):
    torch.manual_seed(seed)
    models = [Model() for _ in range(n)]
    refs = [copy.deepcopy(m) for m in models]
    x = torch.randn(n, 6, 5, dtype=torch.float64)
    y = torch.randn(n, 6, 3, dtype=torch.float64)
    return models, refs, x, y


@pytest.mark.parametrize('opt', ['adam', 'sgd'])
def test_training_equals_separate_optimizers_with_masks(opt):
    models, refs, x, y = _setup()
    net = BatchedTorchNetwork(models, optimizer=opt, device='cpu')
    ropt = [(torch.optim.Adam(r.parameters()) if opt == 'adam' else torch.optim.SGD(r.parameters(), lr=1e-2)) for r in refs]
    active = torch.tensor([True, False, True])
    for it in range(6):
        act = active if it % 2 else None
        net.train_on_batch(x, y, active=act)
        for n, (r, o) in enumerate(zip(refs, ropt)):
            if act is not None and not bool(act[n]):
                continue                                   # an inactive agent keeps weights AND optimizer state
            o.zero_grad()
            torch.nn.MSELoss()(r(x[n]), y[n]).mean().backward()
            o.step()
    pred = net.predict_on_batch(x)
    for n, r in enumerate(refs):
        assert float((pred[n] - r(x[n])).abs().max()) < 1e-14


def test_sample_mask_trains_on_sub_batches():
    """Per-agent sub-batches of varying size (the per-action batches of DynaDSR.replay, agent/dyna_q.py:1101-1131);
    an agent without a marked sample is left untouched."""
    models, refs, x, y = _setup()
    net = BatchedTorchNetwork(models, device='cpu')
    ropt = [torch.optim.Adam(r.parameters()) for r in refs]
    mask = torch.tensor([[1, 0, 1, 1, 0, 0], [0, 0, 0, 0, 0, 0], [1, 1, 1, 1, 1, 1]], dtype=torch.bool)
    for _ in range(3):
        net.train_on_batch(x, y, sample_mask=mask)
        for n, (r, o) in enumerate(zip(refs, ropt)):
            if not bool(mask[n].any()):
                continue
            o.zero_grad()
            torch.nn.MSELoss()(r(x[n][mask[n]]), y[n][mask[n]]).mean().backward()
            o.step()
    pred = net.predict_on_batch(x)
    for n, r in enumerate(refs):
        assert float((pred[n] - r(x[n])).abs().max()) < 1e-14


def test_clone_weights_and_one_dimensional_targets():
    models, refs, x, y = _setup()
    net = BatchedTorchNetwork([Model(5, 1) for _ in range(3)], device='cpu')
    twin = net.clone()
    assert all(torch.equal(a, b) for a, b in zip(net.get_weights(), twin.get_weights()))
    twin.train_on_batch(x, y[:, :, 0])                      # [N, B] targets are reshaped to [N, B, 1] (network_torch.py:150-155)
    assert not torch.equal(net.get_weights()[0], twin.get_weights()[0])      # independent storage
    # soft target update for a subset of agents (agent/dyna_q.py:690-700)
    act = torch.tensor([True, False, False])
    before = [w.clone() for w in net.get_weights()]
    net.set_weights([t + 0.5 * (o - t) for t, o in zip(net.get_weights(), twin.get_weights())], active=act)
    after = net.get_weights()
    assert not torch.equal(after[0][0], before[0][0]) and torch.equal(after[0][1:], before[0][1:])
    with pytest.raises(AssertionError):
        net.predict_on_batch(x[:2])                        # batches carry the agent axis
