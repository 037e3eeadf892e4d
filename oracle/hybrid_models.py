"""The two-layer fp64 networks of the reference's own Dyna-DQN / Dyna-DSR unit tests (unit_tests/test_dyna_dqn.py:19-33,
unit_tests/test_dyna_dsr.py:18-31), used by the golden generator and by the GPU parity tests (test infrastructure)."""
import torch


class Model(torch.nn.Module):
    def __init__(self, units_out, units_in=25, hidden=32):
        super().__init__()
        self.hidden = torch.nn.Linear(units_in, hidden)
        self.output = torch.nn.Linear(hidden, units_out)
        self.double()

    def forward(self, batch):
        return self.output(torch.nn.functional.relu(self.hidden(batch)))


def seeded(units_out, seed):
    """A Model with deterministic initial weights (torch's CPU generator)."""
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = Model(units_out)
    torch.random.set_rng_state(g)
    return m
