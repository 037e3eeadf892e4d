"""Function approximators on the agent axis (reference: cobel/network/__init__.py)."""
from .batched import BatchedTorchNetwork  # noqa: F401
