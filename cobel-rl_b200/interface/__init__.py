"""Environments of the hot path (reference: cobel/interface/__init__.py)."""
from .interface import Interface  # noqa: F401
from .gridworld import Gridworld, WorldDict  # noqa: F401
from .topology import Topology  # noqa: F401
