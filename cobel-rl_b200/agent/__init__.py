"""Tabular agents of the hot path (reference: cobel/agent/__init__.py)."""
from .agent import Agent, Callbacks  # noqa: F401
from .dyna_q import DynaQ  # noqa: F401
from .q import QAgent  # noqa: F401
from .sr import SR  # noqa: F401
from .sfma import SFMA  # noqa: F401
from .pma import PMA  # noqa: F401
from .dyna_dqn import DynaDQN, DynaDSR  # noqa: F401
