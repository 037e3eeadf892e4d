"""Memory modules of the hot path (reference: cobel/memory/__init__.py)."""
from .dyna_q import DynaQMemory  # noqa: F401
from .sfma import SFMAMemory  # noqa: F401
from .pma import PMAMemory  # noqa: F401
