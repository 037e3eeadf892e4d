"""World builders (reference: cobel/misc/gridworld_tools.py, topology_tools.py)."""
from . import gridworld_tools, topology_tools  # noqa: F401
