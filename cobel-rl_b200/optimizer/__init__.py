"""Model-fitting drivers on top of the batched path (reference: cobel/optimizer/__init__.py)."""
from .grid_search import GridSearchOptimizer  # noqa: F401
from .evolution import EAOptimizer  # noqa: F401
