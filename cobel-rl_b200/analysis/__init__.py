"""Analysis helpers on the batched trajectory buffers (reference: cobel/analysis)."""
from .behavior_spatial import get_occupancy_map, match  # noqa: F401
