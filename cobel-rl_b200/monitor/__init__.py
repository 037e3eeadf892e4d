"""Behavioural monitors fed from the batched per-trial statistics (reference: cobel/monitor)."""
from .behavior import EscapeLatencyMonitor, RewardMonitor  # noqa: F401
