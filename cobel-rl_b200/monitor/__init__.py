"""Behavioural monitors fed from the batched per-trial statistics (reference: cobel/monitor)."""
from .behavior import (EscapeLatencyMonitor, QMonitor, ResponseMonitor, RewardMonitor,  # noqa: F401
                       TrajectoryMonitor)
