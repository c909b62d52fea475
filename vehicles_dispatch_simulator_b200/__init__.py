"""B200-native batched vehicle-dispatch environment (hot path of
szlhl1040/Vehicles-Dispatch-Simulator's per-time-slot loop)."""
from ._native import PARITY_THRESHOLD, STAT_NAMES, VdsError  # noqa: F401

__all__ = ["PARITY_THRESHOLD", "STAT_NAMES", "VdsError"]
