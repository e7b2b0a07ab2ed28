"""Import-only stand-ins so LearningFilters/models.py (spectral-GNN baselines, off the hot path) imports unmodified."""
from .. import MessagePassing  # noqa: F401
