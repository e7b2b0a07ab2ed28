"""Import-only stand-in (LearningFilters/models.py:13); the baselines that call it are off the hot path."""


def gcn_norm(*a, **k):
    raise NotImplementedError("off the SignNet hot path")
