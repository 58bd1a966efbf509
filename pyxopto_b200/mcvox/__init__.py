"""Voxelised Monte Carlo simulator - mirror of ``xopto.mcvox``."""
from . import mc  # noqa: F401
