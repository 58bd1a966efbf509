"""Layered (planar) Monte Carlo simulator - mirror of ``xopto.mcml``."""
from . import mc  # noqa: F401
