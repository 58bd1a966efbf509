"""Concentric-cylinder Monte Carlo simulator - mirror of ``xopto.mccyl``."""
from . import mc  # noqa: F401
