"""Shared Monte Carlo core (mirror of ``xopto.mcbase`` for the hot path)."""
