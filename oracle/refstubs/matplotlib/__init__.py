"""Inert stand-in for matplotlib (imported by reference viewers, never used)."""
