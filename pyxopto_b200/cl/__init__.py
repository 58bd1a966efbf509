"""Device access layer (mirror of ``xopto.cl``): device discovery (``clinfo``),
MWC seed generation (``clrng``) and ctypes struct helpers (``cltypes``)."""
from . import clinfo, clrng, cltypes  # noqa: F401
