"""Host utilities shared by the plugins (axis, boundary, geometry, fiber, buffers, LUTs)."""
from . import axis, boundary, geometry, fiber, buffer, lut  # noqa: F401
