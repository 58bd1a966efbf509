"""Layered (planar) Monte Carlo simulator - mirror of ``xopto.mcml``."""
from . import mc  # noqa: F401


# ---- the reference's subpackage layout -------------------------------------------
# xopto.mcml exposes its plugin families as subpackages (xopto.mcml.mcoptions,
# .mcpf, .mcfluence, .mctrace, .mcutil.fiber ...).  The same import statements work
# here: the shared modules of pyxopto_b200.mcbase / pyxopto_b200.cl are registered
# under this package's name.
import sys as _sys
from ..cl import clinfo, clrng, cltypes                      # noqa: E402,F401
from ..mcbase import (mcobject, mcoptions, mctypes, mcpf, mcfluence, mctrace,  # noqa: E402,F401
                      mcsv, mcprogress, mcmaterial, mcutil)

for _name in ('clinfo', 'clrng', 'cltypes', 'mcobject', 'mcoptions', 'mctypes', 'mcpf', 'mcfluence', 'mctrace', 'mcsv', 'mcprogress', 'mcmaterial', 'mcutil'):
    _sys.modules.setdefault(__name__ + '.' + _name, globals()[_name])
for _name in ('axis', 'boundary', 'buffer', 'fiber', 'geometry', 'lut'):
    _sys.modules.setdefault(__name__ + '.mcutil.' + _name, getattr(mcutil, _name))
del _name

# xopto.mcml.mcrun: batch runners (RunMinWeight* / RunMinPacketsTrace) of this geometry
from ..mcbase import mcrun as _mcrun                        # noqa: E402
mcrun = _mcrun.geometry_module(__name__, mc, ('top', 'bottom', 'specular'))
from . import mcrun_helper as _helper                        # noqa: E402
mcrun.McRunHelper, mcrun.helper = _helper.McRunHelper, _helper
_sys.modules.setdefault(__name__ + '.mcrun.helper', _helper)
