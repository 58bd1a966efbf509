"""Plugin objects of the reference package used directly with this engine.

A script written against ``xopto`` builds its layers / materials, sources,
detectors, fluence, trace and surface-layout objects with the reference's own
classes.  Those objects carry OpenCL-C source fragments, not the hand-written CUDA
structs of this engine, so they cannot be bound to the kernel as they are - but
every one of them describes itself completely with ``todict()`` (the reference's
own serialisation protocol, mcobject / ``fromdict`` in every plugin module), and
the classes of this package take the same constructor arguments.  ``adopt``
rebuilds an object tree from that description with the classes of the matching
``pyxopto_b200`` geometry package; the packed structs of the result are
byte-identical to the reference's (tests/test_reference_objects.py, run where the
reference is installed).

The simulator constructors call this themselves: ``pyxopto_b200.mcml.mc.Mc(layers,
source, detectors, ...)`` accepts ``xopto`` objects in any argument position.
"""
import importlib
import inspect

import numpy as np


def _is_native(obj) -> bool:
    """True for everything that is not an object of the reference package: this
    package's own classes and user-written plugin classes (which carry their own
    kernel fragments and are bound as they are, mcsim._user_fragments)."""
    return type(obj).__module__.split('.')[0] != 'xopto'


def _search_path(geometry: str, context=None):
    """Modules searched for a class name: the module of the parent object first
    (``SixAroundOne`` is a detector inside ``Detectors`` and a probe layout inside
    ``SurfaceLayouts``), then the plugin modules of the geometry."""
    pkg = __name__.rsplit('.', 1)[0]
    names = {
        'mcml': ['mcml.mcsource', 'mcml.mcdetector', 'mcml.mcsurface', 'mcml.mclayer'],
        'mcvox': ['mcvox.mcsource', 'mcml.mcdetector', 'mcvox.mcgeometry', 'mcbase.mcmaterial'],
        'mccyl': ['mccyl.mcsource', 'mccyl.mcdetector', 'mccyl.mclayer'],
    }[geometry] + ['mcbase.mcpf', 'mcbase.mcfluence', 'mcbase.mctrace', 'mcbase.mcsv',
                   'mcbase.mcutil.fiber', 'mcbase.mcutil.axis', 'mcbase.mcutil.lut']
    mods = [importlib.import_module(pkg + '.' + n) for n in names]
    if context is not None:
        mods.insert(0, context)
    return mods


def _resolve(type_name: str, geometry: str, context=None):
    for mod in _search_path(geometry, context):
        cls = getattr(mod, type_name, None)
        if isinstance(cls, type):
            return cls, mod
    raise TypeError('pyxopto_b200 has no counterpart of the reference plugin class '
                    '"{}" in the {} geometry.'.format(type_name, geometry))


def _child(src, key):
    """The attribute of the described object that a dictionary entry came from."""
    if src is None:
        return None
    try:
        return getattr(src, key, None)
    except Exception:
        return None


def _build(desc, geometry: str, context=None, src=None):
    """Object tree from a ``todict()`` description; ``src`` is the described object
    where known (constructor arguments its ``todict()`` forgets are read from its
    public attributes, e.g. the ``spacing`` of a SixAroundOne surface layout)."""
    if isinstance(desc, dict) and 'type' in desc:
        cls, mod = _resolve(desc['type'], geometry, context)
        kwargs = {k: _build(v, geometry, mod, _child(src, k))
                  for k, v in desc.items() if k != 'type'}
        # serialised names that differ from the constructor's only by an underscore
        # ('r_axis' -> raxis, as the reference's own fromdict() methods map them)
        params = inspect.signature(cls.__init__).parameters
        if not any(p.kind is p.VAR_KEYWORD for p in params.values()):
            kwargs = {(k if k in params or k.replace('_', '') not in params
                       else k.replace('_', '')): v for k, v in kwargs.items()}
        if src is not None and not _is_native(src):
            for name, p in params.items():
                if name == 'self' or name in kwargs or p.kind in (p.VAR_KEYWORD, p.VAR_POSITIONAL):
                    continue
                value = _child(src, name)
                if isinstance(value, (bool, int, float, str, np.floating, np.integer)):
                    kwargs[name] = value
        return cls(**kwargs)
    if isinstance(desc, (list, tuple)) and any(isinstance(v, dict) and 'type' in v for v in desc):
        out = []
        for i, v in enumerate(desc):
            try:
                item = src[i] if src is not None else None
            except Exception:
                item = None
            out.append(_build(v, geometry, context, item))
        return out
    return desc


def adopt(obj, geometry: str):
    """``obj`` if it already belongs to this package (or is None), else its
    counterpart built from ``obj.todict()`` with the classes of ``geometry``
    ('mcml', 'mcvox' or 'mccyl')."""
    if obj is None or _is_native(obj) or not hasattr(obj, 'todict'):
        return obj
    try:
        desc = obj.todict()
    except AttributeError:
        # (FluenceCyl / FluenceCylt: todict() of the reference reads attributes the
        # class does not have, fluencecyl.py; described from the public properties)
        cls, _ = _resolve(type(obj).__name__, geometry)
        params = [p for p in inspect.signature(cls.__init__).parameters if p != 'self']
        desc = {p: getattr(obj, p) for p in params if hasattr(obj, p)}
        desc = {k: (v.todict() if hasattr(v, 'todict') and not _is_native(v) else v)
                for k, v in desc.items()}
        desc['type'] = type(obj).__name__
    new = _build(desc, geometry, None, obj)
    # state that todict() does not carry
    if hasattr(obj, 'material') and hasattr(new, 'material') and \
            isinstance(getattr(obj, 'material', None), np.ndarray):
        new.material[:] = obj.material              # voxel -> material index array
    return new
