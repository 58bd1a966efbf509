"""Compile-time kernel options (mirror of ``xopto/mcbase/mcoptions.py:119-732``).

Same classes, constructors and class-level instances (``McUseLottery.off``,
``McMethod.ar``, ``McFloatLutMemory.constant_mem`` ...) as the reference, so a
reference script's ``options=[...]`` list works unchanged.  Every option carries
``cl_options = [(NAME, value)]``; the simulator merges the pairs of all plugins,
rejects conflicting duplicates (mcoptions.py:79-85) and maps the result onto the
``XO_*`` defines of the CUDA translation unit.

Options that select an OpenCL memory space or an OpenCL code-generation detail
have no meaning on this engine and are ACCEPTED AND IGNORED (tables are staged in
shared memory whenever they fit, 64-bit atomics are native, structs arrive as
``__grid_constant__`` parameters): ``McIntLutMemory``, ``McFloatLutMemory``,
``McMaterialMemory`` (mcvox), ``McUsePackedStructures``, ``McUseSoft64Atomics``,
``McUseHalfMath``, ``McUseNativeMath`` (throughput mode always uses the MUFU
fast path; deterministic mode never does), ``McUseFluenceCache``, ``McDebugMode``.
"""
from .mcobject import McObject


class McOption(McObject):
    """Base class of kernel options set up through defines (mcoptions.py:119-196)."""

    @staticmethod
    def cl_value(value) -> str:
        if isinstance(value, bool):
            return {False: 'FALSE', True: 'TRUE'}[value]
        if isinstance(value, int):
            return '{:d}'.format(value)
        if isinstance(value, float):
            return 'FP_LITERAL({:.16g})'.format(value)
        if isinstance(value, str):
            return value
        if value is None:
            return ''
        raise TypeError('Option must be of str, bool, int or float type!')

    @classmethod
    def make_define(cls, name, value) -> str:
        if value is None:
            return '#define {}'.format(name)
        return '#define {} {}'.format(name, cls.cl_value(value))

    def __init__(self, name: str, value):
        self.cl_options = [(name, value)]

    name = property(lambda self: self.cl_options[0][0], None, None, 'Option name.')
    value = property(lambda self: self.cl_options[0][1], None, None, 'Option value.')

    def __repr__(self):
        return "{}('{}', {})".format(type(self).__name__, self.name, self.value)

    def __str__(self):
        return self.__repr__()

    def __eq__(self, other):
        return isinstance(other, McOption) and self.cl_options == other.cl_options

    def __hash__(self):
        return hash(tuple(self.cl_options))


class McBoolOption(McOption):
    def __init__(self, name: str, value: bool):
        super().__init__(name, bool(value))


class McIntOption(McOption):
    def __init__(self, name: str, value: int):
        super().__init__(name, int(value))


class McFloatOption(McOption):
    def __init__(self, name: str, value: float):
        super().__init__(name, float(value))


class McTypeOption(McOption):
    """Label-valued option (no quotes in the define)."""

    def __init__(self, name: str, value: str):
        super().__init__(name, str(value))


def _value_repr(self):
    return '{}({})'.format(type(self).__name__, self.value)


class McMethod(McIntOption):
    """Stepping method: 0 albedo weight (default), 1 albedo rejection,
    2 microscopic Beer-Lambert (mcoptions.py:269-318)."""
    albedo_weight = aw = default = McIntOption('MC_METHOD', 0)
    albedo_rejection = ar = McIntOption('MC_METHOD', 1)
    microscopic_beer_lambert = mbl = McIntOption('MC_METHOD', 2)
    _NAMES = {'aw': 0, 'albedo_weight': 0, 'ar': 1, 'albedo_rejection': 1,
              'mbl': 2, 'microscopic_beer_lambert': 2}

    def __init__(self, value: int = 0):
        if isinstance(value, str):
            value = self._NAMES.get(value.lower(), value)
        if value not in (0, 1, 2):
            raise ValueError('Allowed values are 0, 1 or 2!')
        super().__init__('MC_METHOD', value)

    __repr__ = _value_repr


def _bool_option(option_name: str, default: bool, doc: str):
    """Class of a boolean switch with the reference's ``on`` / ``off`` / ``default``
    class-level instances."""
    def __init__(self, value: bool = default):
        McBoolOption.__init__(self, option_name, value)

    cls = type('_', (McBoolOption,), {'__init__': __init__, '__repr__': _value_repr,
                                      '__doc__': doc})
    cls.on = McBoolOption(option_name, True)
    cls.off = McBoolOption(option_name, False)
    cls.default = cls.on if default else cls.off
    return cls


def _named(cls, name):
    cls.__name__ = cls.__qualname__ = name
    cls.__module__ = __name__
    return cls


McUseFluenceCache = _named(_bool_option(
    'MC_USE_FLUENCE_CACHE', False,
    'Fluence cache of the OpenCL kernels (mcoptions.py:321); ignored: fluence grids '
    'are privatised in a shared-memory window by design.'), 'McUseFluenceCache')
McUseHalfMath = _named(_bool_option(
    'MC_USE_HALF_MATH', False,
    'OpenCL half_ math built-ins (mcoptions.py:353); ignored.'), 'McUseHalfMath')
McUseNativeMath = _named(_bool_option(
    'MC_USE_NATIVE_MATH', False,
    'OpenCL native_ math built-ins (mcoptions.py:393); ignored: throughput mode '
    'always uses the MUFU fast path, deterministic mode never does.'), 'McUseNativeMath')
McDebugMode = _named(_bool_option(
    'MC_ENABLE_DEBUG', False, 'Kernel debug printouts (mcoptions.py:516); ignored.'),
    'McDebugMode')
McUseEnhancedRng = _named(_bool_option(
    'MC_USE_ENHANCED_RNG', False,
    'Two MWC steps per uniform draw (mcoptions.py:541).'), 'McUseEnhancedRng')
McUseSoft64Atomics = _named(_bool_option(
    'MC_USE_SOFT_64_ATOMICS', False,
    'Software 64-bit atomics (mcoptions.py:576); ignored: RED.E.ADD.64 is native.'),
    'McUseSoft64Atomics')
McUseLottery = _named(_bool_option(
    'MC_USE_LOTTERY', True, 'Survival lottery at low packet weight (mcoptions.py:600).'),
    'McUseLottery')
McUsePackedStructures = _named(_bool_option(
    'MC_USE_PACKED_STRUCTURES', False,
    'Packed OpenCL structs (mcoptions.py:673); ignored: the plugin structs keep the '
    'natural-alignment layout of the ctypes side.'), 'McUsePackedStructures')
McUseEvents = _named(_bool_option(
    'MC_USE_EVENTS', False,
    'Packet event flags for the trace (mcoptions.py:709).'), 'McUseEvents')
McDeterministic = _named(_bool_option(
    'XO_DETERMINISTIC', False,
    'Engine-specific: deterministic parity mode (static packet schedule, IEEE '
    'arithmetic, portable elementary functions).  Off = throughput mode.'),
    'McDeterministic')


class McMinimumPacketWeight(McFloatOption):
    """Weight below which a packet enters the lottery (mcoptions.py:626)."""
    default = McFloatOption('MC_PACKET_WEIGHT_MIN', 1e-4)

    def __init__(self, value: float = 1e-4):
        super().__init__('MC_PACKET_WEIGHT_MIN', value)

    __repr__ = _value_repr


class McPacketLotteryChance(McFloatOption):
    """Survival probability of the lottery (mcoptions.py:645)."""
    default = McFloatOption('MC_PACKET_LOTTERY_CHANCE', 0.1)

    def __init__(self, value: float = 0.1):
        super().__init__('MC_PACKET_LOTTERY_CHANCE', value)

    __repr__ = _value_repr


_MEMORY = {'global': '__global', '__global': '__global',
           'constant': '__constant', '__constant': '__constant'}


def _memory_option(option_name: str, what: str, default: str, doc: str):
    def __init__(self, value: str = 'global'):
        mem = _MEMORY.get(value)
        if mem is None:
            raise ValueError('{} memory type must be one of "constant" or "global", '
                             'but got "{}"!'.format(what, value))
        McTypeOption.__init__(self, option_name, mem)

    cls = type('_', (McTypeOption,), {'__init__': __init__, '__repr__': _value_repr,
                                      '__doc__': doc})
    cls.constant_mem = McTypeOption(option_name, '__constant')
    cls.global_mem = McTypeOption(option_name, '__global')
    cls.default = cls.constant_mem if default == 'constant' else cls.global_mem
    return cls


McIntLutMemory = _named(_memory_option(
    'MC_INT_LUT_ARRAY_MEMORY', 'Lookup table', 'constant',
    'OpenCL memory space of integer lookup tables (mcoptions.py:424); ignored.'),
    'McIntLutMemory')
McFloatLutMemory = _named(_memory_option(
    'MC_FP_LUT_ARRAY_MEMORY', 'Lookup table', 'constant',
    'OpenCL memory space of float lookup tables (mcoptions.py:470); ignored: the '
    'pool is staged in shared memory whenever it fits (<= 64 KB).'), 'McFloatLutMemory')


def make_defines(options, indent: str = None) -> str:
    """``#define`` lines of a list of (name, value) pairs (mcoptions.py:29-66)."""
    indent = indent or ''
    lines, names = [], []
    for name, value in options:
        if name in names:
            raise ValueError('Option {} defined multiple times!'.format(name))
        names.append(name)
        lines.append('{}{}'.format(indent, McOption.make_define(name, value)))
    return '\n'.join(lines)


def _pairs(item):
    if isinstance(item, (tuple, list)) and len(item) == 2 and isinstance(item[0], str):
        return [tuple(item)]
    co = getattr(item, 'cl_options', None)
    if co is None:
        raise TypeError('Not a kernel option: {!r}'.format(item))
    return list(co() if callable(co) else co)


def resolve_cl_options(*option_lists) -> dict:
    """Merge ``(name, value)`` lists / option objects (of this module or of the
    reference's ``xopto.mcbase.mcoptions``); conflicting duplicates raise
    ValueError (mcoptions.py:68-98)."""
    resolved = {}
    for options in option_lists:
        for item in options or []:
            for name, value in _pairs(item):
                if name in resolved and resolved[name] != value:
                    raise ValueError(
                        'Option {} defined multiple times with different '
                        'values ({} and {})!'.format(name, resolved[name], value))
                resolved[name] = value
    return resolved
