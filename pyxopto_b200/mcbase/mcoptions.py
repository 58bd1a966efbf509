"""Compile-time kernel options (mirror of ``xopto/mcbase/mcoptions.py:269-732``).

Every option resolves to ``(NAME, value)`` pairs; the simulator merges the pairs
of all plugins, rejects conflicting duplicates (mcoptions.py:79-85) and maps the
result onto the ``XO_*`` defines of the CUDA translation unit.
"""


class McOption:
    name = None

    def __init__(self, value):
        self._value = value

    @property
    def value(self):
        return self._value

    def cl_options(self, mc=None):
        return [(self.name, self._value)]

    def __repr__(self):
        return '{}({})'.format(type(self).__name__, self._value)

    def __eq__(self, other):
        return type(self) is type(other) and self._value == other._value

    def __hash__(self):
        return hash((type(self).__name__, self._value))


class McBoolOption(McOption):
    def __init__(self, value: bool):
        super().__init__(bool(value))


def _switch(cls):
    cls.on = cls(True)
    cls.off = cls(False)
    cls.default = cls(cls.default_value)
    return cls


class McMethod(McOption):
    """Stepping method: albedo weight (default), albedo rejection, microscopic Beer-Lambert."""
    name = 'MC_METHOD'
    _NAMES = {'aw': 0, 'albedo_weight': 0, 'ar': 1, 'albedo_rejection': 1,
              'mbl': 2, 'microscopic_beer_lambert': 2}

    def __init__(self, value='albedo_weight'):
        if isinstance(value, str):
            value = self._NAMES[value.lower()]
        super().__init__(int(value))


McMethod.albedo_weight = McMethod.aw = McMethod.default = McMethod(0)
McMethod.albedo_rejection = McMethod.ar = McMethod(1)
McMethod.microscopic_beer_lambert = McMethod.mbl = McMethod(2)


@_switch
class McUseNativeMath(McBoolOption):
    name = 'MC_USE_NATIVE_MATH'
    default_value = False


@_switch
class McUseLottery(McBoolOption):
    name = 'MC_USE_LOTTERY'
    default_value = True


@_switch
class McUseFluenceCache(McBoolOption):
    name = 'MC_USE_FLUENCE_CACHE'
    default_value = False


@_switch
class McDebugMode(McBoolOption):
    name = 'MC_ENABLE_DEBUG'
    default_value = False


@_switch
class McUseEnhancedRng(McBoolOption):
    name = 'MC_USE_ENHANCED_RNG'
    default_value = False


@_switch
class McUseSoft64Atomics(McBoolOption):
    name = 'MC_USE_SOFT_64_ATOMICS'
    default_value = False


@_switch
class McDeterministic(McBoolOption):
    """Engine-specific: deterministic parity mode (static packet schedule, IEEE
    arithmetic, portable elementary functions).  Off = throughput mode."""
    name = 'XO_DETERMINISTIC'
    default_value = False


class McMinimumPacketWeight(McOption):
    name = 'MC_PACKET_WEIGHT_MIN'

    def __init__(self, value: float = 1e-4):
        super().__init__(float(value))


McMinimumPacketWeight.default = McMinimumPacketWeight(1e-4)


class McPacketLotteryChance(McOption):
    name = 'MC_PACKET_LOTTERY_CHANCE'

    def __init__(self, value: float = 0.1):
        super().__init__(float(value))


McPacketLotteryChance.default = McPacketLotteryChance(0.1)


class McFloatLutMemory(McOption):
    """Where float lookup tables live: 'global' | 'constant' (reference names);
    this engine stages them in shared memory whenever they fit."""
    name = 'MC_FP_LUT_MEMORY'

    def __init__(self, value='global'):
        super().__init__(str(value))


McFloatLutMemory.global_mem = McFloatLutMemory.default = McFloatLutMemory('global')
McFloatLutMemory.constant_mem = McFloatLutMemory('constant')


def resolve_cl_options(*option_lists) -> dict:
    """Merge ``(name, value)`` lists; conflicting duplicates raise ValueError."""
    resolved = {}
    for options in option_lists:
        for item in options or []:
            if isinstance(item, McOption):
                pairs = item.cl_options()
            else:
                pairs = [item]
            for name, value in pairs:
                if name in resolved and resolved[name] != value:
                    raise ValueError(
                        'Option {} defined multiple times with different '
                        'values ({} and {})!'.format(name, resolved[name], value))
                resolved[name] = value
    return resolved
