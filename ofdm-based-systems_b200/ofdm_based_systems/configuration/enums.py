"""String enums of the simulation schema; values and names follow the reference's
configuration/enums.py:4-67 so that the shipped config/*.json files load unchanged."""
from enum import Enum


class _StrEnum(str, Enum):
    def __str__(self) -> str:
        return self.value


class ConstellationType(_StrEnum):
    QAM = "QAM"
    PSK = "PSK"


class PrefixType(_StrEnum):
    CYCLIC = "CYCLIC"
    ZERO = "ZERO"
    NONE = "NONE"


class EqualizationMethod(_StrEnum):
    ZF = "ZF"
    MMSE = "MMSE"
    NONE = "NONE"


class ModulationType(_StrEnum):
    OFDM = "OFDM"
    SC_OFDM = "SC-OFDM"


class ChannelType(_StrEnum):
    FLAT = "FLAT"      # quirk Q4: means "the built-in 4-tap default", not a flat channel
    CUSTOM = "CUSTOM"


class NoiseType(_StrEnum):
    AWGN = "AWGN"
    NONE = "NONE"


class PowerAllocationType(_StrEnum):
    UNIFORM = "UNIFORM"
    WATERFILLING = "WATERFILLING"


class AdaptiveModulationMode(_StrEnum):
    FIXED = "FIXED"
    CAPACITY_BASED = "CAPACITY_BASED"
