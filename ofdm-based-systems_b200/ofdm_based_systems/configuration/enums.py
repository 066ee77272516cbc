"""String enums of the simulation schema.  Member names and values are those the shipped config/*.json files and
the reference's callers use (its configuration/enums.py); ``str(member)`` is the JSON value."""
import enum


def _json_enum(name: str, **members: str):
    cls = enum.Enum(name, members, type=str, module=__name__)
    cls.__str__ = lambda self: self.value
    return cls


ConstellationType = _json_enum("ConstellationType", QAM="QAM", PSK="PSK")
PrefixType = _json_enum("PrefixType", CYCLIC="CYCLIC", ZERO="ZERO", NONE="NONE")
EqualizationMethod = _json_enum("EqualizationMethod", ZF="ZF", MMSE="MMSE", NONE="NONE")
ModulationType = _json_enum("ModulationType", OFDM="OFDM", SC_OFDM="SC-OFDM")
ChannelType = _json_enum("ChannelType", FLAT="FLAT", CUSTOM="CUSTOM")   # FLAT = the built-in 4-tap default (quirk Q4)
NoiseType = _json_enum("NoiseType", AWGN="AWGN", NONE="NONE")
PowerAllocationType = _json_enum("PowerAllocationType", UNIFORM="UNIFORM", WATERFILLING="WATERFILLING")
AdaptiveModulationMode = _json_enum("AdaptiveModulationMode", FIXED="FIXED", CAPACITY_BASED="CAPACITY_BASED")
