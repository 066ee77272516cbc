"""Settings schema of config/settings.json and config/simulation_settings*.json.

The field names, defaults and validation messages are the contract of the shipped JSON files and of the reference's
callers (its configuration/models.py); the models themselves are generated from the tables below with
``pydantic.create_model``.  Unknown JSON keys such as ``capacity_scaling_factor`` are ignored (pydantic's default).
"""
import json
import os
from typing import Optional

from pydantic import BaseModel, create_model, field_validator

from ofdm_based_systems.configuration import enums as E


class BaseSettings(BaseModel):
    @classmethod
    def from_json(cls, file_path: str):
        if not os.path.exists(file_path):
            raise FileNotFoundError(f"Configuration file not found: {file_path}")
        with open(file_path, "r", encoding="utf-8") as fh:
            return cls(**json.load(fh))

    # (label, attribute, quote the value, only when not None) rows printed by __str__
    _report = ()

    def __str__(self):
        rows = []
        for label, attr, quoted, optional in self._report:
            value = getattr(self, attr)
            if optional and value is None:
                continue
            rows.append(f"{label}: '{value}'" if quoted else f"{label}: {value}")
        return "\n".join(rows)


def _one_of_bits_or_symbols(cls, v, info):
    bits = info.data.get("num_bits")
    if (bits is None) == (v is None):
        raise ValueError("Either num_bits or num_symbols must be specified." if v is None
                         else "Only one of num_bits or num_symbols should be specified.")
    return v


def _ratio_in_range(cls, v):
    if not 0.0 <= v <= 2.0:
        raise ValueError("prefix_length_ratio must be between 0 and 1 (inclusive).")
    return v


def _order_bound(cls, v):
    if not 2 <= v <= 4096:
        raise ValueError("Constellation order must be between 2 and 4096.")
    if v & (v - 1):
        raise ValueError(f"Constellation order must be a power of 2, got {v}.")
    return v


def _target_ser(cls, v):
    if v <= 0:
        raise ValueError("desired_symbol_error_rate must be positive.")
    if v >= 0.5:
        raise ValueError("desired_symbol_error_rate must be less than 0.5.")
    return v


def _validator(fn, *fields):
    return field_validator(*fields)(classmethod(fn))


class _SettingsText(BaseSettings):
    def __str__(self):
        return f"{self.project_name}\n{self.version}\nDebug Mode: {self.debug}"


Settings = create_model("Settings", __base__=_SettingsText, __module__=__name__,
                        project_name=(str, ...), version=(str, ...), debug=(bool, False))


class _SimulationText(BaseSettings):
    _report = (("Number of Bands", "num_bands", False, False),
               ("Signal-to-Noise Ratios", "signal_noise_ratios", False, False),
               ("Channel Type", "channel_type", False, False),
               ("Channel Model Path", "channel_model_path", True, False),
               ("Noise Type", "noise_type", False, False),
               ("Number of Bits", "num_bits", False, True),
               ("Number of Symbols", "num_symbols", False, True),
               ("Constellation Type", "constellation_type", True, False),
               ("Constellation Order", "constellation_order", False, False),
               ("Prefix Type", "prefix_type", False, False),
               ("Prefix Length Ratio", "prefix_length_ratio", False, False),
               ("Equalization Method", "equalization_method", False, False),
               ("Modulation Type", "modulation_type", False, False),
               ("Power Allocation Type", "power_allocation_type", False, False))


SimulationSettings = create_model(
    "SimulationSettings", __base__=_SimulationText, __module__=__name__,
    __validators__={
        "check_bits_or_symbols": _validator(_one_of_bits_or_symbols, "num_symbols"),
        "validate_prefix_length_ratio": _validator(_ratio_in_range, "prefix_length_ratio"),
        "validate_constellation_order": _validator(_order_bound, "min_constellation_order", "max_constellation_order"),
        "validate_desired_symbol_error_rate": _validator(_target_ser, "desired_symbol_error_rate"),
    },
    # ---- what is simulated
    num_bands=(int, ...),
    signal_noise_ratios=(list[float], ...),
    num_bits=(Optional[int], None),
    num_symbols=(Optional[int], None),
    # ---- channel and noise
    channel_model_path=(str, ...),
    channel_type=(E.ChannelType, E.ChannelType.FLAT),
    noise_type=(E.NoiseType, E.NoiseType.AWGN),
    # ---- waveform
    constellation_order=(int, 16),
    constellation_type=(E.ConstellationType, E.ConstellationType.PSK),
    modulation_type=(E.ModulationType, E.ModulationType.OFDM),
    prefix_type=(E.PrefixType, E.PrefixType.CYCLIC),
    prefix_length_ratio=(float, 0.25),
    equalization_method=(E.EqualizationMethod, E.EqualizationMethod.MMSE),
    # ---- loading
    power_allocation_type=(E.PowerAllocationType, E.PowerAllocationType.UNIFORM),
    adaptive_modulation_mode=(E.AdaptiveModulationMode, E.AdaptiveModulationMode.FIXED),
    min_constellation_order=(int, 4),
    max_constellation_order=(int, 256),
    desired_symbol_error_rate=(float, 1e-3),
)
