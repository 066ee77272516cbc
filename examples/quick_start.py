"""Fixed and adaptive configuration through the reference's own entry points.  Run from the repository root."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ofdm-based-systems_b200"))

from ofdm_based_systems.configuration.enums import (AdaptiveModulationMode, ChannelType, ConstellationType,
                                                    EqualizationMethod, ModulationType, PowerAllocationType, PrefixType)
from ofdm_based_systems.configuration.models import SimulationSettings
from ofdm_based_systems.simulation.models import Simulation

common = dict(num_bands=64, signal_noise_ratios=[15.0, 25.0], channel_type=ChannelType.CUSTOM,
              channel_model_path=os.path.join(ROOT, "config", "channel_models", "Lin-Phoong_P1.npy"),
              constellation_type=ConstellationType.QAM, prefix_type=PrefixType.CYCLIC, prefix_length_ratio=1.0,
              modulation_type=ModulationType.OFDM, equalization_method=EqualizationMethod.MMSE)

fixed = SimulationSettings(num_symbols=64 * 2000, constellation_order=16, **common)
adaptive = SimulationSettings(num_symbols=2000, constellation_order=16, power_allocation_type=PowerAllocationType.WATERFILLING,
                              adaptive_modulation_mode=AdaptiveModulationMode.CAPACITY_BASED, min_constellation_order=4,
                              max_constellation_order=256, capacity_scaling_factor=0.85, **common)

for name, settings in (("fixed 16-QAM", fixed), ("water-filling + adaptive loading", adaptive)):
    for sim in Simulation.create_from_simulation_settings(settings):
        sim.verbose = False
        res = sim.run()
        orders = sorted(set(res["constellation_order_per_subcarrier"]))
        print(f"{name:34s} {res['title']:12s} {res['subtitle']:32s} bits={res['total_bits']:8d} "
              f"BER={res['bit_error_rate']:.3e} SER={res['symbol_error_rate']:.3e} PAPR={res['papr_db']:.2f} dB orders={orders}")
assert len(res) == 29
