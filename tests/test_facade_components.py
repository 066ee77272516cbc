"""The host-side API classes (same names / signatures as the reference) are driven through the
reference's golden vectors exactly the way simulation/models.py:454-606 chains them, with the recorded
noise injected through the ``INoiseModel`` seam.  Every intermediate must equal what the live
reference produced (tests/golden/link_*.npz)."""
import io

import numpy as np
import pytest

from conftest import golden_link_names, load_golden

from ofdm_based_systems.bits_generation import AdaptiveBitsGenerator, RandomBitsGenerator
from ofdm_based_systems.channel.models import ChannelModel
from ofdm_based_systems.constellation import (AdaptiveConstellationMapper, PSKConstellationMapper,
                                              QAMConstellationMapper)
from ofdm_based_systems.equalization.models import MMSEEqualizator, NoEqualizator, ZeroForcingEqualizator
from ofdm_based_systems.modulation.models import OFDMModulator, SingleCarrierOFDMModulator
from ofdm_based_systems.noise.models import INoiseModel, NoNoiseModel
from ofdm_based_systems.prefix.models import CyclicPrefixScheme, NoPrefixScheme, ZeroPaddingPrefixScheme
from ofdm_based_systems.simulation.models import SerialToParallelConverter, read_bits_from_stream  # re-export (overview.py:20)

PREFIX = {"CYCLIC": CyclicPrefixScheme, "ZERO": ZeroPaddingPrefixScheme, "NONE": NoPrefixScheme}
EQ = {"ZF": ZeroForcingEqualizator, "MMSE": MMSEEqualizator, "NONE": NoEqualizator}
MAPPER = {"QAM": QAMConstellationMapper, "PSK": PSKConstellationMapper}


class ReplayNoise(INoiseModel):
    def __init__(self, noise):
        self.noise = noise

    def add_noise(self, signal, snr_db):
        return signal + self.noise


@pytest.mark.parametrize("name", golden_link_names())
def test_component_pipeline_reproduces_reference(name):
    g = load_golden("link", name)
    n = int(g["n_sc"])
    noise_model = ReplayNoise(g["noise"]) if bool(g["awgn"]) else NoNoiseModel()
    channel = ChannelModel(impulse_response=g["taps_raw"], snr_db=float(g["snr_db"]), noise_model=noise_model)
    prefix = PREFIX[str(g["prefix_type"])](prefix_length=int(g["prefix_len"]))
    eq = EQ[str(g["eq"])](channel_frequency_response=np.fft.fft(g["taps_raw"], n), snr_db=float(g["snr_db"]))
    mod_cls = OFDMModulator if str(g["modulator"]) == "OFDM" else SingleCarrierOFDMModulator
    mod = mod_cls(num_subcarriers=n, prefix_scheme=prefix, equalizator=eq)
    if g["orders"].size:
        mapper = AdaptiveConstellationMapper(g["orders"], MAPPER[str(g["scheme"])], n)
    else:
        mapper = MAPPER[str(g["scheme"])](order=int(g["order"]))
    bits = io.BytesIO(g["tx_bytes"].tobytes())
    tx_bits = read_bits_from_stream(bits)
    symbols = mapper.encode(bits)
    s2p = SerialToParallelConverter()
    tx = mod.modulate(s2p.to_parallel(symbols, n))
    rx = channel.transmit(s2p.to_serial(tx))
    z = s2p.to_serial(mod.demodulate(s2p.to_parallel(rx, n + prefix.prefix_length)))
    out = mapper.decode(z)
    rx_bits = read_bits_from_stream(out)
    np.testing.assert_array_equal(channel.impulse_response, g["taps_chan"])
    np.testing.assert_allclose(z, g["received_symbols"], rtol=0, atol=1e-12)
    assert out.getvalue() == g["rx_bytes"].tobytes()
    assert sum(a != b for a, b in zip(tx_bits, rx_bits)) == int(g["bit_errors"])
    assert int(np.sum(symbols != mapper.encode(out))) == int(g["symbol_errors"])
    if "symbols" in g.files:
        np.testing.assert_array_equal(symbols, g["symbols"])
        np.testing.assert_allclose(tx, g["tx"], rtol=0, atol=1e-13)
        np.testing.assert_allclose(rx, g["rx"], rtol=0, atol=1e-12)


def test_tables_and_rules_match_reference(kat):
    for m in (4, 16, 64, 256, 1024):
        np.testing.assert_array_equal(QAMConstellationMapper(order=m).constellation, kat[f"qam{m}"])
    for m in (2, 4, 8, 16, 32):
        np.testing.assert_array_equal(PSKConstellationMapper(order=m).constellation, kat[f"psk{m}"])
    snrs = kat["bl_snr"]
    assert [QAMConstellationMapper.calculate_bit_loading_order(1e-3, s) for s in snrs] == kat["bl_qam_1e-3"].tolist()
    assert [PSKConstellationMapper.calculate_bit_loading_order(1e-3, s) for s in snrs] == kat["bl_psk_1e-3"].tolist()
    from ofdm_based_systems.power_allocation.models import WaterfillingPowerAllocation
    np.testing.assert_array_equal(WaterfillingPowerAllocation(5.0, kat["wf1_gains"], 0.1).allocate(), kat["wf1_power"])
    np.testing.assert_array_equal(WaterfillingPowerAllocation(1.0, kat["wf2_gains"], 0.1).allocate(), kat["wf2_power"])


def test_generators_follow_the_bytes_contract():
    gen = np.random.Generator(np.random.PCG64(5))
    ref = np.random.Generator(np.random.PCG64(5)).bytes(2)
    out = RandomBitsGenerator(generator=gen).generate_bits(11).read()
    assert out == bytes([ref[0], ref[1] & 0b11100000])
    a = AdaptiveBitsGenerator(np.array([2, 4, 0, 6]), 10)
    assert a.get_total_bits() == 120 and len(a.generate_bits().read()) == 15


def test_simulation_plan_matches_reference_setup():
    """Simulation.plan() (the host-side set-up that precedes the CUDA launch) against Simulation.run() of the
    live reference: orders, allocated power, water level, totals (tests/golden/sim_*.npz)."""
    from ofdm_based_systems.configuration.enums import (AdaptiveModulationMode, EqualizationMethod,
                                                        PowerAllocationType)
    from ofdm_based_systems.simulation.models import Simulation
    g = load_golden("sim", "adaptive_wf_mmse")
    taps = load_golden("link", "adaptive_p1_n64_mmse")["taps_raw"]
    sim = Simulation(num_symbols=40, adaptive_modulation_mode=AdaptiveModulationMode.CAPACITY_BASED,
                     power_allocation_type=PowerAllocationType.WATERFILLING, channel_impulse_response=taps,
                     num_subcarriers=64, snr_db=20.0, verbose=False)
    pl = sim.plan()
    np.testing.assert_array_equal(pl["orders"], g["constellation_order_per_subcarrier"])
    np.testing.assert_array_equal(pl["power"], g["allocated_power"])
    assert pl["water_level"] == float(g["water_level"])
    assert pl["total_bits"] == int(g["total_bits"])
    with pytest.raises(ValueError, match="Either num_bits or num_symbols"):
        Simulation()
    with pytest.raises(ValueError, match="Only one of num_bits or num_symbols"):
        Simulation(num_bits=8, num_symbols=8)
