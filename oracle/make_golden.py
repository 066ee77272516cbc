"""Generate tests/golden/*.npz by running the LIVE, UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

Run in the build container (where /root/reference is mounted read-only):

    python oracle/make_golden.py

The reference is imported from /root/reference/src (never copied).  matplotlib is absent from this
image, so a no-op stub (oracle/_mplstub) is put on sys.path for ``simulation/models.py`` to import.

Two kinds of fixtures are written:

* ``link_*.npz``  - the component pipeline of tests/integration/test_end_to_end.py:205-257 driven with
  a seeded ``Generator(PCG64(seed))`` for the bits and a recording ``INoiseModel`` for the noise.
  Inputs (bytes, scaled noise, taps) and every intermediate (X, tx, rx, Y, Z, decoded bytes, counts)
  are stored, so that the oracle AND the CUDA replay kernel can be checked without the reference.
* ``sim_*.npz``   - ``Simulation.run()`` itself (fixed and adaptive / water-filling modes) with the
  module-level default bit generator and the global NumPy RNG seeded, result-dict values stored.
* ``kat.npz``     - constellation tables, Gray tables, bit-loading and water-filling known answers.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("OFDM_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

sys.path.insert(0, os.path.join(REF, "src"))
sys.path.insert(1, os.path.join(HERE, "_mplstub"))

from numpy.random import PCG64, Generator  # noqa: E402

from ofdm_based_systems.bits_generation.models import AdaptiveBitsGenerator, RandomBitsGenerator  # noqa: E402
from ofdm_based_systems.channel.models import ChannelModel  # noqa: E402
from ofdm_based_systems.configuration.enums import (  # noqa: E402
    AdaptiveModulationMode, ConstellationType, EqualizationMethod, ModulationType, NoiseType,
    PowerAllocationType, PrefixType)
from ofdm_based_systems.constellation.adaptive import (  # noqa: E402
    AdaptiveConstellationMapper, calculate_constellation_orders)
from ofdm_based_systems.constellation.models import (  # noqa: E402
    GrayWordCoder, PSKConstellationMapper, QAMConstellationMapper)
from ofdm_based_systems.equalization.models import (  # noqa: E402
    MMSEEqualizator, NoEqualizator, ZeroForcingEqualizator)
from ofdm_based_systems.modulation.models import OFDMModulator, SingleCarrierOFDMModulator  # noqa: E402
from ofdm_based_systems.noise.models import AWGNoiseModel, INoiseModel, NoNoiseModel  # noqa: E402
from ofdm_based_systems.power_allocation.models import (  # noqa: E402
    UniformPowerAllocation, WaterfillingPowerAllocation, calculate_capacity,
    calculate_capacity_per_subcarrier)
from ofdm_based_systems.prefix.models import (  # noqa: E402
    CyclicPrefixScheme, NoPrefixScheme, ZeroPaddingPrefixScheme)
from ofdm_based_systems.serial_parallel.models import SerialToParallelConverter  # noqa: E402
from ofdm_based_systems.simulation.models import Simulation, read_bits_from_stream  # noqa: E402

CHAN = os.path.join(REF, "config", "channel_models")


class RecordingAWGN(AWGNoiseModel):
    """Calls the reference's own add_noise and keeps what it added.  The added noise is recovered
    exactly by replaying the global RNG state the reference consumed."""

    def __init__(self):
        self.noise = None
        self.signal = None

    def add_noise(self, signal, snr_db):
        state = np.random.get_state()
        out = super().add_noise(signal, snr_db)
        after = np.random.get_state()
        np.random.set_state(state)
        re = np.random.normal(size=signal.shape)
        im = np.random.normal(size=signal.shape)
        np.random.set_state(after)
        self.normal_re, self.normal_im = re, im
        power = np.mean(np.abs(signal) ** 2) / (10 ** (snr_db / 10))
        self.noise = np.sqrt(power / 2) * (re + 1j * im)
        assert np.array_equal(signal + self.noise, out), "noise replay does not match the reference"
        self.signal = signal.copy()
        return out


PREFIX = {"CYCLIC": CyclicPrefixScheme, "ZERO": ZeroPaddingPrefixScheme, "NONE": NoPrefixScheme}
EQ = {"ZF": ZeroForcingEqualizator, "MMSE": MMSEEqualizator, "NONE": NoEqualizator}
MAPPER = {"QAM": QAMConstellationMapper, "PSK": PSKConstellationMapper}


def link_case(name, n_sc, order, scheme, taps_raw, prefix_type, prefix_len, eq, snr_db, n_ofdm, seed,
              modulator="OFDM", orders=None, awgn=True, keep=("Y",), amp=None, rx_gain=None, kind="link"):
    """Drive the reference's component classes exactly as simulation/models.py:454-606 does.
    ``amp`` / ``rx_gain`` (kind="loaded"): the applied power loading of the reference's experiments around its own
    components - the two NumPy lines of examples/waterfilling_noise_bump_experiment.py:148 (parallel * sqrt(P)) and
    :165-169 (demodulated / sqrt(P)), which Simulation.run() itself never executes (simulation/models.py:508)."""
    taps_raw = np.asarray(taps_raw, dtype=np.complex128)
    np.random.seed(seed)
    gen = Generator(PCG64(seed))
    noise_model = RecordingAWGN() if awgn else NoNoiseModel()
    channel = ChannelModel(impulse_response=taps_raw, snr_db=snr_db, noise_model=noise_model)
    prefix = PREFIX[prefix_type](prefix_length=prefix_len)
    H_eq = np.fft.fft(taps_raw, n_sc)                      # simulation/models.py:263-266
    equalizer = EQ[eq](channel_frequency_response=H_eq, snr_db=snr_db)
    mod_cls = OFDMModulator if modulator == "OFDM" else SingleCarrierOFDMModulator
    mod = mod_cls(num_subcarriers=n_sc, prefix_scheme=prefix, equalizator=equalizer)
    s2p = SerialToParallelConverter()
    if orders is None:
        mapper = MAPPER[scheme](order=order)
        total_bits = n_ofdm * n_sc * mapper.bits_per_symbol
        bits = RandomBitsGenerator(generator=gen).generate_bits(total_bits)
    else:
        orders = np.asarray(orders, dtype=np.int64)
        mapper = AdaptiveConstellationMapper(orders, MAPPER[scheme], n_sc)
        bgen = AdaptiveBitsGenerator(mapper.get_bits_per_subcarrier(), n_ofdm, generator=gen)
        total_bits = bgen.get_total_bits()
        bits = bgen.generate_bits()
    tx_bytes = bits.getvalue()
    bits_list = read_bits_from_stream(bits)
    symbols = mapper.encode(bits)
    parallel = s2p.to_parallel(symbols, n_sc)
    if amp is not None:
        parallel = parallel * np.asarray(amp, dtype=np.float64)
    tx = mod.modulate(parallel)
    p = np.abs(tx) ** 2
    papr_db = 10 * np.log10(np.max(p) / np.mean(p))
    serial = s2p.to_serial(tx)
    rx = channel.transmit(serial)
    rx_par = s2p.to_parallel(rx, n_sc + prefix.prefix_length)
    demod = mod.demodulate(rx_par)
    if rx_gain is not None:
        demod = demod * np.asarray(rx_gain, dtype=np.float64)
    z = s2p.to_serial(demod)
    rx_stream = mapper.decode(z)
    rx_bytes = rx_stream.getvalue()
    rx_list = read_bits_from_stream(rx_stream)
    bit_errors = sum(a != b for a, b in zip(bits_list, rx_list))
    recoded = mapper.encode(rx_stream)
    symbol_errors = int(np.sum(symbols != recoded))
    # Y (pre-equaliser) recomputed with the reference's own prefix class
    r = np.array([prefix.remove_prefix(row) for row in rx_par])
    Y = np.fft.fft(r, n=n_sc, axis=1, norm="ortho")
    d = dict(name=name, n_sc=n_sc, order=order, scheme=scheme, taps_raw=taps_raw,
             taps_chan=channel.impulse_response, H_eq=H_eq, prefix_type=prefix_type,
             prefix_len=prefix_len, eq=eq, snr_db=float(snr_db), n_ofdm=n_ofdm, seed=seed,
             modulator=modulator, awgn=awgn, total_bits=total_bits,
             orders=(np.zeros(0, dtype=np.int64) if orders is None else orders),
             tx_bytes=np.frombuffer(tx_bytes, dtype=np.uint8),
             noise=(noise_model.noise if awgn else np.zeros(0, dtype=np.complex128)),
             rx_bytes=np.frombuffer(rx_bytes, dtype=np.uint8), received_symbols=z,
             bit_errors=int(bit_errors), symbol_errors=symbol_errors, papr_db=float(papr_db))
    if amp is not None:
        d["amp"] = np.asarray(amp, dtype=np.float64)
    if rx_gain is not None:
        d["rx_gain"] = np.asarray(rx_gain, dtype=np.float64)
    if awgn and "normals" in keep:
        d["normal_re"], d["normal_im"] = noise_model.normal_re, noise_model.normal_im
    inter = dict(symbols=symbols, tx=tx, rx=rx, Y=Y)
    for k in keep:
        if k == "all":
            d.update(inter)
        elif k in inter:
            d[k] = inter[k]
    path = os.path.join(OUT, f"{kind}_{name}.npz")
    np.savez_compressed(path, **d)
    print(f"{name:28s} bits={total_bits:7d} bit_errors={bit_errors:6d} sym_errors={symbol_errors:6d} "
          f"papr={papr_db:.3f} dB  {os.path.getsize(path) / 1024:.0f} KB")


def seed_default_generators(seed):
    """Seed the two RNGs Simulation.run() uses implicitly (quirk Q9): the shared default-argument
    Generator of RandomBitsGenerator / AdaptiveBitsGenerator and the global legacy RNG."""
    for cls in (RandomBitsGenerator, AdaptiveBitsGenerator):
        g = cls.__init__.__defaults__[-1]
        g.bit_generator.state = PCG64(seed).state
    np.random.seed(seed)


def sim_case(name, seed, **kw):
    seed_default_generators(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        res = Simulation(verbose=False, **kw).run()
    d = dict(name=name, seed=seed)
    for k, v in kw.items():
        d["arg_" + k] = (v.value if hasattr(v, "value") else v)
    for k in ("bit_errors", "symbol_errors", "total_bits", "bit_error_rate", "symbol_error_rate", "papr_db",
              "title", "subtitle", "prefix_acronym", "power_allocation_acronym", "bitrate_mbps"):
        d[k] = res[k]
    d["water_level"] = np.nan if res["water_level"] is None else res["water_level"]
    d["constellation_order_per_subcarrier"] = np.array(res["constellation_order_per_subcarrier"], dtype=np.int64)
    d["allocated_power"] = np.array(res["allocated_power"], dtype=np.float64)
    d["received_symbols"] = res["received_symbols"]
    d["keys"] = np.array(sorted(res.keys()))
    path = os.path.join(OUT, f"sim_{name}.npz")
    np.savez_compressed(path, **d)
    print(f"sim {name:24s} bits={res['total_bits']:7d} bit_errors={res['bit_errors']:6d} "
          f"sym_errors={int(res['symbol_errors']):6d} papr={res['papr_db']:.3f}  {os.path.getsize(path) / 1024:.0f} KB")


def kat():
    d = {}
    for m in (4, 16, 64, 256, 1024):
        d[f"qam{m}"] = QAMConstellationMapper(order=m).constellation
    for m in (2, 4, 8, 16, 32):
        d[f"psk{m}"] = PSKConstellationMapper(order=m).constellation
    for b in (2, 3, 4):
        c = GrayWordCoder(bits_per_word=b)
        d[f"gray{b}"] = np.array([c.gray_table[i] for i in range(1 << b)])
        d[f"igray{b}"] = np.array([c.inverse_gray_table[i] for i in range(1 << b)])
    snrs = np.array([0.5, 3, 10, 30, 100, 300, 1e3, 1e4, 1e5])
    with contextlib.redirect_stdout(io.StringIO()):
        d["bl_snr"] = snrs
        d["bl_qam_1e-3"] = np.array([QAMConstellationMapper.calculate_bit_loading_order(1e-3, s) for s in snrs])
        d["bl_psk_1e-3"] = np.array([PSKConstellationMapper.calculate_bit_loading_order(1e-3, s) for s in snrs])
        d["bl_qam_1e-2"] = np.array([QAMConstellationMapper.calculate_bit_loading_order(1e-2, s) for s in snrs])
        d["bl_psk_1e-5"] = np.array([PSKConstellationMapper.calculate_bit_loading_order(1e-5, s) for s in snrs])
    # water-filling KATs + one per shipped channel
    g1, g2 = np.array([1, .8, .6, .4, .2]), np.array([1, .9, .01, .001])
    d["wf1_gains"], d["wf1_power"] = g1, WaterfillingPowerAllocation(5.0, g1, 0.1).allocate()
    d["wf2_gains"], d["wf2_power"] = g2, WaterfillingPowerAllocation(1.0, g2, 0.1).allocate()
    names = sorted(f[:-4] for f in os.listdir(CHAN) if f.endswith(".npy"))
    d["channel_names"] = np.array(names)
    for nm in names:
        h = np.load(os.path.join(CHAN, nm + ".npy"))
        d["chan_" + nm] = h
        for n_sc in (64, 1024):
            for snr in (5.0, 20.0):
                gains = np.abs(np.fft.fft(h, n_sc)) ** 2
                n0 = 10 ** (-snr / 10)
                wf = WaterfillingPowerAllocation(float(n_sc), gains, n0)
                pw = wf.allocate()
                key = f"wf_{nm}_{n_sc}_{int(snr)}"
                d[key + "_power"] = pw
                with contextlib.redirect_stdout(io.StringIO()):
                    d[key + "_orders"] = np.array(
                        [QAMConstellationMapper.calculate_bit_loading_order(1e-3, p * g / n0) for p, g in zip(pw, gains)])
                d[key + "_cap"] = calculate_capacity_per_subcarrier(pw, gains, n0)
    cap = np.array([8.5, 6.2, 3.1, 1.5, 0.2, 9.9, 4.0])
    d["shannon_cap"] = cap
    d["shannon_qam"] = calculate_constellation_orders(cap, 4, 256, 1.0, QAMConstellationMapper)
    d["shannon_psk"] = calculate_constellation_orders(cap, 4, 256, 0.85, PSKConstellationMapper)
    d["uniform_5_4"] = UniformPowerAllocation(5.0, 4).allocate()
    np.savez_compressed(os.path.join(OUT, "kat.npz"), **d)
    print("kat.npz written")


def main():
    os.makedirs(OUT, exist_ok=True)
    ch = {n: np.load(os.path.join(CHAN, n + ".npy")) for n in
          ("severe_multipath", "Lin-Phoong_P1", "Lin-Phoong_P2", "flat_fading", "rayleigh_fading", "two_ray",
           "default_multipath")}
    kat()
    # BASELINE.json config #1 shape (tests/integration/test_end_to_end.py:205-257): N=64, QPSK, CP=16, AWGN, ZF
    link_case("c1_n64_qpsk_cp16_awgn_zf", 64, 4, "QAM", ch["flat_fading"], "CYCLIC", 16, "ZF", 6.0, 40, 11, keep=("all", "normals"))
    # config/simulation_settings_test.json shape: N=64, 16-QAM, ZF, severe_multipath, P=7
    link_case("test_json_n64_16qam_zf", 64, 16, "QAM", ch["severe_multipath"], "CYCLIC", 7, "ZF", 20.0, 60, 12)
    # config #2 shape: N=1024, 16-QAM, 8 taps, MMSE
    link_case("c2_n1024_16qam_mmse", 1024, 16, "QAM", ch["severe_multipath"], "CYCLIC", 7, "MMSE", 16.0, 8, 13)
    # headline: N=1024, 64-QAM, MMSE, 8 taps
    link_case("headline_n1024_64qam_mmse", 1024, 64, "QAM", ch["severe_multipath"], "CYCLIC", 7, "MMSE", 20.0, 10, 14)
    # config #3: custom channel, 64-QAM, ZF vs MMSE
    link_case("c3_n64_64qam_zf_p2", 64, 64, "QAM", ch["Lin-Phoong_P2"], "CYCLIC", 3, "ZF", 25.0, 60, 15)
    link_case("c3_n64_64qam_mmse_p2", 64, 64, "QAM", ch["Lin-Phoong_P2"], "CYCLIC", 3, "MMSE", 25.0, 60, 15)
    # config #5 shape: N=4096, 256-QAM, MMSE
    link_case("c5_n4096_256qam_mmse", 4096, 256, "QAM", ch["severe_multipath"], "CYCLIC", 7, "MMSE", 30.0, 3, 16)
    # zero padding (overlap-add), 6 taps
    link_case("zp_n64_16qam_mmse", 64, 16, "QAM", ch["rayleigh_fading"], "ZERO", 5, "MMSE", 15.0, 50, 17, keep=("all", "normals"))
    link_case("zp_n256_64qam_zf", 256, 64, "QAM", ch["severe_multipath"], "ZERO", 7, "ZF", 24.0, 12, 18)
    # short prefix / no prefix -> inter-symbol interference across OFDM symbols
    link_case("isi_cp3_n256_64qam_mmse", 256, 64, "QAM", ch["severe_multipath"], "CYCLIC", 3, "MMSE", 22.0, 12, 19, keep=("all", "normals"))
    link_case("isi_none_n128_16qam_zf", 128, 16, "QAM", ch["severe_multipath"], "NONE", 0, "ZF", 22.0, 20, 20)
    link_case("isi_zp2_n64_qpsk_mmse", 64, 4, "QAM", ch["rayleigh_fading"], "ZERO", 2, "MMSE", 12.0, 40, 21)
    # long prefix (ratio > 1)
    link_case("cp10_n64_16qam_mmse", 64, 16, "QAM", ch["severe_multipath"], "CYCLIC", 10, "MMSE", 14.0, 40, 22)
    # SC-OFDM (the default config/simulation_settings.json uses it)
    link_case("sc_n64_qpsk_zf_p1", 64, 4, "QAM", ch["Lin-Phoong_P1"], "CYCLIC", 3, "ZF", 8.0, 60, 23, modulator="SC-OFDM", keep=("all", "normals"))
    link_case("sc_n256_16qam_mmse", 256, 16, "QAM", ch["severe_multipath"], "CYCLIC", 7, "MMSE", 18.0, 12, 24,
              modulator="SC-OFDM")
    # PSK
    link_case("psk8_n128_mmse_two_ray", 128, 8, "PSK", ch["two_ray"], "CYCLIC", 1, "MMSE", 14.0, 30, 25, keep=("all", "normals"))
    link_case("psk2_n64_zf_flat", 64, 2, "PSK", ch["flat_fading"], "CYCLIC", 4, "ZF", 4.0, 60, 26)
    link_case("psk16_n64_none_eq_flat", 64, 16, "PSK", ch["flat_fading"], "NONE", 0, "NONE", 20.0, 40, 27)
    # raw, non-unit-energy taps: equaliser sees raw H, channel applies normalised taps (quirk Q3)
    link_case("rawtaps_n64_16qam_mmse", 64, 16, "QAM", np.array([1.0, 0.4, 0.2]), "CYCLIC", 16, "MMSE", 20.0, 40, 28)
    link_case("rawtaps_n64_16qam_zf", 64, 16, "QAM", np.array([1.0, 0.4, 0.2]), "CYCLIC", 16, "ZF", 20.0, 40, 28)
    # no-noise model
    link_case("nonoise_n64_256qam_zf", 64, 256, "QAM", ch["Lin-Phoong_P1"], "CYCLIC", 3, "ZF", 100.0, 20, 29, awgn=False)
    # higher orders
    link_case("n512_256qam_zf", 512, 256, "QAM", ch["default_multipath"], "CYCLIC", 3, "ZF", 28.0, 8, 30)
    link_case("n2048_1024qam_mmse", 2048, 1024, "QAM", ch["two_ray"], "CYCLIC", 1, "MMSE", 34.0, 3, 31)
    # adaptive per-subcarrier loading (config #4 shape, one realisation): orders from the reference's own rule
    for nm, n_sc, snr, seed in (("p1", 64, 20.0, 32), ("severe", 256, 24.0, 33)):
        h = ch["Lin-Phoong_P1"] if nm == "p1" else ch["severe_multipath"]
        gains = np.abs(np.fft.fft(h, n_sc)) ** 2
        n0 = 10 ** (-snr / 10)
        pw = WaterfillingPowerAllocation(float(n_sc), gains, n0).allocate()
        with contextlib.redirect_stdout(io.StringIO()):
            orders = np.array([QAMConstellationMapper.calculate_bit_loading_order(1e-3, p * g / n0)
                               for p, g in zip(pw, gains)], dtype=np.int64)
        n_ofdm = 40 if n_sc == 64 else 16
        link_case(f"adaptive_{nm}_n{n_sc}_mmse", n_sc, 0, "QAM", h, "CYCLIC", len(h) - 1, "MMSE", snr, n_ofdm, seed,
                  orders=orders, keep=("all",) if n_sc == 64 else ("Y",))
    # adaptive with inactive subcarriers and PSK
    orders = np.array([0, 2, 4, 8, 16, 4, 0, 2] * 8, dtype=np.int64)
    link_case("adaptive_psk_n64_zf", 64, 0, "PSK", ch["two_ray"], "CYCLIC", 1, "ZF", 18.0, 24, 34, orders=orders)

    loaded_cases(ch)
    noise_bump_cases(ch)
    ber_study(ch)

    # Simulation.run() itself
    common = dict(num_subcarriers=64, snr_db=18.0)
    sim_case("default_fixed", 41, num_symbols=64 * 40, **common)
    sim_case("fixed_wf_zf_custom", 42, num_symbols=64 * 40, constellation_order=64,
             equalizator_type=EqualizationMethod.ZF, power_allocation_type=PowerAllocationType.WATERFILLING,
             channel_impulse_response=ch["severe_multipath"], **common)
    sim_case("adaptive_wf_mmse", 43, num_symbols=40, adaptive_modulation_mode=AdaptiveModulationMode.CAPACITY_BASED,
             power_allocation_type=PowerAllocationType.WATERFILLING, channel_impulse_response=ch["Lin-Phoong_P1"],
             num_subcarriers=64, snr_db=20.0)
    sim_case("adaptive_uniform_zf", 44, num_symbols=40, adaptive_modulation_mode=AdaptiveModulationMode.CAPACITY_BASED,
             equalizator_type=EqualizationMethod.ZF, channel_impulse_response=ch["severe_multipath"],
             num_subcarriers=64, snr_db=25.0)
    sim_case("sc_zp_psk", 45, num_bits=64 * 3 * 40, constellation_order=8, constellation_scheme=ConstellationType.PSK,
             modulator_type=ModulationType.SC_OFDM, prefix_scheme=PrefixType.ZERO, num_subcarriers=64, snr_db=15.0)
    sim_case("noprefix_nonoise", 46, num_symbols=64 * 20, prefix_scheme=PrefixType.NONE, noise_scheme=NoiseType.NONE,
             equalizator_type=EqualizationMethod.NONE, num_subcarriers=64, snr_db=30.0)


def ber_study(ch):
    """BER statistics of the LIVE reference with its OWN random generators, per OFDM symbol, so that the GPU's
    independent-RNG BER can be placed inside the reference's confidence intervals (north_star):
      * the headline link (N=1024, 64-QAM, MMSE, severe_multipath, CP=7) at 8 SNR points, 1600 OFDM symbols (9.8e6 bits) each;
      * the short-prefix study of docs/OFDM-Based Systems.tex:226-264 as the CURRENT code runs it:
        Lin-Phoong P2, N=64, 64-QAM, 30 dB, {ZF, MMSE} x {CP, ZP} x prefix ratio {0.34, 0.68, 1.00, 1.34}, 1600 OFDM
        symbols (614 400 bits) per entry."""
    def errors_per_symbol(n_sc, order, taps, prefix_type, prefix_len, eq, snr_db, n_ofdm, seed):
        np.random.seed(seed)
        gen = Generator(PCG64(seed))
        channel = ChannelModel(impulse_response=np.asarray(taps, dtype=np.complex128), snr_db=snr_db, noise_model=AWGNoiseModel())
        prefix = PREFIX[prefix_type](prefix_length=prefix_len)
        equalizer = EQ[eq](channel_frequency_response=np.fft.fft(taps, n_sc), snr_db=snr_db)
        mod = OFDMModulator(num_subcarriers=n_sc, prefix_scheme=prefix, equalizator=equalizer)
        s2p = SerialToParallelConverter()
        mapper = QAMConstellationMapper(order=order)
        bps = mapper.bits_per_symbol
        with contextlib.redirect_stdout(io.StringIO()):
            bits = RandomBitsGenerator(generator=gen).generate_bits(n_ofdm * n_sc * bps)
            tx_bits = np.array(read_bits_from_stream(bits), dtype=np.uint8)
            tx = s2p.to_serial(mod.modulate(s2p.to_parallel(mapper.encode(bits), n_sc)))
            rx = s2p.to_parallel(channel.transmit(tx), n_sc + prefix.prefix_length)
            rx_bits = np.array(read_bits_from_stream(mapper.decode(s2p.to_serial(mod.demodulate(rx)))), dtype=np.uint8)
        return np.sum((tx_bits != rx_bits).reshape(n_ofdm, n_sc * bps), axis=1).astype(np.int32)

    head_snrs = np.array([0.0, 4.0, 8.0, 12.0, 16.0, 20.0, 24.0, 28.0])
    # 4 x 400 OFDM symbols per point (the reference's nearest-neighbour classifier holds n x M complex128: 400 symbols at a time)
    head = np.stack([np.concatenate([errors_per_symbol(1024, 64, ch["severe_multipath"], "CYCLIC", 7, "MMSE", snr, 400, 700 + i + 100 * r)
                                     for r in range(4)]) for i, snr in enumerate(head_snrs)])
    ratios = np.array([0.34, 0.68, 1.00, 1.34])
    taps = ch["Lin-Phoong_P2"]
    sp = np.zeros((2, 2, 4, 1600), dtype=np.int32)
    for a, eq in enumerate(("ZF", "MMSE")):
        for b, pre in enumerate(("CYCLIC", "ZERO")):
            for c, ratio in enumerate(ratios):
                sp[a, b, c] = errors_per_symbol(64, 64, taps, pre, int(ratio * (len(taps) - 1)), eq, 30.0, 1600, 800 + 16 * a + 4 * b + c)
    path = os.path.join(OUT, "ber_reference.npz")
    np.savez_compressed(path, headline_snrs=head_snrs, headline_errors=head, headline_bits_per_symbol=1024 * 6,
                        sp_eq=np.array(["ZF", "MMSE"]), sp_prefix=np.array(["CYCLIC", "ZERO"]), sp_ratio=ratios,
                        sp_errors=sp, sp_bits_per_symbol=64 * 6, sp_snr_db=30.0)
    print("headline BER:", np.round(head.sum(axis=1) / (head.shape[1] * 6144), 5))
    print("short-prefix BER (ZF/MMSE x CP/ZP x ratio):\n", np.round(sp.sum(axis=3) / (1600 * 384), 4))


def loaded_cases(ch):
    """Applied power loading: the reference's WaterfillingPowerAllocation feeds tx amplitudes sqrt(P_k)
    and receiver gains 1/sqrt(P_k) around the reference's own mapper / modulator / channel / decoder."""
    for nm, n_sc, order, eq, snr, n_ofdm, seed in (("p2_n64_16qam_zf", 64, 16, "ZF", 16.0, 40, 51),
                                                    ("severe_n1024_64qam_mmse", 1024, 64, "MMSE", 24.0, 6, 52),
                                                    ("rayleigh_n256_16qam_mmse", 256, 16, "MMSE", 14.0, 16, 53)):
        h = {"p2": ch["Lin-Phoong_P2"], "severe": ch["severe_multipath"], "rayleigh": ch["rayleigh_fading"]}[nm.split("_")[0]]
        gains = np.abs(np.fft.fft(h, n_sc)) ** 2
        power = WaterfillingPowerAllocation(1.0, gains, 10 ** (-snr / 10)).allocate()
        power = np.maximum(power * n_sc, 1e-4)          # mean power 1 per subcarrier, floored like the experiment (:146-147)
        amp = np.sqrt(power)
        link_case(nm, n_sc, order, "QAM", h, "CYCLIC", len(h) - 1, eq, snr, n_ofdm, seed, amp=amp, rx_gain=1.0 / amp,
                  kind="loaded")


def noise_bump_cases(ch):
    """SURVEY 8f-2, second half: the post-equaliser stage of examples/waterfilling_noise_bump_experiment.py:120-185 run with
    the reference's OWN components, statement for statement (noise-free channel, power loading, MMSE equaliser, coloured
    noise injected after the equaliser, receiver compensation, division by the square root of the block's mean power,
    fixed-order decoder), recording the bits and the noise matrix it drew.  Two of the script's three scenarios at two of
    its SNR points, shortened to 96 OFDM symbols."""
    h = ch["Lin-Phoong_P2"]
    n_sc, order, n_ofdm = 64, 64, 96
    prefix_len = len(h) - 1
    H = np.fft.fft(h, n_sc)
    gains = np.abs(H) ** 2
    for nm, allocation, bump_db, snr_db, seed in (("uniform_bump3_snr20", "UNIFORM", 3.0, 20.0, 61),
                                                   ("wf_bump3_snr15", "WATERFILLING", 3.0, 15.0, 62),
                                                   ("wf_bump6_snr25", "WATERFILLING", 6.0, 25.0, 63)):
        np.random.seed(seed)
        gen = Generator(PCG64(seed))
        profile = np.ones(n_sc)
        profile[int(0.75 * n_sc):] = 10 ** (bump_db / 10)                      # create_noise_profile, :42-51
        mapper = QAMConstellationMapper(order=order)
        s2p = SerialToParallelConverter()
        prefix = PREFIX["CYCLIC"](prefix_length=prefix_len)
        channel = ChannelModel(impulse_response=h, snr_db=0.0, noise_model=NoNoiseModel())
        total_bits = n_ofdm * n_sc * mapper.bits_per_symbol
        with contextlib.redirect_stdout(io.StringIO()):
            bits = RandomBitsGenerator(generator=gen).generate_bits(total_bits)
            tx_bytes = bits.getvalue()
            bits_list = read_bits_from_stream(bits)
            symbols = mapper.encode(bits)
            parallel = s2p.to_parallel(symbols, n_sc)
            noise_power = 10 ** (-snr_db / 10)
            if allocation == "WATERFILLING":
                power = WaterfillingPowerAllocation(total_power=1.0, channel_gains=gains / profile, noise_power=noise_power).allocate()
                power = np.maximum(power, 1e-4)
                power = power / np.sum(power)
            else:
                power = UniformPowerAllocation(total_power=1.0, num_subcarriers=n_sc).allocate()
            parallel = parallel * np.sqrt(power)
            equalizer = EQ["MMSE"](channel_frequency_response=H, snr_db=snr_db)
            mod = OFDMModulator(num_subcarriers=n_sc, prefix_scheme=prefix, equalizator=equalizer)
            tx = mod.modulate(parallel)
            rx = channel.transmit(s2p.to_serial(tx))
            demod = mod.demodulate(s2p.to_parallel(rx, n_sc + prefix_len))
            noise_std = np.sqrt(noise_power * profile / 2.0)[np.newaxis, :]
            noise_matrix = (np.random.normal(size=demod.shape) + 1j * np.random.normal(size=demod.shape)) * noise_std
            noisy = demod + noise_matrix
            power_sqrt_safe = np.sqrt(power)
            power_sqrt_safe[power_sqrt_safe < 1e-10] = 1.0
            noisy = noisy / power_sqrt_safe
            z = s2p.to_serial(noisy)
            avg_power = np.mean(np.abs(z) ** 2)
            if avg_power > 1e-12:
                z = z / np.sqrt(avg_power)
            rx_stream = mapper.decode(z)
            rx_bytes = rx_stream.getvalue()
            rx_list = read_bits_from_stream(rx_stream)
        bit_errors = sum(a != b for a, b in zip(bits_list, rx_list))
        pw = np.abs(tx) ** 2
        path = os.path.join(OUT, f"noisebump_{nm}.npz")
        np.savez_compressed(path, name=nm, n_sc=n_sc, order=order, n_ofdm=n_ofdm, prefix_len=prefix_len, snr_db=snr_db, seed=seed,
                            taps_raw=h, taps_chan=channel.impulse_response, H_eq=H, eq="MMSE", allocation=allocation,
                            noise_profile=profile, amp=np.sqrt(power), rx_gain=1.0 / power_sqrt_safe, total_bits=total_bits,
                            tx_bytes=np.frombuffer(tx_bytes, dtype=np.uint8), post_noise=noise_matrix,
                            rx_bytes=np.frombuffer(rx_bytes, dtype=np.uint8), received_symbols=z, avg_power=float(avg_power),
                            bit_errors=int(bit_errors), papr_db=float(10 * np.log10(np.max(pw) / np.mean(pw))))
        print(f"noisebump_{nm:24s} bits={total_bits} bit_errors={bit_errors} avg_power={avg_power:.5f} {os.path.getsize(path) / 1024:.0f} KB")


if __name__ == "__main__":
    if "--noisebump-only" in sys.argv:
        os.makedirs(OUT, exist_ok=True)
        noise_bump_cases({"Lin-Phoong_P2": np.load(os.path.join(CHAN, "Lin-Phoong_P2.npy"))})
    elif "--ber-only" in sys.argv:
        os.makedirs(OUT, exist_ok=True)
        names = ["Lin-Phoong_P2", "severe_multipath"]
        ber_study({n: np.load(os.path.join(CHAN, n + ".npy")) for n in names})
    elif "--loaded-only" in sys.argv:
        os.makedirs(OUT, exist_ok=True)
        names = ["Lin-Phoong_P1", "Lin-Phoong_P2", "default_multipath", "flat_fading", "rayleigh_fading", "severe_multipath", "two_ray"]
        loaded_cases({n: np.load(os.path.join(CHAN, n + ".npy")) for n in names})
    else:
        main()
