"""``Simulation``: the orchestrator of one SNR point (reference: simulation/models.py:72-818).

Constructor, registries, set-up arithmetic and the 29-key result dict follow the reference.  The hot
path the reference runs between :454 and :606 (bits -> map -> IFFT -> prefix -> FIR channel -> AWGN ->
FFT -> ZF/MMSE -> demap -> error counts, PAPR) is ONE CUDA kernel launch in libofdm_b200.so; there is no
CPU implementation of it in this package and ``run()`` raises ``RuntimeError`` without a CUDA device.
"""
from __future__ import annotations

import os
import time
from io import BytesIO
from typing import Any, BinaryIO, Dict, List, Optional, Type

import numpy as np
from numpy.typing import NDArray

# The reference's callers and tests reach the component classes through this module's namespace too
# (e.g. ``simulation.models.SerialToParallelConverter``), so every name the reference imports here is kept importable.
from ofdm_based_systems.bits_generation.models import AdaptiveBitsGenerator, IGenerator, RandomBitsGenerator
from ofdm_based_systems.channel.models import ChannelModel
from ofdm_based_systems.configuration.enums import (
    AdaptiveModulationMode, ConstellationType, EqualizationMethod, ModulationType, NoiseType, PowerAllocationType,
    PrefixType)
from ofdm_based_systems.configuration.models import SimulationSettings
from ofdm_based_systems.constellation.adaptive import AdaptiveConstellationMapper, calculate_constellation_orders
from ofdm_based_systems.constellation.models import (
    IConstellationMapper, PSKConstellationMapper, QAMConstellationMapper)
from ofdm_based_systems.equalization.models import MMSEEqualizator, NoEqualizator, ZeroForcingEqualizator
from ofdm_based_systems.modulation.models import IModulator, OFDMModulator, SingleCarrierOFDMModulator
from ofdm_based_systems.noise.models import AWGNoiseModel, NoNoiseModel
from ofdm_based_systems.power_allocation.models import (
    UniformPowerAllocation, WaterfillingPowerAllocation, calculate_capacity_per_subcarrier)
from ofdm_based_systems.prefix.models import CyclicPrefixScheme, NoPrefixScheme, ZeroPaddingPrefixScheme
from ofdm_based_systems.serial_parallel.models import SerialToParallelConverter
from ofdm_based_systems.simulation.plotting import draw_constellation_image

# 4-tap channel used when no impulse response is given (simulation/models.py:236-246; ChannelType.FLAT, quirk Q4)
DEFAULT_IMPULSE_RESPONSE = np.array([
    7.767824138452235072e-01 + 4.560896742466611919e-01j,
    -6.669848996328063551e-02 + 2.839935704583463338e-01j,
    1.398968327715586490e-01 - 1.591963958343969865e-01j,
    2.229949514514480494e-02 + 2.409945439452868821e-01j,
], dtype=np.complex128)

# bound on results["received_symbols"] / the scatter plot: the reference keeps every equalised symbol,
# which does not scale to 1e9+ bits; runs up to this many OFDM-symbol samples return all of them
MAX_RETURNED_SAMPLES = 1 << 20


def read_bits_from_stream(stream: BinaryIO) -> List[int]:
    """Bytes -> list of 0/1 ints, MSB first; rewinds the stream (simulation/models.py:59-69)."""
    data = stream.read()
    stream.seek(0)
    return np.unpackbits(np.frombuffer(data, dtype=np.uint8), bitorder="big").tolist()


class Simulation:
    CONSTELLATION_SCHEME_MAPPERS = {ConstellationType.QAM: QAMConstellationMapper,
                                    ConstellationType.PSK: PSKConstellationMapper}
    MODULATOR_SCHEME_MAPPERS = {ModulationType.OFDM: OFDMModulator, ModulationType.SC_OFDM: SingleCarrierOFDMModulator}
    PREFIX_SCHEME_MAPPERS = {PrefixType.NONE: NoPrefixScheme, PrefixType.CYCLIC: CyclicPrefixScheme,
                             PrefixType.ZERO: ZeroPaddingPrefixScheme}
    EQUALIZATOR_SCHEME_MAPPERS = {EqualizationMethod.NONE: NoEqualizator, EqualizationMethod.ZF: ZeroForcingEqualizator,
                                  EqualizationMethod.MMSE: MMSEEqualizator}
    NOISE_SCHEME_MAPPERS = {NoiseType.AWGN: AWGNoiseModel, NoiseType.NONE: NoNoiseModel}
    POWER_ALLOCATION_MAPPERS = {PowerAllocationType.UNIFORM: UniformPowerAllocation,
                                PowerAllocationType.WATERFILLING: WaterfillingPowerAllocation}

    def __init__(self, num_bits: Optional[int] = None, num_symbols: Optional[int] = None, num_subcarriers: int = 64,
                 constellation_order: int = 16, constellation_scheme: ConstellationType = ConstellationType.QAM,
                 modulator_type: ModulationType = ModulationType.OFDM, prefix_scheme: PrefixType = PrefixType.CYCLIC,
                 prefix_length_ratio: float = 1.0, equalizator_type: EqualizationMethod = EqualizationMethod.MMSE,
                 snr_db: float = 20.0, noise_scheme: NoiseType = NoiseType.AWGN,
                 power_allocation_type: PowerAllocationType = PowerAllocationType.UNIFORM,
                 adaptive_modulation_mode: AdaptiveModulationMode = AdaptiveModulationMode.FIXED,
                 min_constellation_order: int = 4, max_constellation_order: int = 256,
                 desired_symbol_error_rate: float = 1e-3,
                 channel_impulse_response: Optional[NDArray[np.complex128]] = None, verbose: bool = True):
        if num_bits is None and num_symbols is None:
            raise ValueError("Either num_bits or num_symbols must be provided.")
        if num_bits is not None and num_symbols is not None:
            raise ValueError("Only one of num_bits or num_symbols should be provided.")
        self.num_bits = num_bits
        self.num_symbols = num_symbols
        self.num_subcarriers = num_subcarriers
        self.constellation_order = constellation_order
        self.constellation_scheme = constellation_scheme
        self.modulator_type = modulator_type
        self.prefix_scheme = prefix_scheme
        self.prefix_length_ratio = prefix_length_ratio
        self.equalizator_type = equalizator_type
        self.snr_db = snr_db
        self.noise_scheme = noise_scheme
        self.power_allocation_type = power_allocation_type
        self.adaptive_modulation_mode = adaptive_modulation_mode
        self.min_constellation_order = min_constellation_order
        self.max_constellation_order = max_constellation_order
        self.desired_symbol_error_rate = desired_symbol_error_rate
        self.channel_impulse_response = channel_impulse_response
        self.verbose = verbose

    def _log(self, message: str) -> None:
        if self.verbose:
            print(message)

    @classmethod
    def create_from_simulation_settings(cls, simulation_settings: SimulationSettings) -> List["Simulation"]:
        """One Simulation per SNR value; a CUSTOM channel is loaded from its .npy (path relative to CWD)."""
        taps = None
        if simulation_settings.channel_type.value == "CUSTOM":
            path = simulation_settings.channel_model_path
            if not path:
                raise ValueError("channel_model_path must be specified when channel_type is CUSTOM")
            if not path.startswith("/"):
                path = os.path.abspath(path)
            if not os.path.exists(path):
                raise FileNotFoundError(f"Channel model file not found: {path}")
            try:
                taps = np.load(path)
            except Exception as exc:  # noqa: BLE001 - same contract as the reference
                raise ValueError(f"Failed to load channel model from {path}: {exc}")
            print(f"Loaded custom channel impulse response from: {path} ({len(taps)} taps, {taps.dtype})")
        s = simulation_settings
        return [cls(num_bits=s.num_bits, num_symbols=s.num_symbols, num_subcarriers=s.num_bands,
                    constellation_order=s.constellation_order, constellation_scheme=s.constellation_type,
                    modulator_type=s.modulation_type, prefix_scheme=s.prefix_type,
                    prefix_length_ratio=s.prefix_length_ratio, equalizator_type=s.equalization_method, snr_db=snr,
                    noise_scheme=s.noise_type, power_allocation_type=s.power_allocation_type,
                    adaptive_modulation_mode=s.adaptive_modulation_mode,
                    min_constellation_order=s.min_constellation_order, max_constellation_order=s.max_constellation_order,
                    desired_symbol_error_rate=s.desired_symbol_error_rate, channel_impulse_response=taps)
                for snr in s.signal_noise_ratios]

    # ------------------------------------------------------------------------------------------
    def plan(self) -> Dict[str, Any]:
        """Everything ``run()`` derives before the hot loop (simulation/models.py:226-410), on the host in
        fp64: channel, prefix length, equaliser response, loading.  Returned as a dict so that the sweep
        runner (main.py) can launch many SNR points without re-deriving it."""
        from ofdm_based_systems.simulation.sweep import LinkConfig
        taps = (self.channel_impulse_response if self.channel_impulse_response is not None
                else DEFAULT_IMPULSE_RESPONSE)
        taps = np.asarray(taps, dtype=np.complex128)
        n = self.num_subcarriers
        channel = ChannelModel(impulse_response=taps, snr_db=self.snr_db,
                               noise_model=self.NOISE_SCHEME_MAPPERS.get(self.noise_scheme, AWGNoiseModel)())
        prefix_length = 0 if self.prefix_scheme == PrefixType.NONE else int(self.prefix_length_ratio * channel.order)
        prefix = self.PREFIX_SCHEME_MAPPERS.get(self.prefix_scheme, NoPrefixScheme)(prefix_length=prefix_length)
        gains = np.abs(np.fft.fft(taps, n)) ** 2                    # RAW taps (quirk Q3)
        noise_power = 10 ** (-self.snr_db / 10)
        mapper_cls: Type[IConstellationMapper] = self.CONSTELLATION_SCHEME_MAPPERS.get(
            self.constellation_scheme, QAMConstellationMapper)
        water_level: Optional[float] = None
        adaptive = self.adaptive_modulation_mode == AdaptiveModulationMode.CAPACITY_BASED
        if adaptive:
            if self.power_allocation_type == PowerAllocationType.WATERFILLING:
                power = WaterfillingPowerAllocation(total_power=n, channel_gains=gains, noise_power=noise_power).allocate()
                levels = power + noise_power / gains
                water_level = float(np.mean(levels[power > 1e-10]))
            else:
                power = UniformPowerAllocation(total_power=n, num_subcarriers=n).allocate()
            orders = np.array([mapper_cls.calculate_bit_loading_order(ser=self.desired_symbol_error_rate,
                                                                      snr=p * g / noise_power)
                               for p, g in zip(power, gains)], dtype=np.int64)
            mapper: IConstellationMapper = AdaptiveConstellationMapper(constellation_orders=orders,
                                                                       base_mapper_class=mapper_cls, num_subcarriers=n)
            bits_per_ofdm = int(np.sum(mapper.get_bits_per_subcarrier()))
            if self.num_symbols is not None:
                num_ofdm = self.num_symbols                        # quirk Q8: OFDM symbols in adaptive mode
            else:
                if bits_per_ofdm == 0:
                    raise ValueError("All subcarriers have zero order - cannot transmit data")
                num_ofdm = self.num_bits // bits_per_ofdm
            if num_ofdm <= 0:      # AdaptiveBitsGenerator (bits_generation/models.py:95-120) refuses an empty run
                raise ValueError(f"num_ofdm_symbols must be positive, got {num_ofdm}")
            total_bits = bits_per_ofdm * num_ofdm
            if bits_per_ofdm == 0:
                raise ValueError("No active subcarriers (all orders are zero)")
            if (8 * -(-total_bits // 8)) % bits_per_ofdm != 0:     # quirk Q11 (constellation/adaptive.py:168-172)
                raise ValueError(f"Bits length ({8 * -(-total_bits // 8)}) must be multiple of "
                                 f"bits_per_symbol ({bits_per_ofdm})")
            num_constellation_symbols = num_ofdm * n
        else:
            orders = np.full(n, self.constellation_order, dtype=np.int64)
            mapper = mapper_cls(order=self.constellation_order)
            bps = mapper.bits_per_symbol
            total_bits = self.num_bits if self.num_symbols is None else self.num_symbols * int(np.log2(self.constellation_order))
            if total_bits is None:
                raise ValueError("Total bits could not be determined.")
            stream_bits = 8 * -(-total_bits // 8)
            num_constellation_symbols = -(-stream_bits // bps)
            if num_constellation_symbols % n != 0:
                raise ValueError("Length of data must be divisible by number of streams.")
            num_ofdm = num_constellation_symbols // n
            bits_per_ofdm = n * bps
            power = None
        cfg = LinkConfig(num_subcarriers=n, taps_raw=taps, constellation_order=self.constellation_order,
                         constellation_scheme=self.constellation_scheme.value, modulator_type=self.modulator_type.value,
                         prefix_scheme=self.prefix_scheme.value, prefix_length=prefix_length,
                         equalizator_type=self.equalizator_type.value, awgn=self.noise_scheme == NoiseType.AWGN,
                         orders=orders)
        return dict(cfg=cfg, channel=channel, prefix=prefix, mapper=mapper, orders=orders, power=power,
                    water_level=water_level, gains=gains, noise_power=noise_power, total_bits=total_bits,
                    num_ofdm=num_ofdm, num_constellation_symbols=num_constellation_symbols,
                    bits_per_ofdm=bits_per_ofdm, adaptive=adaptive)

    def run(self) -> Dict[str, Any]:
        """One SNR point (simulation/models.py:214-818).  Configuration errors surface before the device is touched."""
        return Simulation.run_sweep([self])[0]

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _link_key(pl: Dict[str, Any]):
        """Simulations whose plans agree on this key run on ONE configured link (they differ in SNR only)."""
        cfg = pl["cfg"]
        return (cfg.num_subcarriers, cfg.taps_raw.tobytes(), cfg.orders.tobytes(), cfg.constellation_scheme,
                cfg.modulator_type, cfg.prefix_scheme, cfg.prefix_length, cfg.equalizator_type, cfg.awgn,
                pl["num_ofdm"], pl["total_bits"], pl["adaptive"])

    @staticmethod
    def run_sweep(simulations: List["Simulation"], group=None) -> List[Dict[str, Any]]:
        """The reference's sequential SNR loop (main.py:234-240) as sharded launches: the simulations that share a link
        (the list ``create_from_simulation_settings`` returns in FIXED mode: one link, K SNR values) run as ONE kernel
        launch with the SNR point as a grid dimension; with an initialised ``torch.distributed`` process group every
        rank takes a contiguous share of the OFDM-symbol range and ONE all-reduce per sweep combines the counters.
        Adaptive loading depends on the SNR, so those simulations are groups of one.  Results come back in input order."""
        from ofdm_based_systems import _native
        from ofdm_based_systems.simulation.sweep import LinkSweep
        plans = [sim.plan() for sim in simulations]          # ValueErrors of the configuration come first
        _native.require_gpu()
        groups: Dict[Any, List[int]] = {}
        for i, pl in enumerate(plans):
            groups.setdefault(Simulation._link_key(pl), []).append(i)
        results: List[Optional[Dict[str, Any]]] = [None] * len(simulations)
        for members in groups.values():
            pl0, sims = plans[members[0]], [simulations[i] for i in members]
            cfg, num_ofdm, total_bits, mapper = pl0["cfg"], pl0["num_ofdm"], pl0["total_bits"], pl0["mapper"]
            start = time.perf_counter()
            seed = int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 31)
            sweep = LinkSweep(cfg)
            link = sweep.link
            try:
                snrs = [float(sim.snr_db) for sim in sims]
                ragged = (not pl0["adaptive"]) and total_bits != pl0["num_constellation_symbols"] * mapper.bits_per_symbol
                sample_ofdm = max(1, min(num_ofdm, MAX_RETURNED_SAMPLES // cfg.num_subcarriers))
                if ragged:
                    pairs = [sim._run_ragged(link, cfg, pl0, cfg.noise_sigma(snr)) for sim, snr in zip(sims, snrs)]
                    counters, received = [c for c, _ in pairs], [z for _, z in pairs]
                else:
                    # ---- the hot path: ONE launch for every SNR point of this link ...
                    counters = sweep.sweep_counters(snrs, num_ofdm, seed=seed, group=group)
                    # ---- ... plus a bounded dump launch per point for results["received_symbols"] / the scatter plot
                    received = []
                    for k, snr in enumerate(snrs):
                        _, dump = link.run_fused(snr, cfg.noise_sigma(snr), sample_ofdm, seed=seed, point=k, dump=("z",))
                        received.append(dump["z"].reshape(-1).astype(np.complex128))
            finally:
                sweep.close()
            elapsed_ms = (time.perf_counter() - start) * 1000 / len(members)
            for i, c, z in zip(members, counters, received):
                results[i] = simulations[i]._report(plans[i], c, z, elapsed_ms, truncated=sample_ofdm < num_ofdm and not ragged)
        return results  # type: ignore[return-value]

    def _report(self, pl: Dict[str, Any], counters, received, elapsed_ms: float, truncated: bool) -> Dict[str, Any]:
        """The reference's prints and its 29-key result dict (simulation/models.py:413-444, 509-524, 597-620, 796-810)
        from the counters of one SNR point."""
        cfg, prefix, mapper, orders = pl["cfg"], pl["prefix"], pl["mapper"], pl["orders"]
        total_bits, water_level = pl["total_bits"], pl["water_level"]
        print("=" * 50)
        print("Starting OFDM-based System Simulation")
        print("=" * 50)
        self._log("Using custom channel impulse response (%d taps)" % len(cfg.taps_raw)
                  if self.channel_impulse_response is not None else "Using default multipath channel (4 taps)")
        print(f"Using prefix length: {cfg.prefix_length}")
        print(f"Signal to noise ratio: {self.snr_db} dB")

        results: Dict[str, Any] = {
            "num_bits": self.num_bits, "num_symbols": self.num_symbols, "num_subcarriers": self.num_subcarriers,
            "constellation_order": self.constellation_order, "constellation_scheme": self.constellation_scheme.name,
            "modulator_type": self.modulator_type.name, "prefix_scheme": self.prefix_scheme.name,
            "prefix_acronym": prefix.acronym, "equalizator_type": self.equalizator_type.name, "snr_db": self.snr_db,
            "noise_scheme": self.noise_scheme.name, "power_allocation_type": self.power_allocation_type.name,
            "power_allocation_acronym": "WF" if self.power_allocation_type == PowerAllocationType.WATERFILLING else "UNIFORM",
            "adaptive_modulation_mode": self.adaptive_modulation_mode.name,
            "constellation_order_per_subcarrier": orders.tolist(),
            "water_level": water_level,     # quirk Q10: stays None in FIXED mode, the value is computed later
            "title": f"{prefix.acronym}-{self.modulator_type.name}-{self.equalizator_type.name}",
            "subtitle": f"{self.constellation_order}{self.constellation_scheme.name}-SNR{self.snr_db}dB-{self.power_allocation_type.name}",
        }

        # power allocation is computed and REPORTED, never applied to the symbols (quirk Q1, :483-509)
        power = pl["power"]
        if not pl["adaptive"]:
            if self.power_allocation_type == PowerAllocationType.WATERFILLING:
                power = WaterfillingPowerAllocation(total_power=1.0, channel_gains=pl["gains"],
                                                    noise_power=pl["noise_power"]).allocate()
            else:
                power = UniformPowerAllocation(total_power=1.0, num_subcarriers=self.num_subcarriers).allocate()
        self._log(f"Power allocation computed: min={power.min():.6f}, max={power.max():.6f}")
        results["allocated_power"] = power.tolist()

        papr_db = np.float64(counters.papr_db)
        print(f"PAPR: {papr_db:.2f} dB")
        results["papr_db"] = papr_db
        bit_errors = int(counters.bit_errors)
        symbol_errors = np.int64(counters.symbol_errors)
        ber = bit_errors / total_bits if total_bits > 0 else 0.0
        n_sym = pl["num_constellation_symbols"]
        ser = symbol_errors / n_sym if n_sym > 0 else 0.0
        print("=" * 50)
        print(f"Bit Errors: {bit_errors} out of {total_bits} bits")
        print(f"Bit Error Rate (BER): {ber:.6f}")
        print(f"Symbol Errors: {symbol_errors} out of {n_sym} symbols")
        print(f"Symbol Error Rate (SER): {ser:.6f}")
        print("=" * 50)
        results.update({"bit_errors": bit_errors, "symbol_errors": symbol_errors, "total_bits": total_bits,
                        "bit_error_rate": ber, "symbol_error_rate": ser, "received_symbols": received})
        if truncated:
            # the reference returns every equalised symbol; beyond MAX_RETURNED_SAMPLES only the first OFDM symbols are
            # returned (and drawn) - this extra key says so, the counters above cover the whole run
            results["received_symbols_truncated"] = True
        results["constellation_plot"] = draw_constellation_image(
            received, mapper.constellation, title=results["title"], ber=ber, snr_db=self.snr_db, papr_db=float(papr_db),
            orders=orders if pl["adaptive"] else None, num_subcarriers=self.num_subcarriers)
        results["transmission_time_ms"] = elapsed_ms
        results["bitrate_mbps"] = total_bits / 1e6          # quirk Q2: not divided by time in the reference either
        print(f"Transmission time: {elapsed_ms:.2f} ms")
        print(f"Bitrate: {results['bitrate_mbps']:.2f} Mbps")
        print("Simulation completed.")
        print("=" * 50)
        return results

    def _run_ragged(self, link, cfg, pl, sigma):
        """num_bits that does not fill its last byte / constellation symbol: the reference zero-pads the
        stream (constellation/models.py:235-237) and zip() stops at the shorter list (:597).  The bits then
        come from the host generator and go through the kernel's replay entry with the compare limit set."""
        total_bits, num_ofdm = pl["total_bits"], pl["num_ofdm"]
        stream = RandomBitsGenerator().generate_bits(total_bits).read()
        noise = None
        if sigma > 0:
            shape = num_ofdm * (cfg.num_subcarriers + cfg.prefix_length)
            noise = sigma * (np.random.normal(size=shape) + 1j * np.random.normal(size=shape))
        counters, dump = link.run_replay(self.snr_db, stream, noise, num_ofdm, compare_limit_bits=8 * len(stream),
                                         dump=("z",))
        return counters, dump["z"].reshape(-1).astype(np.complex128)
