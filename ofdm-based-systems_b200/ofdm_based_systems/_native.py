"""ctypes binding of libofdm_b200.so (C ABI in include/ofdm_b200.h).

This is the ONLY compute backend of ``Simulation.run()`` and of the sweep runner: there is no CPU
fallback.  Importing this module without the built library raises; using it without a CUDA device
raises ``RuntimeError`` at the first call.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("OFDM_B200_LIB", os.path.join(_PKG_ROOT, "libofdm_b200.so"))

PREFIX = {"NONE": 0, "CYCLIC": 1, "ZERO": 2}
MODULATOR = {"OFDM": 0, "SC-OFDM": 1, "SC_OFDM": 1}
EQUALIZER = {"NONE": 0, "ZF": 1, "MMSE": 2}
SCHEME = {"QAM": 0, "PSK": 1}
NOISE_NONE, NOISE_C64, NOISE_C128 = 0, 2, 3


class NativeLibraryMissing(ImportError):
    pass


class LinkDesc(C.Structure):
    _fields_ = [("n_subcarriers", C.c_int32), ("prefix_type", C.c_int32), ("prefix_len", C.c_int32),
                ("modulator", C.c_int32), ("equalizer", C.c_int32), ("scheme", C.c_int32),
                ("n_taps", C.c_int32), ("device", C.c_int32)]


class LinkResult(C.Structure):
    _fields_ = [("bit_errors", C.c_uint64), ("bits", C.c_uint64), ("symbol_errors", C.c_uint64),
                ("symbols", C.c_uint64), ("ofdm_symbols", C.c_uint64), ("tx_samples", C.c_uint64),
                ("tx_power_sum", C.c_double), ("tx_power_max", C.c_double)]


class WaterfillDesc(C.Structure):
    _fields_ = [("n_subcarriers", C.c_int32), ("n_taps", C.c_int32), ("scheme", C.c_int32), ("waterfilling", C.c_int32),
                ("min_order", C.c_int32), ("max_order", C.c_int32), ("snr_db", C.c_double), ("total_power", C.c_double),
                ("gap", C.c_double), ("tolerance", C.c_double), ("order_rule", C.c_int32), ("reserved", C.c_int32),
                ("capacity_scaling", C.c_double)]


class FramesDesc(C.Structure):
    _fields_ = [("n_subcarriers", C.c_int32), ("prefix_len", C.c_int32), ("equalizer", C.c_int32), ("n_taps", C.c_int32),
                ("loading", C.c_int32), ("fixed_order", C.c_int32), ("waterfilling", C.c_int32), ("min_order", C.c_int32),
                ("max_order", C.c_int32), ("device", C.c_int32), ("snr_db", C.c_double), ("gap", C.c_double)]


class LinkDump(C.Structure):
    _fields_ = [("y", C.c_void_p), ("z", C.c_void_p), ("rx_labels", C.c_void_p), ("tx_labels", C.c_void_p),
                ("noise", C.c_void_p)]


class LinkPost(C.Structure):
    _fields_ = [("noise_profile", C.c_void_p), ("recorded_noise", C.c_void_p), ("z_scale", C.c_double),
                ("measure_power", C.c_int32), ("reserved", C.c_int32)]


EXPORTS = (
    "ofdm_b200_last_error", "ofdm_b200_abi_version", "ofdm_b200_device_count", "ofdm_b200_launch_count",
    "ofdm_b200_measure_fp32_tflops", "ofdm_link_create", "ofdm_link_create_loaded", "ofdm_link_destroy", "ofdm_link_bits_per_ofdm_symbol", "ofdm_link_table_bytes",
    "ofdm_link_uses_fast_kernel",
    "ofdm_link_run_fused", "ofdm_link_run_replay", "ofdm_link_launch_fused", "ofdm_link_launch_replay",
    "ofdm_link_run_sweep", "ofdm_link_launch_sweep", "ofdm_link_read_sweep", "ofdm_link_pack_sweep",
    "ofdm_link_set_post", "ofdm_link_read_z_power",
    "ofdm_link_reset_counters", "ofdm_link_read_result", "ofdm_link_counters_device_ptr", "ofdm_link_pack_counters",
    "ofdm_waterfill_bitload_batched", "ofdm_waterfill_bitload_batched_dev", "ofdm_frames_run",
    "ofdm_link_debug_tables", "ofdm_frames_debug_tables", "ofdm_frames_header_floats",
)


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} is missing: build it with `python ofdm-based-systems_b200/build_native.py` "
            "(the OFDM link has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, u64, u32, i32, dbl = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32, C.c_double
    lib.ofdm_b200_last_error.restype = C.c_char_p
    lib.ofdm_b200_abi_version.restype = i32
    lib.ofdm_b200_device_count.restype = i32
    lib.ofdm_b200_launch_count.restype = u64
    lib.ofdm_b200_measure_fp32_tflops.restype = dbl
    lib.ofdm_b200_measure_fp32_tflops.argtypes = [i32]
    lib.ofdm_link_create.argtypes = [C.POINTER(LinkDesc), vp, vp, vp, vp, C.POINTER(vp)]
    lib.ofdm_link_create_loaded.argtypes = [C.POINTER(LinkDesc), vp, vp, vp, vp, vp, C.POINTER(vp)]
    lib.ofdm_link_destroy.argtypes = [vp]
    lib.ofdm_link_destroy.restype = None
    lib.ofdm_link_bits_per_ofdm_symbol.argtypes = [vp]
    lib.ofdm_link_table_bytes.argtypes = [vp]
    lib.ofdm_link_uses_fast_kernel.argtypes = [vp]
    lib.ofdm_link_table_bytes.restype = u64
    lib.ofdm_link_run_fused.argtypes = [vp, dbl, dbl, u64, u32, u64, u64, C.POINTER(LinkDump), C.POINTER(LinkResult)]
    lib.ofdm_link_run_replay.argtypes = [vp, dbl, vp, u64, vp, i32, u64, u64, C.POINTER(LinkDump), C.POINTER(LinkResult)]
    lib.ofdm_link_launch_fused.argtypes = [vp, dbl, dbl, u64, u32, u64, u64, C.POINTER(LinkDump), vp]
    lib.ofdm_link_launch_replay.argtypes = [vp, dbl, vp, u64, vp, i32, u64, u64, C.POINTER(LinkDump), vp]
    lib.ofdm_link_run_sweep.argtypes = [vp, i32, vp, vp, u64, u32, u64, u64, vp]
    lib.ofdm_link_launch_sweep.argtypes = [vp, i32, vp, vp, u64, u32, u64, u64, vp]
    lib.ofdm_link_read_sweep.argtypes = [vp, vp, i32, vp]
    lib.ofdm_link_pack_sweep.argtypes = [vp, vp, i32, i32, vp]
    lib.ofdm_link_set_post.argtypes = [vp, C.POINTER(LinkPost)]
    lib.ofdm_link_read_z_power.argtypes = [vp, vp, C.POINTER(dbl), C.POINTER(u64)]
    lib.ofdm_link_reset_counters.argtypes = [vp, vp]
    lib.ofdm_link_read_result.argtypes = [vp, vp, C.POINTER(LinkResult)]
    lib.ofdm_link_counters_device_ptr.argtypes = [vp]
    lib.ofdm_link_pack_counters.argtypes = [vp, vp, i32, i32, vp]
    lib.ofdm_link_counters_device_ptr.restype = vp
    lib.ofdm_waterfill_bitload_batched.argtypes = [C.POINTER(WaterfillDesc), vp, C.c_int64, vp, vp, vp, vp, vp, vp]
    lib.ofdm_waterfill_bitload_batched_dev.argtypes = [C.POINTER(WaterfillDesc), vp, C.c_int64, vp, vp, vp, vp, vp, vp, vp]
    lib.ofdm_frames_run.argtypes = [C.POINTER(FramesDesc), vp, C.c_int64, u64, u64, u32, u64, C.POINTER(LinkResult), vp, vp, vp]
    lib.ofdm_link_debug_tables.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.ofdm_frames_debug_tables.argtypes = [C.POINTER(FramesDesc), vp, C.c_int64, u64, u64, vp, vp, vp, vp]
    lib.ofdm_frames_header_floats.restype = i32
    if lib.ofdm_b200_abi_version() != 1:
        raise NativeLibraryMissing(f"{LIB_PATH}: ABI version {lib.ofdm_b200_abi_version()} != 1, rebuild it")
    return lib


lib = _load()


def last_error() -> str:
    return (lib.ofdm_b200_last_error() or b"").decode()


def device_count() -> int:
    return int(lib.ofdm_b200_device_count())


def require_gpu() -> None:
    n = device_count()
    if n <= 0:
        raise RuntimeError("ofdm_based_systems needs a CUDA device (B200): the link chain runs only in "
                           f"libofdm_b200.so and has no CPU fallback ({last_error() or 'no device found'})")


def _check(rc: int) -> None:
    if rc == 0:
        return
    msg = last_error()
    if rc in (-1, -3):
        raise ValueError(msg)
    if rc == -4:
        raise MemoryError(msg)
    raise RuntimeError(msg)


@dataclass
class LinkCounters:
    bit_errors: int
    bits: int
    symbol_errors: int
    symbols: int
    ofdm_symbols: int
    tx_samples: int
    tx_power_sum: float
    tx_power_max: float

    @property
    def papr_db(self) -> float:
        """simulation/models.py:519-522."""
        if self.tx_samples == 0 or self.tx_power_sum <= 0:
            return float("inf")
        return float(10 * np.log10(self.tx_power_max / (self.tx_power_sum / self.tx_samples)))

    @classmethod
    def from_struct(cls, r: LinkResult) -> "LinkCounters":
        return cls(r.bit_errors, r.bits, r.symbol_errors, r.symbols, r.ofdm_symbols, r.tx_samples,
                   r.tx_power_sum, r.tx_power_max)


class Link:
    """One configured link resident on one GPU (ofdm_link_create)."""

    def __init__(self, n_subcarriers: int, taps_chan: np.ndarray, h_eq: np.ndarray, orders: np.ndarray, *,
                 prefix_type: str = "CYCLIC", prefix_len: int = 0, modulator: str = "OFDM", equalizer: str = "MMSE",
                 scheme: str = "QAM", amp: Optional[np.ndarray] = None, rx_gain: Optional[np.ndarray] = None,
                 device: int = -1):
        require_gpu()
        taps = np.ascontiguousarray(taps_chan, dtype=np.complex128)
        heq = np.ascontiguousarray(h_eq, dtype=np.complex128)
        ords = np.ascontiguousarray(orders, dtype=np.int32)
        if heq.shape != (n_subcarriers,) or ords.shape != (n_subcarriers,):
            raise ValueError("h_eq and orders must have one entry per subcarrier")
        a = None if amp is None else np.ascontiguousarray(amp, dtype=np.float64)
        g = None if rx_gain is None else np.ascontiguousarray(rx_gain, dtype=np.float64)
        for arr in (a, g):
            if arr is not None and arr.shape != (n_subcarriers,):
                raise ValueError("amp and rx_gain must have one entry per subcarrier")
        self.n_subcarriers, self.prefix_len = int(n_subcarriers), int(prefix_len)
        self.desc = LinkDesc(n_subcarriers, PREFIX[prefix_type], prefix_len, MODULATOR[modulator], EQUALIZER[equalizer],
                             SCHEME[scheme], taps.shape[0], device)
        self._h = C.c_void_p()
        _check(lib.ofdm_link_create_loaded(C.byref(self.desc), taps.ctypes.data, heq.ctypes.data, ords.ctypes.data,
                                           None if a is None else a.ctypes.data, None if g is None else g.ctypes.data,
                                           C.byref(self._h)))
        self.bits_per_ofdm_symbol = int(lib.ofdm_link_bits_per_ofdm_symbol(self._h))
        self.table_bytes = int(lib.ofdm_link_table_bytes(self._h))
        self.uses_fast_kernel = bool(lib.ofdm_link_uses_fast_kernel(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            lib.ofdm_link_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    # ---- host-buffer entry points (synchronous)
    def _dump_arrays(self, n_symbols: int, want: Optional[tuple]):
        if not want:
            return None, {}
        n, npre = self.n_subcarriers, self.n_subcarriers + self.prefix_len
        arrs = {}
        if "y" in want:
            arrs["y"] = np.zeros((n_symbols, n), dtype=np.complex64)
        if "z" in want:
            arrs["z"] = np.zeros((n_symbols, n), dtype=np.complex64)
        if "rx_labels" in want:
            arrs["rx_labels"] = np.zeros((n_symbols, n), dtype=np.uint16)
        if "tx_labels" in want:
            arrs["tx_labels"] = np.zeros((n_symbols, n), dtype=np.uint16)
        if "noise" in want:
            arrs["noise"] = np.zeros((n_symbols, npre), dtype=np.complex64)
        d = LinkDump(*[arrs[k].ctypes.data if k in arrs else None for k in ("y", "z", "rx_labels", "tx_labels", "noise")])
        return d, arrs

    def run_fused(self, snr_db: float, noise_sigma: float, n_symbols: int, *, seed: int = 0x0FD3, point: int = 0,
                  first_symbol: int = 0, dump: Optional[tuple] = None):
        d, arrs = self._dump_arrays(n_symbols, dump)
        res = LinkResult()
        _check(lib.ofdm_link_run_fused(self._h, snr_db, noise_sigma, seed, point, first_symbol, n_symbols,
                                       None if d is None else C.byref(d), C.byref(res)))
        out = LinkCounters.from_struct(res)
        return (out, arrs) if dump else out

    def run_replay(self, snr_db: float, bits: bytes, noise: Optional[np.ndarray], n_symbols: int, *,
                   compare_limit_bits: int = 0, dump: Optional[tuple] = None):
        buf = np.frombuffer(bits, dtype=np.uint8) if not isinstance(bits, np.ndarray) else np.ascontiguousarray(bits, dtype=np.uint8)
        if noise is None:
            nptr, ndt = None, NOISE_NONE
        else:
            noise = np.ascontiguousarray(noise)
            if noise.dtype == np.complex64:
                ndt = NOISE_C64
            elif noise.dtype == np.complex128:
                ndt = NOISE_C128
            else:
                raise ValueError("noise must be complex64 or complex128")
            if noise.size != n_symbols * (self.n_subcarriers + self.prefix_len):
                raise ValueError("noise must hold (N + P) samples per OFDM symbol")
            nptr = noise.ctypes.data
        d, arrs = self._dump_arrays(n_symbols, dump)
        res = LinkResult()
        _check(lib.ofdm_link_run_replay(self._h, snr_db, buf.ctypes.data, buf.size, nptr, ndt, n_symbols,
                                        compare_limit_bits, None if d is None else C.byref(d), C.byref(res)))
        out = LinkCounters.from_struct(res)
        return (out, arrs) if dump else out

    # ---- post-equaliser stage (ofdm_link_set_post): coloured noise after the equaliser, block-wide renormalisation
    def set_post(self, noise_profile: Optional[np.ndarray] = None, *, recorded_noise: Optional[np.ndarray] = None,
                 z_scale: float = 1.0, measure_power: bool = False) -> None:
        """examples/waterfilling_noise_bump_experiment.py:163-183.  ``noise_profile`` [N]: variance multipliers of the noise
        injected after the equaliser (variance 10^(-snr_db/10) * profile[k]); ``recorded_noise`` [n_symbols, N] complex128:
        the matrix the reference drew (replay mode); ``z_scale``: factor applied before the demapper; ``measure_power``:
        accumulate sum |z|^2 for ``read_z_power``.  All defaults = remove the stage."""
        prof = None if noise_profile is None else np.ascontiguousarray(noise_profile, dtype=np.float64)
        if prof is not None and prof.shape != (self.n_subcarriers,):
            raise ValueError("noise_profile must have one entry per subcarrier")
        rec = None if recorded_noise is None else np.ascontiguousarray(recorded_noise, dtype=np.complex128)
        self._post_keepalive = (prof, rec)        # the library reads the recorded matrix at run time
        post = LinkPost(None if prof is None else prof.ctypes.data, None if rec is None else rec.ctypes.data,
                        float(z_scale), int(bool(measure_power)), 0)
        _check(lib.ofdm_link_set_post(self._h, C.byref(post)))
        self.uses_fast_kernel = bool(lib.ofdm_link_uses_fast_kernel(self._h))

    def read_z_power(self, stream: int = 0):
        """(sum |z|^2, number of equalised values) accumulated since the counters were last reset."""
        total, count = C.c_double(), C.c_uint64()
        _check(lib.ofdm_link_read_z_power(self._h, stream, C.byref(total), C.byref(count)))
        return float(total.value), int(count.value)

    def run_fused_renormalised(self, snr_db: float, noise_sigma: float, n_symbols: int, *, noise_profile=None,
                               seed: int = 0x0FD3, point: int = 0, first_symbol: int = 0, dump: Optional[tuple] = None):
        """The two passes of the reference's block-wide renormalisation (:178-181) over the same Philox streams: the first
        measures the mean power of the equalised, compensated values, the second slices them scaled by 1/sqrt(mean)."""
        self.set_post(noise_profile, measure_power=True)
        self.run_fused(snr_db, noise_sigma, n_symbols, seed=seed, point=point, first_symbol=first_symbol)
        total, count = self.read_z_power()
        avg = total / max(count, 1)
        self.set_post(noise_profile, z_scale=1.0 / np.sqrt(avg) if avg > 1e-12 else 1.0)
        try:
            return self.run_fused(snr_db, noise_sigma, n_symbols, seed=seed, point=point, first_symbol=first_symbol, dump=dump)
        finally:
            self.set_post()

    # ---- a whole SNR sweep in one launch (ofdm_link_run_sweep / _launch_sweep / _read_sweep / _pack_sweep)
    @staticmethod
    def _sweep_arrays(snr_dbs, noise_sigmas):
        snr = np.ascontiguousarray(snr_dbs, dtype=np.float64).reshape(-1)
        sig = np.ascontiguousarray(noise_sigmas, dtype=np.float64).reshape(-1)
        if snr.size == 0 or snr.size != sig.size:
            raise ValueError("a sweep needs one noise sigma per SNR point and at least one point")
        return snr, sig

    def run_sweep(self, snr_dbs, noise_sigmas, n_symbols: int, *, seed: int = 0x0FD3, first_point: int = 0,
                  first_symbol: int = 0):
        """Every SNR point over the same symbol range in one kernel launch; list of LinkCounters, one per point."""
        snr, sig = self._sweep_arrays(snr_dbs, noise_sigmas)
        res = (LinkResult * snr.size)()
        _check(lib.ofdm_link_run_sweep(self._h, snr.size, snr.ctypes.data, sig.ctypes.data, seed, first_point,
                                       first_symbol, n_symbols, res))
        return [LinkCounters.from_struct(r) for r in res]

    def launch_sweep(self, snr_dbs, noise_sigmas, n_symbols: int, *, seed: int = 0x0FD3, first_point: int = 0,
                     first_symbol: int = 0, stream: int = 0) -> int:
        snr, sig = self._sweep_arrays(snr_dbs, noise_sigmas)
        _check(lib.ofdm_link_launch_sweep(self._h, snr.size, snr.ctypes.data, sig.ctypes.data, seed, first_point,
                                          first_symbol, n_symbols, stream))
        return int(snr.size)

    def read_sweep(self, n_points: int, stream: int = 0):
        res = (LinkResult * n_points)()
        _check(lib.ofdm_link_read_sweep(self._h, stream, n_points, res))
        return [LinkCounters.from_struct(r) for r in res]

    def pack_sweep(self, payload_dev: int, rank: int, world: int, stream: int = 0) -> None:
        _check(lib.ofdm_link_pack_sweep(self._h, payload_dev, rank, world, stream))

    # ---- device-pointer entry points (asynchronous on a CUDA stream handle)
    def reset_counters(self, stream: int = 0) -> None:
        _check(lib.ofdm_link_reset_counters(self._h, stream))

    def launch_fused(self, snr_db: float, noise_sigma: float, n_symbols: int, *, seed: int = 0x0FD3, point: int = 0,
                     first_symbol: int = 0, stream: int = 0) -> None:
        _check(lib.ofdm_link_launch_fused(self._h, snr_db, noise_sigma, seed, point, first_symbol, n_symbols, None, stream))

    def launch_replay(self, snr_db: float, bits_dev: int, n_bytes: int, noise_dev: int, noise_dtype: int,
                      n_symbols: int, *, compare_limit_bits: int = 0, stream: int = 0) -> None:
        _check(lib.ofdm_link_launch_replay(self._h, snr_db, bits_dev, n_bytes, noise_dev, noise_dtype, n_symbols,
                                           compare_limit_bits, None, stream))

    def read_result(self, stream: int = 0) -> LinkCounters:
        res = LinkResult()
        _check(lib.ofdm_link_read_result(self._h, stream, C.byref(res)))
        return LinkCounters.from_struct(res)

    def pack_counters(self, payload_row_dev: int, rank: int, world: int, stream: int = 0) -> None:
        _check(lib.ofdm_link_pack_counters(self._h, payload_row_dev, rank, world, stream))

    def debug_tables(self) -> dict:
        """Test hook: the folded fp32 tables of the register-resident kernel as the device holds them."""
        n = self.n_subcarriers
        out = dict(eq=np.zeros((n, 4), np.float32), taps=np.zeros((8, 2), np.float32), taps3=np.zeros((8, 4), np.float32))
        per_sc = dict(level=np.zeros((n, 2), np.float32), masks=np.zeros(n // 4, np.uint32))
        rc = lib.ofdm_link_debug_tables(self._h, out["eq"].ctypes.data, per_sc["level"].ctypes.data,
                                        per_sc["masks"].ctypes.data, out["taps"].ctypes.data, out["taps3"].ctypes.data)
        if rc == 0:
            out.update(per_sc)
        else:   # one order on every subcarrier: no per-subcarrier level / mask tables
            _check(lib.ofdm_link_debug_tables(self._h, out["eq"].ctypes.data, None, None, out["taps"].ctypes.data,
                                              out["taps3"].ctypes.data))
        return out

    @property
    def counters_device_ptr(self) -> int:
        return int(lib.ofdm_link_counters_device_ptr(self._h) or 0)


def _reject_null_channels(taps: np.ndarray) -> None:
    """channel/models.py:41-43: a realisation whose taps are all zero has no unit-energy normalisation."""
    if taps.size and not np.all(np.any(taps != 0, axis=-1)):
        raise ValueError("Impulse response cannot be all zeros.")


def bit_loading_gap(ser: float, scheme: str = "QAM") -> float:
    """The SNR gap of the reference's bit-loading rules: QAM Qinv(ser/4)^2/3 (constellation/models.py:301-304),
    PSK gamma* = Qinv(ser/2)^2 / (2 pi^2) (:462-464).  scipy's norm.isf, like the reference."""
    from scipy.stats import norm
    if scheme == "QAM":
        return float(norm.isf(ser / 4) ** 2 / 3)
    return float(norm.isf(ser / 2) ** 2 / (2 * np.pi ** 2))


def waterfill_bitload_batched(taps: np.ndarray, n_subcarriers: int, snr_db: float, *, ser: float = 1e-3,
                              total_power: Optional[float] = None, scheme: str = "QAM", waterfilling: bool = True,
                              min_order: int = 0, max_order: int = 0, tolerance: float = 1e-8,
                              order_rule: str = "gap", capacity_scaling: float = 1.0):
    """Power allocation + constellation orders for a batch of channel realisations on the GPU.
    taps: [F, L] complex RAW taps.  order_rule "gap" = calculate_bit_loading_order (constellation/models.py:297-321,
    459-474); "capacity" = calculate_constellation_orders (constellation/adaptive.py:271-329, needs min / max order).
    Returns dict(power [F,N], orders [F,N], water_level [F], h_eq [F,N], iterations [F], capacity [F,N])."""
    require_gpu()
    taps = np.ascontiguousarray(np.atleast_2d(taps), dtype=np.complex128)
    f, l = taps.shape
    n = int(n_subcarriers)
    _reject_null_channels(taps)
    desc = WaterfillDesc(n, l, SCHEME[scheme], int(bool(waterfilling)), int(min_order), int(max_order), float(snr_db),
                         float(n if total_power is None else total_power), bit_loading_gap(ser, scheme), float(tolerance),
                         {"gap": 0, "capacity": 1}[order_rule], 0, float(capacity_scaling))
    power = np.empty((f, n), dtype=np.float64)
    orders = np.empty((f, n), dtype=np.int32)
    level = np.empty(f, dtype=np.float64)
    h_eq = np.empty((f, n), dtype=np.complex128)
    iters = np.empty(f, dtype=np.int32)
    cap = np.empty((f, n), dtype=np.float64)
    _check(lib.ofdm_waterfill_bitload_batched(C.byref(desc), taps.ctypes.data, f, power.ctypes.data, orders.ctypes.data,
                                              level.ctypes.data, h_eq.ctypes.data, iters.ctypes.data, cap.ctypes.data))
    if waterfilling and not np.all(np.isfinite(level)):
        # a spectral null: the reference refuses it (power_allocation/models.py:117-121)
        bad = int(np.flatnonzero(~np.isfinite(level))[0])
        raise ValueError(f"All channel gains must be positive, got min={float(np.min(np.abs(h_eq[bad]) ** 2))}")
    return dict(power=power, orders=orders.astype(np.int64), water_level=level, h_eq=h_eq, iterations=iters,
                capacity=cap)


def compare_allocations_batched(taps: np.ndarray, n_subcarriers: int, snr_db: float, *, total_power: float = 1.0):
    """compare_allocations (power_allocation/models.py:296-334) for a batch of channel realisations: uniform vs
    water-filling capacity, summed over the subcarriers of each realisation."""
    uni = waterfill_bitload_batched(taps, n_subcarriers, snr_db, total_power=total_power, waterfilling=False)
    wf = waterfill_bitload_batched(taps, n_subcarriers, snr_db, total_power=total_power, waterfilling=True)
    cu, cw = uni["capacity"].sum(axis=1), wf["capacity"].sum(axis=1)
    return {"uniform_capacity": cu, "waterfilling_capacity": cw, "capacity_gain": cw - cu,
            "capacity_gain_percent": np.where(cu > 0, 100 * (cw - cu) / np.where(cu > 0, cu, 1), 0.0)}


def run_frames(n_subcarriers: int, n_frames: int, symbols_per_frame: int, snr_db: float, *, taps: Optional[np.ndarray] = None,
               n_taps: int = 8, prefix_len: Optional[int] = None, equalizer: str = "MMSE", order: Optional[int] = None,
               waterfilling: bool = True, min_order: int = 4, max_order: int = 256, ser: float = 1e-3, seed: int = 0x0FD3,
               point: int = 0, first_frame: int = 0, device: int = -1, per_frame: bool = True,
               want_orders: bool = True, want_taps: bool = True):
    """A batch of channel realisations in one launch (ofdm_frames_run).  ``taps`` [F, L] complex RAW taps, or None for
    a fresh Rayleigh draw per frame on the device.  ``order`` = one QAM order on every subcarrier; None = per-frame
    gap-rule orders (water-filling or uniform power) bounded to [min_order, max_order].
    Returns dict(total LinkCounters, frames [F] list of LinkCounters, orders [F, N], taps [F, L]); the last three are
    None when not requested (they are the only per-frame device -> host traffic)."""
    require_gpu()
    if taps is not None:
        taps = np.ascontiguousarray(np.atleast_2d(taps), dtype=np.complex128)
        if taps.shape[0] != n_frames:
            raise ValueError("taps must hold one row of raw taps per frame")
        n_taps = taps.shape[1]
        _reject_null_channels(taps)
    n = int(n_subcarriers)
    desc = FramesDesc(n, int(n_taps - 1 if prefix_len is None else prefix_len), EQUALIZER[equalizer], int(n_taps),
                      0 if order is not None else 1, int(order or 0), int(bool(waterfilling)), int(min_order), int(max_order),
                      int(device), float(snr_db), bit_loading_gap(ser, "QAM"))
    total = LinkResult()
    frames = (LinkResult * n_frames)() if per_frame else None
    orders = np.empty((n_frames, n), dtype=np.int32) if want_orders else None
    taps_out = np.empty((n_frames, n_taps), dtype=np.complex128) if want_taps else None
    _check(lib.ofdm_frames_run(C.byref(desc), None if taps is None else taps.ctypes.data, n_frames, symbols_per_frame, seed,
                               point, first_frame, C.byref(total), frames, None if orders is None else orders.ctypes.data,
                               None if taps_out is None else taps_out.ctypes.data))
    return dict(total=LinkCounters.from_struct(total),
                frames=[LinkCounters.from_struct(r) for r in frames] if per_frame else None,
                orders=None if orders is None else orders.astype(np.int64), taps=taps_out)


def frames_debug_tables(n_subcarriers: int, taps: np.ndarray, snr_db: float, *, prefix_len: Optional[int] = None,
                        equalizer: str = "MMSE", order: Optional[int] = None, waterfilling: bool = True,
                        min_order: int = 4, max_order: int = 256, ser: float = 1e-3, device: int = -1) -> dict:
    """Test hook: the per-frame fp32 tables ofdm_frames_run builds on the device for these raw taps [F, L]."""
    require_gpu()
    taps = np.ascontiguousarray(np.atleast_2d(taps), dtype=np.complex128)
    f, l = taps.shape
    n = int(n_subcarriers)
    desc = FramesDesc(n, int(l - 1 if prefix_len is None else prefix_len), EQUALIZER[equalizer], int(l),
                      0 if order is not None else 1, int(order or 0), int(bool(waterfilling)), int(min_order), int(max_order),
                      int(device), float(snr_db), bit_loading_gap(ser, "QAM"))
    hf = int(lib.ofdm_frames_header_floats())
    out = dict(eq=np.zeros((f, n, 4), np.float32), level=np.zeros((f, n, 2), np.float32),
               masks=np.zeros((f, n // 4), np.uint32), hdr=np.zeros((f, hf), np.float32))
    _check(lib.ofdm_frames_debug_tables(C.byref(desc), taps.ctypes.data, f, 0, 0, out["eq"].ctypes.data,
                                        out["level"].ctypes.data, out["masks"].ctypes.data, out["hdr"].ctypes.data))
    out.update(taps=out["hdr"][:, :16].reshape(f, 8, 2), taps3=out["hdr"][:, 16:48].reshape(f, 8, 4),
               sigma=out["hdr"][:, 48], mmse_c=out["hdr"][:, 49])
    return out


def measure_fp32_tflops(iters: int = 4096) -> float:
    require_gpu()
    return float(lib.ofdm_b200_measure_fp32_tflops(iters))


def launch_count() -> int:
    return int(lib.ofdm_b200_launch_count())
