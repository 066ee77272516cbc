"""CPU oracle for the per-OFDM-symbol link chain  --  TEST INFRASTRUCTURE ONLY.

This file is a NumPy fp64 *restatement* of the algorithm the reference implements in
``src/ofdm_based_systems`` (paths below are relative to the reference checkout).  It exists so that
the CUDA path can be checked against something that runs anywhere (the reference itself is Python
and does not travel to the GPU box).  Nothing under ``ofdm-based-systems_b200/`` imports it; only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` do.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the live, unmodified reference in the build
container, drives it with seeded bits and recorded noise and stores inputs + every intermediate in
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function here against those
files and against the known-answer vectors the reference's own tests hold (Gray table, ZP
overlap-add, tail-bit masking, MSB-first bit order, ZF exact division, bit-loading tables,
water-filling KATs).

Each function cites the reference lines it follows.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np
from scipy.stats import norm

QAM, PSK = "QAM", "PSK"
PREFIX_NONE, PREFIX_CYCLIC, PREFIX_ZERO = "NONE", "CYCLIC", "ZERO"
EQ_NONE, EQ_ZF, EQ_MMSE = "NONE", "ZF", "MMSE"
MOD_OFDM, MOD_SC = "OFDM", "SC-OFDM"


# --------------------------------------------------------------------------------------------
# bits  (bits_generation/models.py:27-55, simulation/models.py:59-69)
# --------------------------------------------------------------------------------------------
def generate_bits(num_bits: int, rng: np.random.Generator) -> bytes:
    """``RandomBitsGenerator.generate_bits`` (bits_generation/models.py:27-55): ceil(n/8) generator
    bytes, unused low bits of the last byte cleared."""
    num_bytes = math.ceil(num_bits / 8)
    raw = bytearray(rng.bytes(num_bytes))
    keep = num_bits % 8
    if keep > 0 and num_bytes > 0:
        raw[-1] &= (0xFF << (8 - keep)) & 0xFF
    return bytes(raw)


def unpack_bits(data: bytes) -> np.ndarray:
    """``read_bits_from_stream`` (simulation/models.py:59-69): MSB-first, 8 bits per byte."""
    return np.unpackbits(np.frombuffer(data, dtype=np.uint8), bitorder="big").astype(np.int64)


def pack_bits(bits: np.ndarray) -> bytes:
    """Byte packing at the end of ``decode`` (constellation/models.py:272-292): MSB-first, the last
    partial byte is left-aligned and zero padded."""
    return np.packbits(np.asarray(bits, dtype=np.uint8), bitorder="big").tobytes()


# --------------------------------------------------------------------------------------------
# constellations  (constellation/models.py:70-109, 180-218, 356-380)
# --------------------------------------------------------------------------------------------
def gray(x):
    """``GrayWordCoder.gray_table`` (constellation/models.py:75-77)."""
    return x ^ (x >> 1)


def qam_constellation(order: int) -> np.ndarray:
    """``QAMConstellationMapper.generate_constellation`` (constellation/models.py:180-218) followed by
    ``GrayWordCoder.reorder_constellation`` (:94-109), restated literally (loop form)."""
    side = int(np.sqrt(order))
    if side * side != order:
        raise ValueError("Order must be a perfect square (e.g., 4, 16, 64).")
    levels = np.arange(-side + 1, side, 2)
    natural = [complex(i, q) for q in levels[::-1] for i in levels]
    const = np.zeros(order, dtype=np.complex128)
    for b in range(order):
        const[b] = natural[gray(b)]
    out = np.zeros_like(const)
    for r in range(side):
        row = const[r * side:(r + 1) * side]
        out[r * side:(r + 1) * side] = row[::-1] if r % 2 == 1 else row
    out /= np.sqrt(np.mean(np.abs(out) ** 2))
    return out


def qam_constellation_closed_form(order: int) -> np.ndarray:
    """Closed form of the table above: hi = b >> m, lo = b & (s-1),
    I = -(s-1) + 2*gray(lo), Q = (s-1) - 2*gray(hi), divided by sqrt(2(M-1)/3)."""
    side = int(np.sqrt(order))
    m = int(np.log2(side))
    b = np.arange(order)
    hi, lo = b >> m, b & (side - 1)
    pts = (-(side - 1) + 2 * gray(lo)) + 1j * ((side - 1) - 2 * gray(hi))
    return pts / np.sqrt(np.mean(np.abs(pts) ** 2))


def psk_constellation(order: int) -> np.ndarray:
    """``PSKConstellationMapper.generate_constellation`` (constellation/models.py:356-380):
    constellation[gray(k)] = exp(j*2*pi*k/M)."""
    bps = np.log2(order)
    if bps != int(bps) or order < 2:
        raise ValueError("PSK order must be a power of 2 (e.g., 2, 4, 8, 16).")
    k = np.arange(order)
    const = np.zeros(order, dtype=np.complex128)
    const[gray(k)] = np.exp(1j * (2 * np.pi * k / order))
    return const


def constellation(order: int, scheme: str = QAM) -> np.ndarray:
    return qam_constellation(order) if scheme == QAM else psk_constellation(order)


def bits_per_symbol(order: int) -> int:
    """``bits_per_symbol`` property (constellation/models.py:171-172)."""
    return int(np.log2(order)) if order > 0 else 0


def nn_classify(const: np.ndarray, symbols: np.ndarray, chunk: int = 1 << 15) -> np.ndarray:
    """``NNClassifier.classify`` (constellation/models.py:19-27): argmin over |z - c|, first index on
    ties.  Returns the INDEX (== label, because the table is indexed by label)."""
    symbols = np.asarray(symbols, dtype=np.complex128).ravel()
    out = np.empty(symbols.shape[0], dtype=np.int64)
    for s in range(0, symbols.shape[0], chunk):
        d = np.abs(symbols[s:s + chunk, None] - const[None, :])
        out[s:s + chunk] = np.argmin(d, axis=1)
    return out


# --------------------------------------------------------------------------------------------
# fixed-order map / demap  (constellation/models.py:220-295, 382-457)
# --------------------------------------------------------------------------------------------
def labels_from_bits(data: bytes, bps: int) -> np.ndarray:
    """Bit unpacking + zero padding + label = sum(bit_i << (bps-1-i)) (constellation/models.py:226-243)."""
    bits = unpack_bits(data)
    if bits.size % bps != 0:
        bits = np.concatenate([bits, np.zeros(bps - bits.size % bps, dtype=np.int64)])
    return bits.reshape(-1, bps).dot(1 << np.arange(bps - 1, -1, -1))


def encode_fixed(data: bytes, order: int, scheme: str = QAM) -> Tuple[np.ndarray, np.ndarray]:
    """``encode`` (constellation/models.py:220-249 / 382-411). Returns (symbols, labels)."""
    labels = labels_from_bits(data, bits_per_symbol(order))
    return constellation(order, scheme)[labels], labels


def bits_from_labels(labels: np.ndarray, bps: int) -> np.ndarray:
    """label -> bits MSB-first (constellation/models.py:267-270)."""
    shifts = np.arange(bps - 1, -1, -1)
    return ((np.asarray(labels)[:, None] >> shifts[None, :]) & 1).reshape(-1)


def decode_fixed(symbols: np.ndarray, order: int, scheme: str = QAM) -> Tuple[bytes, np.ndarray]:
    """``decode`` (constellation/models.py:251-295 / 413-457). Returns (packed bytes, labels)."""
    labels = nn_classify(constellation(order, scheme), symbols)
    return pack_bits(bits_from_labels(labels, bits_per_symbol(order))), labels


# --------------------------------------------------------------------------------------------
# adaptive (per-subcarrier) map / demap  (constellation/adaptive.py:130-265)
# --------------------------------------------------------------------------------------------
def adaptive_bits_per_subcarrier(orders: np.ndarray) -> np.ndarray:
    """constellation/adaptive.py:77-80."""
    return np.array([int(np.log2(o)) if o > 0 else 0 for o in orders], dtype=np.int64)


def encode_adaptive(data: bytes, orders: np.ndarray, scheme: str = QAM) -> Tuple[np.ndarray, np.ndarray]:
    """``AdaptiveConstellationMapper.encode`` (constellation/adaptive.py:130-201): bits are consumed
    OFDM-symbol-major, subcarrier-minor, bps_k bits each; order-0 subcarriers carry 0+0j.
    Returns (symbols[S*N], labels[S, N] with -1 on inactive subcarriers)."""
    orders = np.asarray(orders, dtype=np.int64)
    bps = adaptive_bits_per_subcarrier(orders)
    per_sym = int(bps.sum())
    if per_sym == 0:
        raise ValueError("No active subcarriers (all orders are zero)")
    bits = unpack_bits(data)
    if bits.size % per_sym != 0:
        raise ValueError(
            f"Bits length ({bits.size}) must be multiple of bits_per_symbol ({per_sym})")
    n_sym, n_sc = bits.size // per_sym, orders.size
    bits = bits.reshape(n_sym, per_sym)
    offs = np.concatenate([[0], np.cumsum(bps)])
    labels = np.full((n_sym, n_sc), -1, dtype=np.int64)
    out = np.zeros((n_sym, n_sc), dtype=np.complex128)
    tables: Dict[int, np.ndarray] = {int(o): constellation(int(o), scheme) for o in np.unique(orders) if o > 0}
    for k in range(n_sc):
        if bps[k] == 0:
            continue
        chunk = bits[:, offs[k]:offs[k + 1]]
        lab = chunk.dot(1 << np.arange(bps[k] - 1, -1, -1))
        labels[:, k] = lab
        out[:, k] = tables[int(orders[k])][lab]
    return out.reshape(-1), labels


def decode_adaptive(symbols: np.ndarray, orders: np.ndarray, scheme: str = QAM) -> Tuple[bytes, np.ndarray]:
    """``AdaptiveConstellationMapper.decode`` (constellation/adaptive.py:203-265): per (symbol,
    subcarrier) NN demap with that subcarrier's table; inactive subcarriers skipped; a trailing
    partial byte is DROPPED (:259-263).  Returns (bytes, labels[S, N] with -1 on inactive)."""
    orders = np.asarray(orders, dtype=np.int64)
    n_sc = orders.size
    symbols = np.asarray(symbols, dtype=np.complex128).ravel()
    if symbols.size % n_sc != 0:
        raise ValueError(
            f"Symbols length ({symbols.size}) must be multiple of num_subcarriers ({n_sc})")
    z = symbols.reshape(-1, n_sc)
    bps = adaptive_bits_per_subcarrier(orders)
    labels = np.full(z.shape, -1, dtype=np.int64)
    cols = []
    for k in range(n_sc):
        if bps[k] == 0:
            continue
        lab = nn_classify(constellation(int(orders[k]), scheme), z[:, k])
        labels[:, k] = lab
        shifts = np.arange(bps[k] - 1, -1, -1)
        cols.append((lab[:, None] >> shifts[None, :]) & 1)
    bits = np.concatenate(cols, axis=1).reshape(-1) if cols else np.zeros(0, dtype=np.int64)
    full = (bits.size // 8) * 8
    return pack_bits(bits[:full]), labels


# --------------------------------------------------------------------------------------------
# prefix / modulation  (prefix/models.py:29-113, modulation/models.py:19-91)
# --------------------------------------------------------------------------------------------
def add_prefix(rows: np.ndarray, prefix_len: int, prefix_type: str) -> np.ndarray:
    """Row-wise ``add_prefix``: CP prepends the last P samples (prefix/models.py:34-44), ZP appends P
    zeros (:60-69), NONE is the identity (:109-110)."""
    if prefix_type == PREFIX_NONE or (prefix_type == PREFIX_CYCLIC and prefix_len == 0):
        return rows
    if prefix_type == PREFIX_CYCLIC:
        return np.concatenate([rows[:, rows.shape[1] - prefix_len:], rows], axis=1)
    return np.concatenate([rows, np.zeros((rows.shape[0], prefix_len), dtype=rows.dtype)], axis=1)


def remove_prefix(rows: np.ndarray, prefix_len: int, prefix_type: str) -> np.ndarray:
    """Row-wise ``remove_prefix``: CP strip (prefix/models.py:46-52); ZP overlap-add, i.e. the
    [I_N | I_P;0] matrix product of :87-101 written as r[n] = y[n] + y[n+N] for n < P."""
    if prefix_type == PREFIX_NONE:
        return rows
    if prefix_type == PREFIX_CYCLIC:
        return rows[:, prefix_len:]
    n = rows.shape[1] - prefix_len
    out = rows[:, :n].copy()
    out[:, :prefix_len] += rows[:, n:]
    return out


def modulate(parallel: np.ndarray, prefix_len: int, prefix_type: str, modulator: str = MOD_OFDM) -> np.ndarray:
    """``OFDMModulator.modulate`` (modulation/models.py:27-39): ortho IFFT along axis 1 + prefix;
    ``SingleCarrierOFDMModulator.modulate`` (:66-72): prefix only."""
    x = np.fft.ifft(parallel, axis=1, norm="ortho") if modulator == MOD_OFDM else parallel
    return add_prefix(x, prefix_len, prefix_type)


def papr_db(tx: np.ndarray) -> float:
    """simulation/models.py:519-522: over every tx sample, prefix included."""
    p = np.abs(tx) ** 2
    avg = np.mean(p)
    return float(10 * np.log10(np.max(p) / avg)) if avg > 0 else float("inf")


# --------------------------------------------------------------------------------------------
# channel + noise  (channel/models.py:37-62, noise/models.py:12-22)
# --------------------------------------------------------------------------------------------
def normalize_taps(h: np.ndarray) -> np.ndarray:
    """``ChannelModel.normalize_impulse_response`` (channel/models.py:37-44)."""
    h = np.asarray(h, dtype=np.complex128)
    power = np.sum(np.abs(h) ** 2)
    if power == 0:
        raise ValueError("Impulse response cannot be all zeros.")
    return h / np.sqrt(power)


def channel_convolve(serial: np.ndarray, h_norm: np.ndarray) -> np.ndarray:
    """First half of ``ChannelModel.transmit`` (channel/models.py:52-55): causal linear convolution
    over the WHOLE serial stream, truncated to the input length."""
    return np.convolve(serial, h_norm, mode="full")[: serial.shape[0]].astype(np.complex128)


def awgn_noise(signal: np.ndarray, snr_db: float, normal_re: np.ndarray, normal_im: np.ndarray) -> np.ndarray:
    """``AWGNoiseModel.add_noise`` (noise/models.py:13-22) with the two standard-normal draws supplied
    by the caller (the reference draws the real array first): returns the noise that is ADDED."""
    signal_power = np.mean(np.abs(signal) ** 2)
    noise_power = signal_power / (10 ** (snr_db / 10))
    return np.sqrt(noise_power / 2) * (normal_re + 1j * normal_im)


# --------------------------------------------------------------------------------------------
# equalisers + demodulation  (equalization/models.py:22-68, modulation/models.py:41-55, 74-91)
# --------------------------------------------------------------------------------------------
def equalize_rows(Y: np.ndarray, H: Optional[np.ndarray], eq: str, snr_db: Optional[float]) -> np.ndarray:
    """Row-wise ``equalize``.  ZF: Y / where(H == 0, 1e-10, H) (equalization/models.py:33-35).
    MMSE: per ROW sigma2 = mean|Y_row|^2 / snr_lin / mean|H|^2 (:39-49, inf if mean|H|^2 == 0),
    Z = Y * conj(H) / (|H|^2 + sigma2) (:59-63).  NONE: identity (:66-68)."""
    if eq == EQ_NONE:
        return Y
    if eq == EQ_ZF:
        return Y / np.where(H == 0, 1e-10, H)[None, :]
    if snr_db is None:
        raise ValueError("SNR in dB must be provided to calculate noise variance.")
    gain = np.mean(np.abs(H) ** 2)
    sig = np.mean(np.abs(Y) ** 2, axis=1)
    sigma2 = np.full_like(sig, np.inf) if gain == 0 else (sig / (10 ** (snr_db / 10))) / gain
    filt = np.conj(H)[None, :] / ((np.abs(H) ** 2)[None, :] + sigma2[:, None])
    return Y * filt


def demodulate(rx_parallel: np.ndarray, n_sc: int, prefix_len: int, prefix_type: str, eq: str,
               H_eq: Optional[np.ndarray], snr_db: Optional[float], modulator: str = MOD_OFDM,
               return_freq: bool = False):
    """``OFDMModulator.demodulate`` (modulation/models.py:41-55): strip, ortho FFT, equalise per row;
    SC-OFDM (:74-91) adds an ortho IFFT after the equaliser."""
    r = remove_prefix(rx_parallel, prefix_len, prefix_type)
    Y = np.fft.fft(r, n=n_sc, axis=1, norm="ortho")
    Z = equalize_rows(Y, H_eq, eq, snr_db)
    out = np.fft.ifft(Z, n=n_sc, axis=1, norm="ortho") if modulator == MOD_SC else Z
    return (out, Y, Z) if return_freq else out


# --------------------------------------------------------------------------------------------
# power allocation + bit loading  (power_allocation/models.py:61-69,140-225; constellation/models.py:297-321,459-474)
# --------------------------------------------------------------------------------------------
def uniform_power(total_power: float, n_sc: int) -> np.ndarray:
    """``UniformPowerAllocation.allocate`` (power_allocation/models.py:61-69)."""
    return np.full(n_sc, total_power / n_sc, dtype=np.float64)


def waterfilling(total_power: float, gains: np.ndarray, noise_power: float, tol: float = 1e-8,
                 return_info: bool = False):
    """``WaterfillingPowerAllocation.allocate`` + ``_find_water_level`` (power_allocation/models.py:140-225):
    floor_k = N0 / (g_k * N)  [the extra /N is the reference's, :161]; bisection on mu in
    [0, P_tot + max floor], <= 100 iterations, stop when |sum max(0, mu - floor) - P_tot| < tol;
    P = max(0, mu - floor) rescaled to sum to P_tot."""
    gains = np.asarray(gains, dtype=np.float64)
    floor = noise_power / (gains * len(gains))
    lo, hi = 0.0, total_power + np.max(floor)
    mu = (lo + hi) / 2
    iters = 0
    for iters in range(1, 101):
        mu = (lo + hi) / 2
        s = np.sum(np.maximum(0, mu - floor))
        if np.abs(s - total_power) < tol:
            break
        if s < total_power:
            lo = mu
        else:
            hi = mu
    power = np.maximum(0, mu - floor)
    s = np.sum(power)
    if s > 0:
        power = power * (total_power / s)
    return (power, mu, iters) if return_info else power


def reported_water_level(power: np.ndarray, gains: np.ndarray, noise_power: float) -> float:
    """simulation/models.py:311-313 / 493-495: mean(P_k + N0/g_k) over P_k > 1e-10 (no /N here)."""
    lvl = power + noise_power / gains
    return float(np.mean(lvl[power > 1e-10]))


def bit_loading_qam(ser: float, snr: float) -> int:
    """``QAMConstellationMapper.calculate_bit_loading_order`` (constellation/models.py:297-321)."""
    gamma = (1 / 3) * (norm.isf(ser / 4) ** 2)
    bits = int(np.round(np.log2(1 + (snr / gamma))))
    if bits % 2 != 0:
        bits -= 1
    return 0 if bits <= 0 else 2 ** bits


def bit_loading_psk(ser: float, snr: float) -> int:
    """``PSKConstellationMapper.calculate_bit_loading_order`` (constellation/models.py:459-474)."""
    q_inv = norm.isf(ser / 2)
    gamma_star = (q_inv ** 2) / (2 * (np.pi ** 2))
    with np.errstate(divide="ignore", invalid="ignore"):
        gamma = (np.sqrt(snr * gamma_star)) / (1 - np.sqrt(gamma_star / (snr + 1e-10)))
        bits = int(np.floor(np.log2(1 + snr / (gamma + 1e-10)) + 1e-10))
    return 0 if bits <= 0 else 2 ** bits


def bit_loading_orders(power: np.ndarray, gains: np.ndarray, noise_power: float, ser: float,
                       scheme: str = QAM) -> np.ndarray:
    """simulation/models.py:337-352: snr_k = P_k * g_k / N0, then the gap rule per subcarrier."""
    f = bit_loading_qam if scheme == QAM else bit_loading_psk
    return np.array([f(ser, p * g / noise_power) for p, g in zip(power, gains)], dtype=np.int64)


def capacity_per_subcarrier(power, gains, noise_power):
    """``calculate_capacity_per_subcarrier`` (power_allocation/models.py:264-293)."""
    return np.log2(1 + power * gains / noise_power + 1e-12)


def capacity(power, gains, noise_power) -> float:
    """``calculate_capacity`` (power_allocation/models.py:228-261)."""
    return float(np.sum(capacity_per_subcarrier(power, gains, noise_power)))


def shannon_orders(cap: np.ndarray, min_order: int, max_order: int, scaling: float, scheme: str = QAM) -> np.ndarray:
    """``calculate_constellation_orders`` (constellation/adaptive.py:271-329)."""
    b = np.clip(cap * scaling, 0, np.log2(max_order))
    b = (b // 2 * 2) if scheme == QAM else np.floor(b)
    b = np.where(b < np.log2(min_order), 0, b)
    return np.where(b > 0, 2 ** b, 0).astype(np.int64)


# --------------------------------------------------------------------------------------------
# the whole link  (simulation/models.py:226-606 minus printing / plotting)
# --------------------------------------------------------------------------------------------
@dataclass
class LinkSetup:
    """What ``Simulation.run()`` derives before the hot loop (simulation/models.py:226-410)."""
    n_sc: int
    taps_raw: np.ndarray                  # as given (custom .npy or the 4-tap default)
    snr_db: float
    order: int = 16                       # fixed mode
    scheme: str = QAM
    modulator: str = MOD_OFDM
    prefix_type: str = PREFIX_CYCLIC
    prefix_ratio: float = 1.0
    eq: str = EQ_MMSE
    awgn: bool = True
    orders: Optional[np.ndarray] = None   # adaptive mode: per-subcarrier orders (len n_sc)
    prefix_len_override: Optional[int] = None  # component-pipeline callers pick P directly
    # applied power loading, which Simulation.run() never does (simulation/models.py:508) but the reference's
    # experiments do: tx amplitudes sqrt(P_k) (examples/overview.py:142, examples/waterfilling_noise_bump_experiment.py:148)
    # and the receiver's compensation 1/sqrt(P_k) on the equalised subcarriers (waterfilling_noise_bump_experiment.py:165-169)
    amp: Optional[np.ndarray] = None
    rx_gain: Optional[np.ndarray] = None
    taps_chan: np.ndarray = field(init=False)
    H_eq: np.ndarray = field(init=False)
    prefix_len: int = field(init=False)

    def __post_init__(self):
        self.taps_raw = np.asarray(self.taps_raw, dtype=np.complex128)
        self.taps_chan = normalize_taps(self.taps_raw)                  # channel/models.py:14-16
        order_l = len(self.taps_chan) - 1                                # channel/models.py:22-24
        p = int(self.prefix_ratio * order_l)                             # simulation/models.py:251
        if self.prefix_type == PREFIX_NONE:
            p = 0                                                        # :252-253
        self.prefix_len = p if self.prefix_len_override is None else int(self.prefix_len_override)
        self.H_eq = np.fft.fft(self.taps_raw, self.n_sc)                 # :263-266 (RAW taps, quirk Q3)

    @property
    def adaptive(self) -> bool:
        return self.orders is not None


DEFAULT_TAPS = np.array([                                                # simulation/models.py:237-245
    7.767824138452235072e-01 + 4.560896742466611919e-01j,
    -6.669848996328063551e-02 + 2.839935704583463338e-01j,
    1.398968327715586490e-01 - 1.591963958343969865e-01j,
    2.229949514514480494e-02 + 2.409945439452868821e-01j,
], dtype=np.complex128)


def adaptive_setup(n_sc: int, taps_raw: np.ndarray, snr_db: float, ser: float, scheme: str = QAM,
                   waterfill: bool = True):
    """simulation/models.py:278-352: gains from RAW taps, N0 = 10^(-snr/10), P_tot = N,
    power allocation, gap-rule orders.  Returns (orders, power, water_level or None)."""
    gains = np.abs(np.fft.fft(np.asarray(taps_raw, dtype=np.complex128), n_sc)) ** 2
    n0 = 10 ** (-snr_db / 10)
    if waterfill:
        power = waterfilling(n_sc, gains, n0)
        level = reported_water_level(power, gains, n0)
    else:
        power, level = uniform_power(n_sc, n_sc), None
    return bit_loading_orders(power, gains, n0, ser, scheme), power, level


def run_link(setup: LinkSetup, tx_bytes: bytes, total_bits: int, noise: Optional[np.ndarray] = None,
             normals: Optional[Tuple[np.ndarray, np.ndarray]] = None, post_noise: Optional[np.ndarray] = None,
             renormalise: bool = False) -> Dict[str, object]:
    """The hot path ``Simulation.run()`` executes between simulation/models.py:454 and :606, with the
    bits supplied by the caller and the noise either supplied already scaled (``noise``: complex
    array over the serial stream = replay) or as the two standard-normal arrays the reference would
    draw (``normals``), or absent (no-noise model).
    ``post_noise`` ([S, N] complex, already scaled) and ``renormalise`` restate the post-equaliser stage of
    examples/waterfilling_noise_bump_experiment.py:163-183: coloured noise added to the equalised subcarriers, then the
    receiver gains, then the whole block divided by the square root of its mean power."""
    n, p = setup.n_sc, setup.prefix_len
    if setup.adaptive:
        symbols, tx_labels = encode_adaptive(tx_bytes, setup.orders, setup.scheme)
    else:
        symbols, tx_labels = encode_fixed(tx_bytes, setup.order, setup.scheme)
    if symbols.size % n != 0:
        raise ValueError("Length of data must be divisible by number of streams.")  # serial_parallel/models.py:13
    parallel = symbols.reshape(-1, n)                                                 # :471
    if setup.amp is not None:
        parallel = parallel * np.asarray(setup.amp, dtype=np.float64)[None, :]        # noise_bump_experiment.py:148
    tx = modulate(parallel, p, setup.prefix_type, setup.modulator)                    # :514
    papr = papr_db(tx)                                                                # :519-524
    serial = tx.reshape(-1)                                                           # :529
    conv = channel_convolve(serial, setup.taps_chan)                                  # :538
    if noise is None and normals is not None and setup.awgn:
        noise = awgn_noise(conv, setup.snr_db, normals[0], normals[1])
    rx = conv + noise if (noise is not None and setup.awgn) else conv
    rx_par = rx.reshape(-1, n + p)                                                    # :546-548
    out, Y, Zf = demodulate(rx_par, n, p, setup.prefix_type, setup.eq, setup.H_eq, setup.snr_db,
                            setup.modulator, return_freq=True)                        # :554
    if post_noise is not None:
        out = out + np.asarray(post_noise).reshape(out.shape)                         # noise_bump_experiment.py:163-171
    if setup.rx_gain is not None:
        if setup.modulator != MOD_OFDM:
            raise ValueError("receiver power compensation is per subcarrier: OFDM modulator only")
        out = out * np.asarray(setup.rx_gain, dtype=np.float64)[None, :]              # noise_bump_experiment.py:173-176
    z = out.reshape(-1)
    z_avg_power = float(np.mean(np.abs(z) ** 2)) if z.size else 0.0                   # :178
    if renormalise and z_avg_power > 1e-12:
        z = z / np.sqrt(z_avg_power)                                                  # :179-181
    if setup.adaptive:
        rx_bytes, rx_labels = decode_adaptive(z, setup.orders, setup.scheme)          # :591
    else:
        rx_bytes, rx_labels = decode_fixed(z, setup.order, setup.scheme)
    tb, rb = unpack_bits(tx_bytes), unpack_bits(rx_bytes)                             # :457, :593
    m = min(tb.size, rb.size)
    bit_errors = int(np.sum(tb[:m] != rb[:m]))                                        # :597 (zip truncates)
    # SER (:604-606): re-encode the decisions, compare constellation points
    if setup.adaptive:
        recoded, _ = encode_adaptive(rx_bytes, setup.orders, setup.scheme)
    else:
        recoded, _ = encode_fixed(rx_bytes, setup.order, setup.scheme)
    symbol_errors = int(np.sum(symbols != recoded))
    return dict(symbols=symbols, tx_labels=tx_labels, tx=tx, conv=conv, noise=noise, rx=rx, Y=Y, Zf=Zf,
                received_symbols=z, rx_bytes=rx_bytes, rx_labels=rx_labels, bit_errors=bit_errors,
                symbol_errors=symbol_errors, total_bits=total_bits,
                bit_error_rate=(bit_errors / total_bits if total_bits > 0 else 0.0),
                symbol_error_rate=(symbol_errors / symbols.size if symbols.size else 0.0),
                papr_db=papr, z_avg_power=z_avg_power, noise_power=(None if noise is None else float(np.mean(np.abs(conv) ** 2) / 10 ** (setup.snr_db / 10))))


# --------------------------------------------------------------------------------------------
# decision-boundary distance (used by the parity tests to define "away from boundaries")
# --------------------------------------------------------------------------------------------
def qam_boundary_distance(z: np.ndarray, order: int) -> np.ndarray:
    """Distance of each point from the nearest slicer threshold of square M-QAM, in constellation
    units (thresholds at even multiples of 1/k, k = sqrt(2(M-1)/3); outer cells are open)."""
    side = int(np.sqrt(order))
    k = np.sqrt(2 * (order - 1) / 3)
    d = np.full(z.shape, np.inf)
    for comp in (z.real * k, z.imag * k):
        thr = np.arange(-(side - 2), side - 1, 2, dtype=np.float64)     # interior thresholds
        if thr.size:
            d = np.minimum(d, np.min(np.abs(comp[..., None] - thr), axis=-1) / k)
    return d


def psk_boundary_distance(z: np.ndarray, order: int) -> np.ndarray:
    """Angular distance (radians, scaled by |z|) from the nearest PSK decision boundary."""
    ang = np.angle(z) * order / (2 * np.pi)
    frac = np.abs((ang - 0.5) - np.round(ang - 0.5))
    return frac * (2 * np.pi / order) * np.abs(z)


# --------------------------------------------------------------------------------------------
# Simulation.run() including its set-up  (simulation/models.py:214-606)
# --------------------------------------------------------------------------------------------
def simulate(*, num_bits=None, num_symbols=None, num_subcarriers=64, constellation_order=16,
             constellation_scheme=QAM, modulator_type=MOD_OFDM, prefix_scheme=PREFIX_CYCLIC,
             prefix_length_ratio=1.0, equalizator_type=EQ_MMSE, snr_db=20.0, noise_scheme="AWGN",
             power_allocation_type="UNIFORM", adaptive_modulation_mode="FIXED",
             desired_symbol_error_rate=1e-3, channel_impulse_response=None,
             bit_rng: Optional[np.random.Generator] = None, noise_rng=None) -> Dict[str, object]:
    """``Simulation.run()`` with explicit RNGs: ``bit_rng`` plays the shared default Generator of
    RandomBitsGenerator (bits_generation/models.py:24), ``noise_rng`` the global legacy RNG of
    noise/models.py:19-21 (an ``np.random.RandomState`` or the ``np.random`` module)."""
    if num_bits is None and num_symbols is None:
        raise ValueError("Either num_bits or num_symbols must be provided.")
    if num_bits is not None and num_symbols is not None:
        raise ValueError("Only one of num_bits or num_symbols should be provided.")
    taps = DEFAULT_TAPS if channel_impulse_response is None else np.asarray(channel_impulse_response)
    n = num_subcarriers
    adaptive = adaptive_modulation_mode == "CAPACITY_BASED"
    gains = np.abs(np.fft.fft(taps, n)) ** 2                                          # :277-278
    n0 = 10 ** (-snr_db / 10)                                                         # :279
    water_level = None
    if adaptive:                                                                      # :289-395
        wf = power_allocation_type == "WATERFILLING"
        orders, power, water_level = adaptive_setup(n, taps, snr_db, desired_symbol_error_rate,
                                                    constellation_scheme, waterfill=wf)
        bps = adaptive_bits_per_subcarrier(orders)
        if num_symbols is not None:
            n_ofdm = num_symbols
        else:
            if int(bps.sum()) == 0:
                raise ValueError("All subcarriers have zero order - cannot transmit data")
            n_ofdm = num_bits // int(bps.sum())
        total_bits = int(bps.sum() * n_ofdm)
    else:                                                                             # :397-410
        orders = None
        total_bits = num_bits if num_symbols is None else num_symbols * int(np.log2(constellation_order))
        if power_allocation_type == "WATERFILLING":                                   # :483-495
            power = waterfilling(1.0, gains, n0)
        else:
            power = uniform_power(1.0, n)
    setup = LinkSetup(n_sc=n, taps_raw=taps, snr_db=snr_db, order=constellation_order, scheme=constellation_scheme,
                      modulator=modulator_type, prefix_type=prefix_scheme, prefix_ratio=prefix_length_ratio,
                      eq=equalizator_type, awgn=(noise_scheme == "AWGN"), orders=orders)
    bit_rng = bit_rng if bit_rng is not None else np.random.default_rng()
    tx_bytes = generate_bits(total_bits, bit_rng)
    normals = None
    if setup.awgn:
        noise_rng = noise_rng if noise_rng is not None else np.random
        if setup.adaptive:
            n_ofdm_sym = total_bits // int(adaptive_bits_per_subcarrier(orders).sum())
        else:
            bps1 = bits_per_symbol(constellation_order)
            n_ofdm_sym = -(-(8 * math.ceil(total_bits / 8)) // bps1) // n
        shape = (n_ofdm_sym * (n + setup.prefix_len),)
        normals = (noise_rng.normal(size=shape), noise_rng.normal(size=shape))
    out = run_link(setup, tx_bytes, total_bits, normals=normals)
    out.update(constellation_order_per_subcarrier=(orders if adaptive else np.full(n, constellation_order, dtype=np.int64)),
               allocated_power=power, water_level=water_level, setup=setup, tx_bytes=tx_bytes)
    return out
