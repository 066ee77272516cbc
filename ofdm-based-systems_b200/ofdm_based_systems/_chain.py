"""Host-side fp64 functional core of the link chain.

The reference spreads these few-line operations over small strategy classes (serial_parallel, prefix, channel,
noise, equalization); here they are plain functions, shared by the API shells that keep the reference's class
names and by the planner in ``simulation/models.py``.  They configure and document the CUDA link - the hot loop
itself runs only in libofdm_b200.so.  Error texts are part of the reference's contract (its tests match them).
"""
from __future__ import annotations

import numpy as np


def _need_dims(array, ndim: int, message: str) -> None:
    if array.ndim != ndim:
        raise ValueError(message)


# ------------------------------------------------------------------ serial <-> parallel (serial_parallel/models.py:7-21)
def split_streams(data, num_streams: int):
    _need_dims(data, 1, "Input data must be a 1D array.")
    if num_streams <= 0:
        raise ValueError("Number of streams must be a positive integer.")
    rows, rest = divmod(len(data), num_streams)
    if rest:
        raise ValueError("Length of data must be divisible by number of streams.")
    return data.reshape(rows, num_streams)


def join_streams(data):
    _need_dims(data, 2, "Input data must be a 2D array.")
    return data.reshape(-1).copy()


# ------------------------------------------------------------------ guard intervals on one row (prefix/models.py:34-101)
SHORT_ROW = "Input symbols length must be greater than prefix length."


def check_prefix_length(prefix_length: int) -> int:
    if prefix_length < 0:
        raise ValueError("Prefix length must be a non-negative integer.")
    return prefix_length


def _row(symbols):
    _need_dims(symbols, 1, "Input symbols must be a 1D array.")
    return symbols


def cyclic_extend(symbols, p: int):
    n = len(_row(symbols))
    if n < p:
        raise ValueError(SHORT_ROW)
    return symbols if p == 0 else np.concatenate((symbols[n - p:], symbols))     # p == 0: the same object


def cyclic_strip(symbols, p: int):
    if len(_row(symbols)) <= p:
        raise ValueError(SHORT_ROW)
    return symbols[p:]


def zero_extend(symbols, p: int):
    return np.concatenate((_row(symbols), np.zeros(p, dtype=symbols.dtype)))


def zero_fold(symbols, p: int):
    """Overlap-add of the trailing p samples onto the first p: the product with the reference's explicit
    [I_N | I_P; 0] matrix (prefix/models.py:87-101) written as a sum."""
    n = len(_row(symbols)) - p
    if n <= 0:
        raise ValueError(SHORT_ROW)
    out = np.array(symbols[:n], copy=True)
    out[:p] += symbols[n:]
    return out


# ------------------------------------------------------------------ channel and noise (channel/models.py:37-62, noise/models.py:13-22)
def unit_energy(impulse_response):
    energy = np.sum(np.abs(impulse_response) ** 2)
    if energy == 0:
        raise ValueError("Impulse response cannot be all zeros.")
    return impulse_response / np.sqrt(energy)


def causal_fir(signal, taps):
    _need_dims(signal, 1, "Signal must be serial (1D array)")
    return np.convolve(signal, taps, mode="full")[: signal.shape[0]].astype(np.complex128)


def awgn(signal, snr_db: float):
    """sigma^2 from the MEASURED stream power; legacy global NumPy RNG, real part drawn first."""
    sigma2 = np.mean(np.abs(signal) ** 2) / (10 ** (snr_db / 10))
    re = np.random.normal(size=signal.shape)
    im = np.random.normal(size=signal.shape)
    return signal + np.sqrt(sigma2 / 2) * (re + 1j * im)


# ------------------------------------------------------------------ one-tap equalisers (equalization/models.py:23-63)
def same_shape(received, response) -> None:
    if received.shape != response.shape:
        raise ValueError("Received symbols and channel frequency response must have the same shape.")


def zero_forcing(received, response):
    same_shape(received, response)
    return received / np.where(response == 0, 1e-10, response)


def mmse_noise_variance(received, response, snr_db) -> float:
    if snr_db is None:
        raise ValueError("SNR in dB must be provided to calculate noise variance.")
    gain = np.mean(np.abs(response) ** 2)
    noise = np.mean(np.abs(received) ** 2) / (10 ** (snr_db / 10))
    return float("inf") if gain == 0 else float(noise / gain)


def mmse(received, response, snr_db):
    sigma2 = mmse_noise_variance(received, response, snr_db)      # per row, from the row itself
    same_shape(received, response)
    return received * (np.conj(response) / (np.abs(response) ** 2 + sigma2))
