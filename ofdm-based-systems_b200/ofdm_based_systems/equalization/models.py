"""Per-row frequency-domain equalisers (reference: equalization/models.py:8-68)."""
from abc import ABC, abstractmethod
from typing import Optional

import numpy as np
from numpy.typing import NDArray


class IEqualizator(ABC):
    def __init__(self, channel_frequency_response: NDArray[np.complex128], snr_db: Optional[float] = None):
        self.channel_frequency_response = channel_frequency_response
        self.snr_db = snr_db

    @abstractmethod
    def equalize(self, received_symbols: NDArray[np.complex128]) -> NDArray[np.complex128]:
        ...

    def _check_shape(self, received_symbols) -> None:
        if received_symbols.shape != self.channel_frequency_response.shape:
            raise ValueError("Received symbols and channel frequency response must have the same shape.")


class ZeroForcingEqualizator(IEqualizator):
    def equalize(self, received_symbols):
        self._check_shape(received_symbols)
        h = self.channel_frequency_response
        return received_symbols / np.where(h == 0, 1e-10, h)


class MMSEEqualizator(IEqualizator):
    def calculate_noise_variance(self, received_signal) -> float:
        if self.snr_db is None:
            raise ValueError("SNR in dB must be provided to calculate noise variance.")
        gain = np.mean(np.abs(self.channel_frequency_response) ** 2)
        noise = np.mean(np.abs(received_signal) ** 2) / (10 ** (self.snr_db / 10))
        return float("inf") if gain == 0 else float(noise / gain)

    def equalize(self, received_symbols):
        sigma2 = self.calculate_noise_variance(received_symbols)     # per row, from the row itself
        self._check_shape(received_symbols)
        h = self.channel_frequency_response
        return received_symbols * (np.conj(h) / (np.abs(h) ** 2 + sigma2))


class NoEqualizator(IEqualizator):
    def equalize(self, received_symbols):
        return received_symbols
