"""FIR multipath channel + noise on the serial stream (reference: channel/models.py:7-62)."""
import numpy as np
from numpy.typing import NDArray

from ofdm_based_systems.noise.models import AWGNoiseModel, INoiseModel


class ChannelModel:
    def __init__(self, impulse_response: NDArray[np.complex128], snr_db: float,
                 noise_model: INoiseModel = AWGNoiseModel()):
        self.impulse_response = self.normalize_impulse_response(impulse_response)
        self.snr_db = snr_db
        self.noise_model = noise_model
        self.frequency_response_cache: dict[int, NDArray[np.complex128]] = {}

    @property
    def order(self) -> int:
        return len(self.impulse_response) - 1

    def get_frequency_response(self, n_fft: int) -> NDArray[np.complex128]:
        cache = self.frequency_response_cache
        if n_fft not in cache:
            cache[n_fft] = np.fft.fft(self.impulse_response, n=n_fft)
        return cache[n_fft]

    def get_gains(self, n_fft: int) -> NDArray[np.float64]:
        return np.abs(self.get_frequency_response(n_fft)) ** 2

    def normalize_impulse_response(self, impulse_response):
        energy = np.sum(np.abs(impulse_response) ** 2)
        if energy == 0:
            raise ValueError("Impulse response cannot be all zeros.")
        return impulse_response / np.sqrt(energy)

    def transmit(self, signal: NDArray[np.complex128]) -> NDArray[np.complex128]:
        if signal.ndim != 1:
            raise ValueError("Signal must be serial (1D array)")
        faded = np.convolve(signal, self.impulse_response, mode="full")[: signal.shape[0]]
        return self.noise_model.add_noise(faded.astype(np.complex128), self.snr_db)
