"""Channel shell: unit-energy taps, causal FIR over the serial stream, then the noise model."""
import numpy as np

from ofdm_based_systems import _chain
from ofdm_based_systems.noise.models import AWGNoiseModel, INoiseModel


class ChannelModel:
    def __init__(self, impulse_response, snr_db: float, noise_model: INoiseModel = AWGNoiseModel()):
        self.impulse_response = self.normalize_impulse_response(impulse_response)
        self.snr_db, self.noise_model = snr_db, noise_model
        self.frequency_response_cache = {}

    normalize_impulse_response = staticmethod(_chain.unit_energy)

    order = property(lambda self: len(self.impulse_response) - 1)

    def get_frequency_response(self, n_fft: int):
        cache = self.frequency_response_cache
        if n_fft not in cache:
            cache[n_fft] = np.fft.fft(self.impulse_response, n=n_fft)
        return cache[n_fft]

    def get_gains(self, n_fft: int):
        return np.abs(self.get_frequency_response(n_fft)) ** 2

    def transmit(self, signal):
        return self.noise_model.add_noise(_chain.causal_fir(signal, self.impulse_response), self.snr_db)
