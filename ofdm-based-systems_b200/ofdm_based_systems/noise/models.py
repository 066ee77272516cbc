"""Noise-model shells.  In the fused CUDA mode the same sigma^2 rule (``_chain.awgn``) is applied to Philox /
Box-Muller samples generated in registers; replay mode records what these classes add."""
import abc

from ofdm_based_systems import _chain


class INoiseModel(abc.ABC):
    @abc.abstractmethod
    def add_noise(self, signal, snr_db):
        """signal + noise for the given SNR (dB) of the measured stream power"""


class AWGNoiseModel(INoiseModel):
    def add_noise(self, signal, snr_db):
        return _chain.awgn(signal, snr_db)


class NoNoiseModel(INoiseModel):
    def add_noise(self, signal, snr_db):
        return signal            # the input object itself
