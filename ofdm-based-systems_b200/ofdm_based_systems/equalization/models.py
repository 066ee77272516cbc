"""Equaliser shells over ``_chain.zero_forcing`` / ``mmse``; the CUDA kernel applies the same one-tap rules from a
per-subcarrier table (csrc/link_fast.cuh, csrc/link_kernel.cuh)."""
import abc

from ofdm_based_systems import _chain


class IEqualizator(abc.ABC):
    def __init__(self, channel_frequency_response, snr_db=None):
        self.channel_frequency_response, self.snr_db = channel_frequency_response, snr_db

    @abc.abstractmethod
    def equalize(self, received_symbols):
        """one row of received subcarriers -> one row of equalised subcarriers"""


class ZeroForcingEqualizator(IEqualizator):
    def equalize(self, received_symbols):
        return _chain.zero_forcing(received_symbols, self.channel_frequency_response)


class MMSEEqualizator(IEqualizator):
    def calculate_noise_variance(self, received_signal) -> float:
        return _chain.mmse_noise_variance(received_signal, self.channel_frequency_response, self.snr_db)

    def equalize(self, received_symbols):
        return _chain.mmse(received_symbols, self.channel_frequency_response, self.snr_db)


class NoEqualizator(IEqualizator):
    def equalize(self, received_symbols):
        return received_symbols            # the input object itself
