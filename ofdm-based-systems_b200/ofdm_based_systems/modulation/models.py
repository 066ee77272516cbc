"""OFDM / SC-OFDM modulators on (num_ofdm_symbols, N[+P]) arrays (reference: modulation/models.py:9-91)."""
from abc import ABC, abstractmethod

import numpy as np
from numpy.typing import NDArray

from ofdm_based_systems.equalization.models import IEqualizator
from ofdm_based_systems.prefix.models import IPrefixScheme


class IModulator(ABC):
    @abstractmethod
    def modulate(self, symbols: NDArray[np.complex128]) -> NDArray[np.complex128]:
        ...

    @abstractmethod
    def demodulate(self, symbols: NDArray[np.complex128]) -> NDArray[np.complex128]:
        ...


class _PrefixedModulator(IModulator):
    num_subcarriers: int
    prefix_scheme: IPrefixScheme
    equalizator: IEqualizator

    def _with_prefix(self, rows):
        return np.array([self.prefix_scheme.add_prefix(r) for r in rows])

    def _to_equalised_frequency(self, rows):
        stripped = np.array([self.prefix_scheme.remove_prefix(r) for r in rows])
        freq = np.fft.fft(stripped, n=self.num_subcarriers, axis=1, norm="ortho")
        return np.array([self.equalizator.equalize(r) for r in freq])


class OFDMModulator(_PrefixedModulator):
    def __init__(self, num_subcarriers: int, prefix_scheme: IPrefixScheme, equalizator: IEqualizator):
        self.num_subcarriers = num_subcarriers
        self.prefix_scheme = prefix_scheme
        self.equalizator = equalizator

    def modulate(self, symbols):
        if symbols.shape[1] != self.num_subcarriers:
            raise ValueError(f"Number of symbols must be {self.num_subcarriers}")
        return self._with_prefix(np.fft.ifft(symbols, n=self.num_subcarriers, axis=1, norm="ortho"))

    def demodulate(self, symbols):
        return self._to_equalised_frequency(symbols)


class SingleCarrierOFDMModulator(_PrefixedModulator):
    def __init__(self, prefix_scheme: IPrefixScheme, equalizator: IEqualizator, num_subcarriers: int):
        self.prefix_scheme = prefix_scheme
        self.equalizator = equalizator
        self.num_subcarriers = num_subcarriers

    def modulate(self, symbols):
        return self._with_prefix(symbols)

    def demodulate(self, symbols):
        return np.fft.ifft(self._to_equalised_frequency(symbols), n=self.num_subcarriers, axis=1, norm="ortho")
