"""Prefix schemes on 1-D rows (reference: prefix/models.py:7-113)."""
from abc import ABC, abstractmethod

import numpy as np
from numpy.typing import NDArray


class IPrefixScheme(ABC):
    prefix_length: int

    def __init__(self, prefix_length: int = 0):
        if prefix_length < 0:
            raise ValueError("Prefix length must be a non-negative integer.")
        self.prefix_length = prefix_length

    @property
    @abstractmethod
    def acronym(self) -> str:
        ...

    @abstractmethod
    def add_prefix(self, symbols: NDArray[np.complex128]) -> NDArray[np.complex128]:
        ...

    @abstractmethod
    def remove_prefix(self, symbols: NDArray[np.complex128]) -> NDArray[np.complex128]:
        ...


def _require_row(symbols) -> None:
    if symbols.ndim != 1:
        raise ValueError("Input symbols must be a 1D array.")


class CyclicPrefixScheme(IPrefixScheme):
    @property
    def acronym(self) -> str:
        return "CP"

    def add_prefix(self, symbols):
        _require_row(symbols)
        p = self.prefix_length
        if len(symbols) < p:
            raise ValueError("Input symbols length must be greater than prefix length.")
        if p == 0:
            return symbols                       # same object, as in the reference
        return np.concatenate((symbols[len(symbols) - p:], symbols))

    def remove_prefix(self, symbols):
        _require_row(symbols)
        if len(symbols) <= self.prefix_length:
            raise ValueError("Input symbols length must be greater than prefix length.")
        return symbols[self.prefix_length:]


class ZeroPaddingPrefixScheme(IPrefixScheme):
    @property
    def acronym(self) -> str:
        return "ZP"

    def add_prefix(self, symbols):
        _require_row(symbols)
        return np.concatenate((symbols, np.zeros(self.prefix_length, dtype=symbols.dtype)))

    def remove_prefix(self, symbols):
        """Overlap-add: the trailing P samples are folded onto the first P (the reference builds the
        explicit [I_N | I_P; 0] matrix, prefix/models.py:87-101; the product is this sum)."""
        _require_row(symbols)
        p = self.prefix_length
        if len(symbols) <= p:
            raise ValueError("Input symbols length must be greater than prefix length.")
        n = len(symbols) - p
        out = np.array(symbols[:n], copy=True)
        out[:p] += symbols[n:]
        return out


class NoPrefixScheme(IPrefixScheme):
    @property
    def acronym(self) -> str:
        return ""

    def add_prefix(self, symbols):
        return symbols

    def remove_prefix(self, symbols):
        return symbols
