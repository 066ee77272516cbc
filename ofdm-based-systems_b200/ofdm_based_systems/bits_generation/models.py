"""Bit sources (reference: bits_generation/models.py:12-163).  ``Simulation.run()`` does not call these
on the GPU path (bits come from Philox in registers); they serve component-level callers and the
replay mode, which consumes exactly these byte streams."""
import math
from abc import ABC, abstractmethod
from io import BytesIO
from typing import BinaryIO, Tuple

import numpy as np
from numpy.random import PCG64, Generator
from numpy.typing import NDArray


class IGenerator(ABC):
    @abstractmethod
    def generate_bits(self, num_bits: int) -> BinaryIO:
        ...


def _random_stream(generator: Generator, num_bits: int) -> BinaryIO:
    """ceil(num_bits / 8) generator bytes, MSB first, the unused low bits of the last byte cleared."""
    data = generator.bytes(math.ceil(num_bits / 8))
    tail = num_bits % 8
    if tail:
        keep = (0xFF << (8 - tail)) & 0xFF
        data = data[:-1] + bytes([data[-1] & keep])
    stream = BytesIO(data)
    stream.seek(0)
    return stream


class RandomBitsGenerator(IGenerator):
    def __init__(self, generator: Generator = Generator(PCG64())):
        self.generator = generator

    def generate_bits(self, num_bits: int) -> BinaryIO:
        return _random_stream(self.generator, num_bits)


class AdaptiveBitsGenerator(IGenerator):
    """Exactly sum(bits_per_subcarrier) * num_ofdm_symbols bits (per-subcarrier loading)."""

    def __init__(self, bits_per_subcarrier: NDArray[np.int64], num_ofdm_symbols: int,
                 generator: Generator = Generator(PCG64())):
        if len(bits_per_subcarrier) == 0:
            raise ValueError("bits_per_subcarrier cannot be empty")
        if num_ofdm_symbols <= 0:
            raise ValueError(f"num_ofdm_symbols must be positive, got {num_ofdm_symbols}")
        self.bits_per_subcarrier = np.array(bits_per_subcarrier, dtype=np.int64)
        self.num_ofdm_symbols = num_ofdm_symbols
        self.generator = generator

    def generate_bits(self, num_bits: int = 0) -> BinaryIO:
        return _random_stream(self.generator, self.get_total_bits())

    def get_total_bits(self) -> int:
        return int(np.sum(self.bits_per_subcarrier) * self.num_ofdm_symbols)

    @staticmethod
    def calculate_requirements(constellation_orders: NDArray[np.int64], num_ofdm_symbols: int
                               ) -> Tuple[int, NDArray[np.int64]]:
        bits = np.array([int(np.log2(o)) if o > 0 else 0 for o in constellation_orders], dtype=np.int64)
        return int(np.sum(bits) * num_ofdm_symbols), bits
