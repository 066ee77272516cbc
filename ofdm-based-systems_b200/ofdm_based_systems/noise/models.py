"""Noise models (reference: noise/models.py:6-27).  In the fused CUDA mode the same sigma^2 rule is
applied with Philox/Box-Muller samples generated in registers."""
from abc import ABC, abstractmethod

import numpy as np
from numpy.typing import NDArray


class INoiseModel(ABC):
    @abstractmethod
    def add_noise(self, signal: NDArray[np.complex128], snr_db: float) -> NDArray[np.complex128]:
        ...


class AWGNoiseModel(INoiseModel):
    def add_noise(self, signal: NDArray[np.complex128], snr_db: float) -> NDArray[np.complex128]:
        # sigma^2 from the MEASURED stream power; legacy global RNG, real part drawn first
        sigma2 = np.mean(np.abs(signal) ** 2) / (10 ** (snr_db / 10))
        re = np.random.normal(size=signal.shape)
        im = np.random.normal(size=signal.shape)
        return signal + np.sqrt(sigma2 / 2) * (re + 1j * im)


class NoNoiseModel(INoiseModel):
    def add_noise(self, signal: NDArray[np.complex128], snr_db: float) -> NDArray[np.complex128]:
        return signal
