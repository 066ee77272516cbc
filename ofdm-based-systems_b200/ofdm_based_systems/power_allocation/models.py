"""Power allocation across subcarriers (reference: power_allocation/models.py:13-334).

Host-side fp64 API objects.  The batched CUDA counterpart (one channel realisation per warp) is
``ofdm_waterfill_bitload_batched`` in libofdm_b200.so; this class is what configures a single link."""
from abc import ABC, abstractmethod

import numpy as np
from numpy.typing import NDArray


class IPowerAllocation(ABC):
    @abstractmethod
    def allocate(self) -> NDArray[np.float64]:
        ...


class UniformPowerAllocation(IPowerAllocation):
    def __init__(self, total_power: float, num_subcarriers: int):
        if total_power < 0:
            raise ValueError(f"Total power must be non-negative, got {total_power}")
        if num_subcarriers <= 0:
            raise ValueError(f"Number of subcarriers must be positive, got {num_subcarriers}")
        self.total_power = total_power
        self.num_subcarriers = num_subcarriers

    def allocate(self) -> NDArray[np.float64]:
        return np.full(self.num_subcarriers, self.total_power / self.num_subcarriers, dtype=np.float64)


class WaterfillingPowerAllocation(IPowerAllocation):
    """P_k = max(0, mu - floor_k) with floor_k = N0 / (g_k * N)  (the extra 1/N is the reference's,
    power_allocation/models.py:161), mu found by bisection, result rescaled to the exact budget."""

    def __init__(self, total_power: float, channel_gains: NDArray[np.float64], noise_power: float,
                 tolerance: float = 1e-8):
        if total_power < 0:
            raise ValueError(f"Total power must be non-negative, got {total_power}")
        if noise_power < 0:
            raise ValueError(f"Noise power must be non-negative, got {noise_power}")
        if len(channel_gains) == 0:
            raise ValueError("Channel gains array cannot be empty")
        if np.any(channel_gains <= 0):
            raise ValueError("All channel gains must be positive, "
                             f"got min={np.min(channel_gains)}, max={np.max(channel_gains)}")
        self.total_power = total_power
        self.channel_gains = np.array(channel_gains, dtype=np.float64)
        self.noise_power = noise_power
        self.tolerance = tolerance
        self.num_subcarriers = len(channel_gains)

    def allocate(self) -> NDArray[np.float64]:
        floor = self.noise_power / (self.channel_gains * len(self.channel_gains))
        level = self._find_water_level(floor)
        power = np.maximum(0, level - floor)
        total = np.sum(power)
        if total > 0:
            power = power * (self.total_power / total)
        return power

    def _find_water_level(self, floor: NDArray[np.float64]) -> float:
        lo, hi = 0.0, self.total_power + np.max(floor)
        level = (lo + hi) / 2
        for _ in range(100):
            level = (lo + hi) / 2
            poured = np.sum(np.maximum(0, level - floor))
            if np.abs(poured - self.total_power) < self.tolerance:
                return level
            if poured < self.total_power:
                lo = level
            else:
                hi = level
        return level


def calculate_capacity_per_subcarrier(power_allocation, channel_gains, noise_power) -> NDArray[np.float64]:
    return np.log2(1 + power_allocation * channel_gains / noise_power + 1e-12)


def calculate_capacity(power_allocation, channel_gains, noise_power) -> float:
    return np.sum(np.log2(1 + power_allocation * channel_gains / noise_power + 1e-12))


def compare_allocations(uniform, waterfilling, channel_gains, noise_power) -> dict:
    cap_u = calculate_capacity(uniform, channel_gains, noise_power)
    cap_w = calculate_capacity(waterfilling, channel_gains, noise_power)
    return {"uniform_capacity": cap_u, "waterfilling_capacity": cap_w, "capacity_gain": cap_w - cap_u,
            "capacity_gain_percent": 100 * (cap_w - cap_u) / cap_u if cap_u > 0 else 0}
