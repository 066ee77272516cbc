"""Bit sources of the chain (host side; the fused CUDA mode draws its bits from Philox in registers)."""
from ofdm_based_systems.bits_generation import models as _m

__all__ = ["IGenerator", "RandomBitsGenerator", "AdaptiveBitsGenerator"]
globals().update({name: getattr(_m, name) for name in __all__})
