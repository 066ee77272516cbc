from ofdm_based_systems.bits_generation.models import AdaptiveBitsGenerator, IGenerator, RandomBitsGenerator

__all__ = ["IGenerator", "RandomBitsGenerator", "AdaptiveBitsGenerator"]
