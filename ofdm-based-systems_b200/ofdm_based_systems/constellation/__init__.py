"""Constellation mappers: public names of ``constellation.models`` and ``constellation.adaptive``."""
from ofdm_based_systems.constellation import adaptive as _a, models as _m

__all__ = ["AdaptiveConstellationMapper", "calculate_constellation_orders", "GrayWordCoder", "IConstellationMapper",
           "ISymbolClassifier", "IWordCoder", "NNClassifier", "NoWordCoder", "PSKConstellationMapper",
           "QAMConstellationMapper"]
globals().update({name: getattr(_a, name, None) or getattr(_m, name) for name in __all__})
