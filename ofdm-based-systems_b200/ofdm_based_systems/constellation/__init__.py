from ofdm_based_systems.constellation.adaptive import AdaptiveConstellationMapper, calculate_constellation_orders
from ofdm_based_systems.constellation.models import (
    GrayWordCoder, IConstellationMapper, ISymbolClassifier, IWordCoder, NNClassifier, NoWordCoder,
    PSKConstellationMapper, QAMConstellationMapper)

__all__ = ["AdaptiveConstellationMapper", "calculate_constellation_orders", "GrayWordCoder", "IConstellationMapper",
           "ISymbolClassifier", "IWordCoder", "NNClassifier", "NoWordCoder", "PSKConstellationMapper",
           "QAMConstellationMapper"]
