"""Guard-interval shells over the row operations of ``_chain``; (name, extend, strip) per scheme."""
import abc

from ofdm_based_systems import _chain


class IPrefixScheme(abc.ABC):
    def __init__(self, prefix_length: int = 0):
        self.prefix_length = _chain.check_prefix_length(prefix_length)

    @property
    @abc.abstractmethod
    def acronym(self) -> str:
        """short tag used in titles and file names"""

    @abc.abstractmethod
    def add_prefix(self, symbols):
        """row of N samples -> row of N + prefix_length samples"""

    @abc.abstractmethod
    def remove_prefix(self, symbols):
        """row of N + prefix_length samples -> row of N samples"""


def _scheme(name: str, tag: str, extend, strip):
    body = {"acronym": property(lambda self: tag),
            "add_prefix": lambda self, symbols: extend(symbols, self.prefix_length),
            "remove_prefix": lambda self, symbols: strip(symbols, self.prefix_length),
            "__module__": __name__}
    return type(name, (IPrefixScheme,), body)


CyclicPrefixScheme = _scheme("CyclicPrefixScheme", "CP", _chain.cyclic_extend, _chain.cyclic_strip)
ZeroPaddingPrefixScheme = _scheme("ZeroPaddingPrefixScheme", "ZP", _chain.zero_extend, _chain.zero_fold)
NoPrefixScheme = _scheme("NoPrefixScheme", "", lambda symbols, p: symbols, lambda symbols, p: symbols)
