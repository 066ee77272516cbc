"""QAM / PSK mappers (reference: constellation/models.py:11-474).

Host-side API objects: they build the tables, validate orders and implement the gap-rule bit
loading; ``Simulation.run()`` hands their parameters (order, scheme) to the CUDA link, which maps
and demaps in closed form.  The bit order (MSB first), the label -> point tables (Gray-encoded
position, odd rows mirrored) and the nearest-neighbour decision follow the reference exactly."""
from abc import ABC, abstractmethod
from functools import cached_property
from io import BytesIO
from typing import BinaryIO, Dict, List, Tuple, Type, Union

import numpy as np
from numpy.typing import NDArray
from scipy.stats import norm


class ISymbolClassifier(ABC):
    @abstractmethod
    def classify(self, constellation: NDArray[np.complex128], symbols: NDArray[np.complex128]
                 ) -> NDArray[np.complex128]:
        ...


class NNClassifier(ISymbolClassifier):
    """Nearest constellation point by Euclidean distance, first index on ties."""

    def classify(self, constellation, symbols):
        nearest = np.empty(len(symbols), dtype=np.int64)
        step = max(1, (1 << 22) // max(len(constellation), 1))      # bounded n x M scratch
        for lo in range(0, len(symbols), step):
            block = symbols[lo:lo + step]
            nearest[lo:lo + step] = np.argmin(np.abs(block[:, np.newaxis] - constellation[np.newaxis, :]), axis=1)
        return constellation[nearest]


class IWordCoder(ABC):
    def __init__(self, bits_per_word: int):
        self.bits_per_word = bits_per_word

    @abstractmethod
    def encode(self, word: int) -> int:
        ...

    @abstractmethod
    def decode(self, coded_word: int) -> int:
        ...

    @abstractmethod
    def reorder_constellation(self, constellation: NDArray[np.complex128], constellation_name: str
                              ) -> NDArray[np.complex128]:
        ...


class NoWordCoder(IWordCoder):
    @property
    def size(self) -> int:
        return 1 << self.bits_per_word

    def encode(self, word: int) -> int:
        if not 0 <= word < self.size:
            raise ValueError(f"Word must be in range [0, {self.size})")
        return word

    def decode(self, coded_word: int) -> int:
        if not 0 <= coded_word < self.size:
            raise ValueError(f"Coded word must be in range [0, {self.size})")
        return coded_word

    def reorder_constellation(self, constellation, constellation_name):
        return constellation


class GrayWordCoder(IWordCoder):
    @property
    def size(self) -> int:
        return 1 << self.bits_per_word

    @cached_property
    def gray_table(self) -> Dict[int, int]:
        return {w: w ^ (w >> 1) for w in range(self.size)}

    @cached_property
    def inverse_gray_table(self) -> Dict[int, int]:
        return {g: w for w, g in self.gray_table.items()}

    def encode(self, word: int) -> int:
        if not 0 <= word < self.size:
            raise ValueError(f"Word must be in range [0, {self.size})")
        return self.gray_table[word]

    def decode(self, coded_word: int) -> int:
        if not 0 <= coded_word < self.size:
            raise ValueError(f"Gray word must be in range [0, {self.size})")
        return self.inverse_gray_table[coded_word]

    def reorder_constellation(self, constellation, constellation_name):
        """Square QAM only: mirror every odd row (boustrophedon), which cancels the carry the Gray code
        leaks from the row bits into the column index."""
        if constellation_name != QAMConstellationMapper.__name__:
            return constellation
        side = int(np.sqrt(len(constellation)))
        grid = np.array(constellation).reshape(side, side)
        grid[1::2] = grid[1::2, ::-1]
        return grid.reshape(-1)


class IConstellationMapper(ABC):
    constellation: NDArray[np.complex128]
    constellation_map: Dict[Tuple[float, float], int]

    def __init__(self, order: int, word_coder: Type[IWordCoder] = GrayWordCoder,
                 classifier: Type[ISymbolClassifier] = NNClassifier):
        self.order = order
        self.word_coder = word_coder(bits_per_word=self.bits_per_symbol)
        self.classifier = classifier()

    @property
    @abstractmethod
    def constellation_name(self) -> str:
        ...

    @property
    @abstractmethod
    def bits_per_symbol(self) -> int:
        ...

    @abstractmethod
    def encode(self, bits: BinaryIO) -> NDArray[np.complex128]:
        ...

    @abstractmethod
    def decode(self, symbols: NDArray[np.complex128] | np.complex128) -> BinaryIO:
        ...

    @classmethod
    @abstractmethod
    def calculate_bit_loading_order(cls, ser: float, snr: float) -> int:
        ...


def _bits_of(stream_or_list: Union[BinaryIO, List[int]], group: int) -> np.ndarray:
    """MSB-first bits of a byte stream, zero padded to a multiple of ``group`` (lists pass through)."""
    if isinstance(stream_or_list, list):
        return np.asarray(stream_or_list, dtype=np.int64)
    raw = np.frombuffer(stream_or_list.read(), dtype=np.uint8)
    bits = np.unpackbits(raw, bitorder="big").astype(np.int64)
    if bits.size % group:
        bits = np.concatenate([bits, np.zeros(group - bits.size % group, dtype=np.int64)])
    return bits


def _labels_to_stream(labels: np.ndarray, bits_per_symbol: int) -> BinaryIO:
    shifts = np.arange(bits_per_symbol - 1, -1, -1)
    bits = ((labels[:, None] >> shifts[None, :]) & 1).astype(np.uint8).reshape(-1)
    return BytesIO(np.packbits(bits, bitorder="big").tobytes())       # last partial byte left-aligned


class _TableMapper(IConstellationMapper):
    """encode / decode shared by QAM and PSK: table lookup one way, NN search + dict lookup back."""

    def _table_key(self, point) -> Tuple[float, float]:
        return (float(point.real), float(point.imag))

    def _build_map(self, constellation) -> Dict[Tuple[float, float], int]:
        return {self._table_key(pt): i for i, pt in enumerate(constellation)}

    def encode(self, bits: Union[BinaryIO, List[int]]) -> NDArray[np.complex128]:
        k = self.bits_per_symbol
        chunks = _bits_of(bits, k).reshape(-1, k)
        labels = chunks.dot(1 << np.arange(k - 1, -1, -1))
        return self.constellation[labels]

    def decode(self, symbols) -> BinaryIO:
        if np.isscalar(symbols):
            symbols = np.array([symbols], dtype=np.complex128)
        else:
            symbols = np.asarray(symbols, dtype=np.complex128)
        decided = self.classifier.classify(self.constellation, symbols)
        labels = np.array([self.constellation_map[self._table_key(pt)] for pt in decided], dtype=np.int64)
        return _labels_to_stream(labels.reshape(-1), self.bits_per_symbol)


class QAMConstellationMapper(_TableMapper):
    """Square M-QAM, unit average energy.  Label b sits at grid position gray(b) of the row-major,
    top-to-bottom / left-to-right grid, with odd rows mirrored (closed form: column gray(b & (s-1)),
    row gray(b >> log2 s))."""

    def __init__(self, order: int, word_coder: Type[IWordCoder] = GrayWordCoder,
                 classifier: Type[ISymbolClassifier] = NNClassifier):
        super().__init__(order, word_coder, classifier)
        self.validate_order()
        self.constellation, self.constellation_map = self.generate_constellation()

    @property
    def constellation_name(self) -> str:
        return f"{self.order}-QAM"

    @property
    def bits_per_symbol(self) -> int:
        return int(np.log2(self.order))

    def validate_order(self) -> None:
        if int(np.sqrt(self.order)) ** 2 != self.order:
            raise ValueError("Order must be a perfect square (e.g., 4, 16, 64).")

    def generate_constellation(self):
        side = int(np.sqrt(self.order))
        axis = np.arange(-side + 1, side, 2)
        grid = (axis[np.newaxis, :] + 1j * axis[::-1, np.newaxis]).reshape(-1)      # rows: +Q first; cols: -I first
        coded = np.array([self.word_coder.encode(b) for b in range(self.order)])
        points = self.word_coder.reorder_constellation(grid[coded].astype(np.complex128),
                                                       QAMConstellationMapper.__name__)
        points = points / np.sqrt(np.mean(np.abs(points) ** 2))
        return points, self._build_map(points)

    @classmethod
    def calculate_bit_loading_order(cls, ser: float, snr: float) -> int:
        """Gap approximation: Gamma = Qinv(ser/4)^2 / 3, bits = round(log2(1 + snr / Gamma)) made even."""
        gap = norm.isf(ser / 4) ** 2 / 3
        bits = int(np.round(np.log2(1 + snr / gap)))
        bits -= bits % 2
        return 0 if bits <= 0 else 2 ** bits


class PSKConstellationMapper(_TableMapper):
    """M-PSK on the unit circle; label gray(k) sits at angle 2 pi k / M."""

    def __init__(self, order: int, word_coder: Type[IWordCoder] = GrayWordCoder,
                 classifier: Type[ISymbolClassifier] = NNClassifier):
        super().__init__(order, word_coder, classifier)
        self.validate_order()
        self.constellation, self.constellation_map = self.generate_constellation()

    @property
    def constellation_name(self) -> str:
        return f"{self.order}-PSK"

    @property
    def bits_per_symbol(self) -> int:
        return int(np.log2(self.order))

    def validate_order(self) -> None:
        k = np.log2(self.order)
        if k != int(k) or self.order < 2:
            raise ValueError("PSK order must be a power of 2 (e.g., 2, 4, 8, 16).")

    def generate_constellation(self):
        circle = np.exp(1j * (2 * np.pi * np.arange(self.order) / self.order))
        points = np.zeros(self.order, dtype=np.complex128)
        for k in range(self.order):
            points[self.word_coder.encode(k)] = circle[k]
        points = self.word_coder.reorder_constellation(points, PSKConstellationMapper.__name__)
        return points, self._build_map(points)

    @classmethod
    def calculate_bit_loading_order(cls, ser: float, snr: float) -> int:
        q = norm.isf(ser / 2)
        g_star = q ** 2 / (2 * np.pi ** 2)
        gap = np.sqrt(snr * g_star) / (1 - np.sqrt(g_star / (snr + 1e-10)))
        bits = int(np.floor(np.log2(1 + snr / (gap + 1e-10)) + 1e-10))
        return 0 if bits <= 0 else 2 ** bits
