"""Per-subcarrier constellation orders (reference: constellation/adaptive.py:16-329).  Bits are
consumed OFDM-symbol-major, subcarrier-minor, log2(order_k) bits each; order 0 carries 0+0j."""
from io import BytesIO
from typing import BinaryIO, Dict, List, Tuple, Type, Union

import numpy as np
from numpy.typing import NDArray

from ofdm_based_systems.constellation.models import IConstellationMapper


class AdaptiveConstellationMapper(IConstellationMapper):
    def __init__(self, constellation_orders: NDArray[np.int64], base_mapper_class: Type[IConstellationMapper],
                 num_subcarriers: int):
        if len(constellation_orders) != num_subcarriers:
            raise ValueError(f"constellation_orders length ({len(constellation_orders)}) "
                             f"must match num_subcarriers ({num_subcarriers})")
        self.constellation_orders = np.array(constellation_orders, dtype=np.int64)
        self.base_mapper_class = base_mapper_class
        self.num_subcarriers = num_subcarriers
        self.mappers: Dict[int, Tuple[List[int], IConstellationMapper]] = {}
        active_orders = [int(o) for o in np.unique(constellation_orders) if o > 0]
        for order in active_orders:
            members = np.flatnonzero(np.asarray(constellation_orders) == order).tolist()
            self.mappers[order] = (members, base_mapper_class(order=order))
        self.bits_per_subcarrier = np.array([int(np.log2(o)) if o > 0 else 0 for o in constellation_orders],
                                            dtype=np.int64)
        pool: List[complex] = []
        for order in active_orders:
            pool.extend(self.mappers[order][1].constellation.tolist())
        self.constellation = np.unique(np.array(pool, dtype=np.complex128))
        self.constellation_map = {(float(p.real), float(p.imag)): i for i, p in enumerate(self.constellation)}

    @property
    def order(self) -> int:
        return int(np.max(self.constellation_orders))

    @property
    def constellation_name(self) -> str:
        used = np.unique(self.constellation_orders[self.constellation_orders > 0])
        family = self.base_mapper_class.__name__.replace("ConstellationMapper", "")
        if len(used) == 0:
            return "No-Transmission"
        if len(used) == 1:
            return f"{int(used[0])}-{family}"
        return f"Adaptive-{int(used.min())}-to-{int(used.max())}-{family}"

    @property
    def bits_per_symbol(self) -> int:
        return int(np.max(self.bits_per_subcarrier))

    def get_bits_per_subcarrier(self) -> NDArray[np.int64]:
        return self.bits_per_subcarrier

    def get_constellation_orders(self) -> NDArray[np.int64]:
        return self.constellation_orders

    def encode(self, bits: Union[BinaryIO, List[int]]) -> NDArray[np.complex128]:
        if isinstance(bits, list):
            flat = np.asarray(bits, dtype=np.int64)
        else:
            flat = np.unpackbits(np.frombuffer(bits.read(), dtype=np.uint8), bitorder="big").astype(np.int64)
        per_ofdm = int(np.sum(self.bits_per_subcarrier))
        if per_ofdm == 0:
            raise ValueError("No active subcarriers (all orders are zero)")
        if flat.size % per_ofdm != 0:
            raise ValueError(f"Bits length ({flat.size}) must be multiple of bits_per_symbol ({per_ofdm})")
        frames = flat.reshape(-1, per_ofdm)
        out = np.zeros((frames.shape[0], self.num_subcarriers), dtype=np.complex128)
        start = np.concatenate([[0], np.cumsum(self.bits_per_subcarrier)])
        for k, width in enumerate(self.bits_per_subcarrier):
            if width == 0:
                continue
            labels = frames[:, start[k]:start[k + 1]].dot(1 << np.arange(width - 1, -1, -1))
            out[:, k] = self.mappers[int(self.constellation_orders[k])][1].constellation[labels]
        return out.reshape(-1)

    def decode(self, symbols: Union[NDArray[np.complex128], np.complex128]) -> BinaryIO:
        if np.isscalar(symbols):
            symbols = np.array([symbols], dtype=np.complex128)
        else:
            symbols = np.asarray(symbols, dtype=np.complex128)
        if len(symbols) % self.num_subcarriers != 0:
            raise ValueError(f"Symbols length ({len(symbols)}) must be multiple of "
                             f"num_subcarriers ({self.num_subcarriers})")
        grid = symbols.reshape(-1, self.num_subcarriers)
        columns = []
        for k, width in enumerate(self.bits_per_subcarrier):
            if width == 0:
                continue
            mapper = self.mappers[int(self.constellation_orders[k])][1]
            raw = np.frombuffer(mapper.decode(grid[:, k]).read(), dtype=np.uint8)   # width bits per symbol, packed
            unpacked = np.unpackbits(raw, bitorder="big")[: grid.shape[0] * width]
            columns.append(unpacked.reshape(grid.shape[0], width))
        bits = np.concatenate(columns, axis=1).reshape(-1) if columns else np.zeros(0, dtype=np.uint8)
        whole = (bits.size // 8) * 8                                  # a trailing partial byte is dropped
        return BytesIO(np.packbits(bits[:whole], bitorder="big").tobytes())

    def calculate_bit_loading_order(self, ser: float, snr: float) -> int:
        raise NotImplementedError("This method is not implemented in AdaptiveConstellationMapper.")


def calculate_constellation_orders(capacity: NDArray[np.float64], min_order: int, max_order: int,
                                   scaling_factor: float, base_mapper_class: Type[IConstellationMapper]
                                   ) -> NDArray[np.int64]:
    """Shannon-capacity rule: scale, clip to log2(max_order), QAM -> even bits, below log2(min_order) -> off."""
    from ofdm_based_systems.constellation.models import QAMConstellationMapper
    bits = np.clip(capacity * scaling_factor, 0, np.log2(max_order))
    bits = bits // 2 * 2 if base_mapper_class == QAMConstellationMapper else np.floor(bits)
    bits = np.where(bits < np.log2(min_order), 0, bits)
    return np.where(bits > 0, 2 ** bits, 0).astype(np.int64)
