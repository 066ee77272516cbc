"""Serial <-> parallel reshaping (reference: serial_parallel/models.py:5-21).  Inside the CUDA kernel
this stage is pure indexing; the class exists for callers that drive the chain component by component."""
import numpy as np
from numpy.typing import NDArray


class SerialToParallelConverter:
    @staticmethod
    def to_parallel(data: NDArray[np.complex128], num_streams: int) -> NDArray[np.complex128]:
        if data.ndim != 1:
            raise ValueError("Input data must be a 1D array.")
        if num_streams <= 0:
            raise ValueError("Number of streams must be a positive integer.")
        rows, rest = divmod(len(data), num_streams)
        if rest:
            raise ValueError("Length of data must be divisible by number of streams.")
        return data.reshape(rows, num_streams)

    @staticmethod
    def to_serial(data: NDArray[np.complex128]) -> NDArray[np.complex128]:
        if data.ndim != 2:
            raise ValueError("Input data must be a 2D array.")
        return data.reshape(-1).copy()
