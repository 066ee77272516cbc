"""API shell over ``_chain.split_streams`` / ``join_streams``.  Inside the CUDA kernel this stage is pure indexing."""
from ofdm_based_systems import _chain


class SerialToParallelConverter:
    to_parallel = staticmethod(_chain.split_streams)
    to_serial = staticmethod(_chain.join_streams)
