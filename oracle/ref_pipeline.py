"""The UNMODIFIED reference's own component pipeline for the headline workload, timed on host cores.
TEST / BENCH INFRASTRUCTURE ONLY (bench.py `cpu_baseline` and `--impl reference`).

It imports the reference package from oracle/_ref/src (a byte-for-byte copy made by oracle/build_ref.py in the
build container; /root/reference itself does not exist on the GPU box) and drives exactly the calls
``Simulation.run()`` makes between simulation/models.py:454 and :606 - the pipeline of the reference's own
integration test (tests/integration/test_end_to_end.py:205-257) - without the plotting: RandomBitsGenerator ->
QAMConstellationMapper.encode -> SerialToParallelConverter -> OFDMModulator.modulate (CyclicPrefixScheme) ->
ChannelModel.transmit (AWGNoiseModel) -> OFDMModulator.demodulate (MMSEEqualizator) -> mapper.decode ->
read_bits_from_stream -> error count.  matplotlib is absent from this image; simulation/models.py imports it at
module level, so the no-op stub oracle/_mplstub is put on sys.path for that import only."""
from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.path.join(HERE, "_ref", "src")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "ofdm_based_systems"))


_REF = None


def _import_reference():
    global _REF
    if _REF is not None:
        return _REF
    if not available():
        raise RuntimeError("oracle/_ref is missing: run oracle/build_ref.py where /root/reference is mounted")
    # the product package has the same name: the reference copy must come first on the path of THIS process
    for mod in [m for m in sys.modules if m == "ofdm_based_systems" or m.startswith("ofdm_based_systems.")]:
        f = getattr(sys.modules[mod], "__file__", None)
        if f and not f.startswith(REF_SRC):
            raise RuntimeError("the product package is already imported in this process; run the reference in its own process")
    sys.path.insert(0, REF_SRC)
    sys.path.insert(1, os.path.join(HERE, "_mplstub"))
    from ofdm_based_systems.bits_generation.models import RandomBitsGenerator
    from ofdm_based_systems.channel.models import ChannelModel
    from ofdm_based_systems.constellation.models import QAMConstellationMapper
    from ofdm_based_systems.equalization.models import MMSEEqualizator
    from ofdm_based_systems.modulation.models import OFDMModulator
    from ofdm_based_systems.noise.models import AWGNoiseModel
    from ofdm_based_systems.prefix.models import CyclicPrefixScheme
    from ofdm_based_systems.serial_parallel.models import SerialToParallelConverter
    from ofdm_based_systems.simulation.models import read_bits_from_stream
    _REF = dict(locals())
    return _REF


def run_chunk(args):
    """One bounded sample of the headline workload on this core: (seed, n_ofdm, n_sc, order, snr_db, taps) ->
    (bits, bit_errors, seconds).  Everything inside is the reference's code."""
    seed, n_ofdm, n_sc, order, snr_db, taps = args
    import contextlib
    import io
    ref = _import_reference()
    np.random.seed(seed)
    bps = int(np.log2(order))
    total_bits = n_ofdm * n_sc * bps
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):        # the reference prints per call
        stream = ref["RandomBitsGenerator"]().generate_bits(total_bits)
        bits = ref["read_bits_from_stream"](stream)
        mapper = ref["QAMConstellationMapper"](order=order)
        symbols = mapper.encode(stream)
        sp = ref["SerialToParallelConverter"]()
        channel = ref["ChannelModel"](impulse_response=np.asarray(taps, dtype=np.complex128), snr_db=snr_db,
                                      noise_model=ref["AWGNoiseModel"]())
        prefix = ref["CyclicPrefixScheme"](prefix_length=int(1.0 * channel.order))
        equalizer = ref["MMSEEqualizator"](channel_frequency_response=np.fft.fft(taps, n_sc), snr_db=snr_db)
        modulator = ref["OFDMModulator"](num_subcarriers=n_sc, prefix_scheme=prefix, equalizator=equalizer)
        tx = sp.to_serial(modulator.modulate(sp.to_parallel(symbols, n_sc)))
        rx = channel.transmit(tx)
        z = sp.to_serial(modulator.demodulate(sp.to_parallel(rx, n_sc + prefix.prefix_length)))
        decoded = ref["read_bits_from_stream"](mapper.decode(z))
        errors = sum(b1 != b2 for b1, b2 in zip(bits, decoded))
    return total_bits, int(errors), time.perf_counter() - t0


if __name__ == "__main__":
    taps = np.load(os.path.join(os.path.dirname(HERE), "config", "channel_models", "severe_multipath.npy"))
    print(run_chunk((1, 20, 1024, 64, 20.0, taps)))
