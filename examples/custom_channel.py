"""A custom impulse response, ZF vs MMSE and the three guard-interval schemes over an SNR grid: one kernel launch per
point (`LinkSweep`), 2e8 bits per point."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ofdm-based-systems_b200"))

from ofdm_based_systems.simulation.sweep import LinkConfig, LinkSweep

taps = np.load(os.path.join(ROOT, "config", "channel_models", "severe_multipath.npy"))
snrs = [10.0, 15.0, 20.0, 25.0, 30.0]
n, order = 256, 64
print(f"severe_multipath.npy: {len(taps)} taps; N={n}, {order}-QAM, {200_000_000 // (n * 6)} OFDM symbols per point")
print("scheme              " + "".join(f"{s:>11.0f} dB" for s in snrs))
for prefix, P in (("CYCLIC", 7), ("ZERO", 7), ("CYCLIC", 3), ("NONE", 0)):
    for eq in ("ZF", "MMSE"):
        cfg = LinkConfig(num_subcarriers=n, taps_raw=taps, constellation_order=order, prefix_scheme=prefix, prefix_length=P,
                         equalizator_type=eq)
        sweep = LinkSweep(cfg)
        res = sweep.sweep(snrs, 200_000_000 // (n * 6), seed=7)
        kind = "fast" if sweep.link.uses_fast_kernel else "general"
        sweep.close()
        print(f"{prefix:6s} P={P} {eq:4s} {kind:7s}" + "".join(f"{r['bit_error_rate']:14.3e}" for r in res))
