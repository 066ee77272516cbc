"""BASELINE config #4: a fresh 8-tap Rayleigh realisation per frame, water-filling + gap-rule bit loading bounded to
QPSK .. 256-QAM, everything on the GPU (tap draw, water-filling, table build, link kernel), one call per SNR point."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ofdm-based-systems_b200"))

from ofdm_based_systems._native import run_frames
from ofdm_based_systems.simulation.sweep import FrameSweep

snrs = [5.0, 10.0, 15.0, 20.0, 25.0, 30.0]
sweep = FrameSweep(64, n_taps=8, equalizer="MMSE", waterfilling=True, min_order=4, max_order=256, ser=1e-3)
sweep.sweep(snrs[:1], 16, 10)
t0 = time.perf_counter()
res = sweep.sweep(snrs, 20_000, 200, seed=1)
dt = time.perf_counter() - t0
print(f"20 000 realisations x 200 OFDM symbols x {len(snrs)} SNR points in {dt * 1e3:.0f} ms")
for r in res:
    print(f"  {r['snr_db']:5.1f} dB  bits={r['total_bits']:.3e}  mean bits/subcarrier={r['total_bits'] / r['num_constellation_symbols']:.2f}  "
          f"BER={r['bit_error_rate']:.3e}  SER={r['symbol_error_rate']:.3e}")
one = run_frames(64, 3, 200, 20.0, n_taps=8, waterfilling=True, min_order=4, max_order=256, seed=1)
print("orders of the first realisation at 20 dB:", one["orders"][0].tolist())
