"""Single-config launcher for ncu: python tools/profile_fused.py [N] [order] [symbols] [launches]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ofdm-based-systems_b200"))
from ofdm_based_systems import _native as nat
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
order = int(sys.argv[2]) if len(sys.argv) > 2 else 64
nsym = int(sys.argv[3]) if len(sys.argv) > 3 else 162761
launches = int(sys.argv[4]) if len(sys.argv) > 4 else 3
taps = np.load(os.path.join(ROOT, "config", "channel_models", "severe_multipath.npy"))
taps = taps / np.sqrt(np.sum(np.abs(taps) ** 2))
link = nat.Link(n, taps, np.fft.fft(taps, n), np.full(n, order), prefix_type="CYCLIC", prefix_len=7, equalizer="MMSE")
sigma = float(np.sqrt(1 / 10 ** 2.0 / 2))
for i in range(launches):
    r = link.run_fused(20.0, sigma, nsym, seed=i)
print(r.bits, r.bit_errors / r.bits)
