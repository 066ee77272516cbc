"""Time the fused headline launch with the device-side API: python tools/time_fused.py [N] [order] [symbols] [reps] [OFDM|SC-OFDM]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ofdm-based-systems_b200"))
from ofdm_based_systems import _native as nat
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
order = int(sys.argv[2]) if len(sys.argv) > 2 else 64
nsym = int(sys.argv[3]) if len(sys.argv) > 3 else 162761
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
modulator = sys.argv[5] if len(sys.argv) > 5 else "OFDM"
taps = np.load(os.path.join(ROOT, "config", "channel_models", "severe_multipath.npy"))
taps = taps / np.sqrt(np.sum(np.abs(taps) ** 2))
link = nat.Link(n, taps, np.fft.fft(taps, n), np.full(n, order), prefix_type="CYCLIC", prefix_len=7, equalizer="MMSE", modulator=modulator)
sigma = float(np.sqrt(1 / 10 ** 2.0 / 2))
for i in range(3):
    link.run_fused(20.0, sigma, nsym, seed=i)
link.reset_counters()
t0 = time.perf_counter()
for i in range(reps):
    link.launch_fused(20.0, sigma, nsym, seed=10 + i)
r = link.read_result()
dt = (time.perf_counter() - t0) / reps
bps = int(np.log2(order))
F = {1024: 197084, 4096: 868828}.get(n, 0)
print(f"variant={os.environ.get('OFDM_B200_FAST_VARIANT','0')} N={n} M={order}: {dt*1e3:.3f} ms/launch  {nsym*n*bps/dt:.4e} bits/s  "
      f"alg {F*nsym/dt/1e12:.2f} TFLOP/s  BER={r.bit_errors/r.bits:.5f}")
