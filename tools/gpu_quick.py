"""Developer probe run on the GPU box: FP32 peak and a first throughput number."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ofdm-based-systems_b200"))
from ofdm_based_systems import _native as nat

print("devices", nat.device_count())
print("fp32 TFLOP/s (FFMA chain)", nat.measure_fp32_tflops(8192))
kat = np.load(os.path.join(ROOT, "tests", "golden", "kat.npz"))
taps = kat["chan_severe_multipath"]; taps = taps / np.sqrt(np.sum(np.abs(taps) ** 2))
for n, order, P in ((1024, 64, 7), (64, 4, 16), (4096, 256, 7), (256, 16, 7)):
    link = nat.Link(n, taps, np.fft.fft(taps, n), np.full(n, order), prefix_type="CYCLIC", prefix_len=P, equalizer="MMSE")
    bps = int(np.log2(order))
    nsym = int(2e9 // (n * bps))
    sigma = float(np.sqrt(1 / 10 ** 2.0 / 2))
    link.run_fused(20.0, sigma, 1000)
    for rep in range(3):
        t0 = time.perf_counter()
        r = link.run_fused(20.0, sigma, nsym, seed=rep)
        dt = time.perf_counter() - t0
        flops = {1024: 197084, 4096: 868828}.get(n, 0) * nsym
        print(f"N={n} M={order}: {nsym} symbols {r.bits/dt:.3e} bits/s  BER={r.bit_errors/r.bits:.5f} "
              f"wall={dt*1e3:.1f} ms  alg TFLOP/s={flops/dt/1e12:.2f}")
    link.close()
