"""Sustained-load check: the headline launch back to back for ~20 s, clocks / power / throttle reasons sampled with
nvidia-smi, throughput per second of wall time.  python tools/sustained.py [seconds]"""
import os, subprocess, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ofdm-based-systems_b200"))
from ofdm_based_systems import _native as nat

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
taps = np.load(os.path.join(ROOT, "config", "channel_models", "severe_multipath.npy"))
taps = taps / np.sqrt(np.sum(np.abs(taps) ** 2))
link = nat.Link(1024, taps, np.fft.fft(taps, 1024), np.full(1024, 64), prefix_type="CYCLIC", prefix_len=7, equalizer="MMSE")
sigma = float(np.sqrt(1 / 10 ** 2.0 / 2))
rows = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap,"
                         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown", "--format=csv,noheader,nounits",
                         "-lms", "500", "-i", "0"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append(l.strip()) for l in proc.stdout], daemon=True).start()
link.run_fused(20.0, sigma, 162761)
t0 = time.perf_counter()
marks = []
n = 0
while time.perf_counter() - t0 < seconds:
    for _ in range(200):
        link.launch_fused(20.0, sigma, 162761, seed=n)
        n += 1
    r = link.read_result()
    marks.append((time.perf_counter() - t0, n))
proc.terminate()
prev_t, prev_n = 0.0, 0
rates = []
for t, k in marks:
    rates.append((k - prev_n) * 162761 * 6144 / (t - prev_t))
    prev_t, prev_n = t, k
print(f"{n} launches in {marks[-1][0]:.1f} s: {n * 162761 * 6144 / marks[-1][0]:.4e} bits/s sustained; per 200-launch window "
      f"min {min(rates):.4e} max {max(rates):.4e}; BER {r.bit_errors / r.bits:.5f}")
print("nvidia-smi samples (sm MHz, W, C, sw_power_cap, hw_slowdown, sw_thermal):")
for row in rows[:: max(1, len(rows) // 12)]:
    print("  ", row)
