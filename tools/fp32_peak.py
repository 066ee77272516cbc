"""The roofline denominator with its own clock record: the FFMA-chain microbenchmark (ofdm_b200_measure_fp32_tflops)
back to back for a few seconds, nvidia-smi clocks / power / throttle reasons sampled meanwhile.
    python tools/fp32_peak.py [seconds]"""
import os, subprocess, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ofdm-based-systems_b200"))
from ofdm_based_systems import _native as nat

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
rows = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap,"
                         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown", "--format=csv,noheader,nounits",
                         "-lms", "200", "-i", "0"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append(l.strip()) for l in proc.stdout], daemon=True).start()
vals = []
t0 = time.perf_counter()
while time.perf_counter() - t0 < seconds:
    vals.append(nat.measure_fp32_tflops(8192))      # best of 4 timed launches of 8192 x 128 FFMA per thread, 148 x 8 blocks x 256 threads
proc.terminate()
print(f"{len(vals)} measurements in {time.perf_counter() - t0:.1f} s: min {min(vals):.2f} median {sorted(vals)[len(vals) // 2]:.2f} max {max(vals):.2f} TFLOP/s "
      f"(nominal 148 SMs x 128 lanes x 2 x 1.965 GHz = 74.45)")
print("nvidia-smi samples (sm MHz, max sm MHz, W, C, sw_power_cap, hw_slowdown, sw_thermal):")
for row in rows[:: max(1, len(rows) // 12)]:
    print("  ", row)
