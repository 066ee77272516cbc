"""Throughput of frame batches (ofdm_frames_run): python tools/bench_frames.py
BASELINE config #4 shape (fresh Rayleigh realisation per frame, water-filling + adaptive loading) and config #2
shape (N=1024 16-QAM MMSE, 8-tap Rayleigh per frame); wall time of the whole synchronous call (tap generation,
water-filling, table build, link kernel, counters D2H)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ofdm-based-systems_b200"))
from ofdm_based_systems import _native as nat

CASES = [
    ("c4 N=64 adaptive WF, 1000 sym/frame", dict(n_subcarriers=64, n_frames=20000, symbols_per_frame=1000, snr_db=20.0)),
    ("c4 N=64 adaptive WF, 64 sym/frame", dict(n_subcarriers=64, n_frames=200000, symbols_per_frame=64, snr_db=20.0)),
    ("c4 N=1024 adaptive WF, 100 sym/frame", dict(n_subcarriers=1024, n_frames=4000, symbols_per_frame=100, snr_db=20.0)),
    ("c2 N=1024 16-QAM Rayleigh/frame, 100 sym/frame", dict(n_subcarriers=1024, n_frames=4000, symbols_per_frame=100, snr_db=14.0, order=16)),
    ("N=4096 256-QAM Rayleigh/frame, 32 sym/frame", dict(n_subcarriers=4096, n_frames=2000, symbols_per_frame=32, snr_db=28.0, order=256)),
]
for name, kw in CASES:
    nat.run_frames(**kw, per_frame=False, want_orders=False, want_taps=False)
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        out = nat.run_frames(**kw, seed=rep, per_frame=False, want_orders=False, want_taps=False)
        best = min(best, time.perf_counter() - t0)
    r = out["total"]
    print(f"{name}: {r.bits:.3e} bits in {best*1e3:.1f} ms = {r.bits/best:.3e} bits/s  BER={r.bit_errors/r.bits:.3e}  "
          f"mean bits/subcarrier={r.bits/r.symbols:.2f}")
