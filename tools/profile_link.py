"""Launcher for ncu captures and wall-clock A/B timing of ONE link shape:

    python tools/profile_link.py --n 64 --order 4 --taps flat_fading --prefix 16 --eq ZF [--modulator SC-OFDM]
                                 [--bits 1e9] [--launches 3] [--time REPS] [--points K]

--time REPS prints the mean time of REPS back-to-back device-side launches (host clock around launch + one read-back,
so launch latency is included once), the bits/s and the algorithmic TFLOP/s with SURVEY 8(d)'s flop count.
--points K runs K SNR points per launch (the sweep entry).
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ofdm-based-systems_b200"))
from ofdm_based_systems import _native as nat  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1024)
ap.add_argument("--order", type=int, default=64)
ap.add_argument("--taps", default="severe_multipath")
ap.add_argument("--prefix", type=int, default=-1)
ap.add_argument("--prefix-type", default="CYCLIC")
ap.add_argument("--eq", default="MMSE")
ap.add_argument("--modulator", default="OFDM")
ap.add_argument("--scheme", default="QAM")
ap.add_argument("--snr", type=float, default=20.0)
ap.add_argument("--bits", type=float, default=1e9)
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--time", type=int, default=0)
ap.add_argument("--points", type=int, default=1)
ap.add_argument("--adaptive", action="store_true", help="per-subcarrier orders drawn from {0, 4, 16, 64, 256} (loading tables)")
a = ap.parse_args()

taps = np.load(os.path.join(ROOT, "config", "channel_models", a.taps + ".npy")).astype(np.complex128)
L = len(taps)
prefix = L - 1 if a.prefix < 0 else a.prefix
bps = int(np.log2(a.order))
nsym = int(-(-a.bits // (a.n * bps)))
h_eq = np.fft.fft(taps, a.n)
tn = taps / np.sqrt(np.sum(np.abs(taps) ** 2))
orders = np.random.default_rng(1).choice([0, 4, 16, 64, 256], size=a.n) if a.adaptive else np.full(a.n, a.order)
if a.adaptive:
    bps = float(np.mean([int(np.log2(o)) if o > 1 else 0 for o in orders]))
    nsym = int(-(-a.bits // (a.n * bps)))
link = nat.Link(a.n, tn, h_eq, orders, prefix_type=a.prefix_type, prefix_len=prefix, equalizer=a.eq,
                modulator=a.modulator, scheme=a.scheme)
sigma = float(np.sqrt(1 / 10 ** (a.snr / 10) / 2))
c_eq = {"MMSE": 14, "ZF": 6, "NONE": 0}[a.eq]
f_sym = 10 * a.n * int(np.log2(a.n)) + 8 * L * (a.n + prefix) + 4 * (a.n + prefix) + a.n * (2 + c_eq + 8)
snrs = [a.snr + 0.5 * i for i in range(a.points)]
sigs = [float(np.sqrt(1 / 10 ** (s / 10) / 2)) for s in snrs]


def go(seed, sync):
    if a.points > 1:
        if sync:
            return link.run_sweep(snrs, sigs, nsym, seed=seed)[0]
        link.launch_sweep(snrs, sigs, nsym, seed=seed)
        return None
    if sync:
        return link.run_fused(a.snr, sigma, nsym, seed=seed)
    link.launch_fused(a.snr, sigma, nsym, seed=seed)
    return None


for i in range(a.launches):
    r = go(i, True)
print(f"fast_kernel={link.uses_fast_kernel} N={a.n} M={a.order} L={L} P={prefix} {a.eq} {a.modulator}: {r.bits} bits, BER {r.bit_errors / max(r.bits, 1):.5f}")
if a.time:
    t0 = time.perf_counter()
    for i in range(a.time):
        go(100 + i, False)
    if a.points > 1:
        link.read_sweep(a.points)
    else:
        link.read_result()
    dt = (time.perf_counter() - t0) / a.time
    bits = int(nsym * a.n * bps * a.points)
    print(f"TIME N={a.n} M={a.order} L={L} P={prefix} {a.eq} {a.modulator} points={a.points}: {dt * 1e3:.4f} ms/launch "
          f"{bits / dt:.4e} bits/s  F_sym={f_sym}  alg {f_sym * nsym * a.points / dt / 1e12:.2f} TFLOP/s "
          f"({100 * f_sym * nsym * a.points / dt / 1e12 / 72.6:.1f} % of 72.6)")
