"""BER-vs-SNR tables of the five BASELINE.json configs at their stated sizes, on one GPU:
    python tools/run_baseline_configs.py > gpurun_out/baseline_configs.md
Each point is one launch (LinkSweep / FrameSweep); wall time per config is printed beside the table."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ofdm-based-systems_b200"))
from ofdm_based_systems.simulation.sweep import FrameSweep, LinkConfig, LinkSweep

chan = lambda name: np.load(os.path.join(ROOT, "config", "channel_models", name + ".npy"))


def table(title, snrs, res, dt, extra=""):
    bits = sum(r["total_bits"] for r in res)
    print(f"\n### {title}\n\n{bits:.3e} bits in {dt:.2f} s = {bits / dt:.3e} bits/s (wall, whole sweep){extra}\n")
    print("| SNR (dB) | bits | bit errors | BER | SER | PAPR (dB) |\n|---:|---:|---:|---:|---:|---:|")
    for s, r in zip(snrs, res):
        print(f"| {s:g} | {r['total_bits']:.3e} | {r['bit_errors']} | {r['bit_error_rate']:.3e} | {r['symbol_error_rate']:.3e} | {r['papr_db']:.2f} |")


def link_sweep(title, cfg, snrs, n_symbols, extra=""):
    sw = LinkSweep(cfg)
    extra += f"; fast kernel: {sw.link.uses_fast_kernel}"
    sw.sweep(snrs[:1], 1000)
    t0 = time.perf_counter()
    res = sw.sweep(snrs, n_symbols, seed=2026)
    dt = time.perf_counter() - t0
    sw.close()
    table(title, snrs, res, dt, extra)


print("# BASELINE.json configs on one B200 (tools/run_baseline_configs.py)")
# 1. 64-subcarrier QPSK, CP = 16, AWGN (one-tap) channel, ZF, 0-20 dB
link_sweep("Config 1: N=64 QPSK CP=16 AWGN ZF", LinkConfig(64, chan("flat_fading"), 4, prefix_length=16, equalizator_type="ZF"),
           list(np.arange(0.0, 21.0, 2.0)), 10 ** 9 // 128, "; theory Q(sqrt(SNR))")
# 2. 1024-subcarrier 16-QAM, 8-tap Rayleigh multipath, MMSE, 1e9 bits per SNR point (one realisation: severe_multipath.npy,
#    then a fresh realisation per frame)
link_sweep("Config 2a: N=1024 16-QAM 8-tap multipath (severe_multipath.npy) MMSE, 1e9 bits / point",
           LinkConfig(1024, chan("severe_multipath"), 16, prefix_length=7, equalizator_type="MMSE"),
           list(np.arange(0.0, 31.0, 2.0)), 244141)
snrs = list(np.arange(0.0, 31.0, 5.0))
fs = FrameSweep(1024, n_taps=8, equalizer="MMSE", order=16)
fs.sweep(snrs[:1], 64, 10)
t0 = time.perf_counter()
res = fs.sweep(snrs, 2442, 100, seed=2026)
t_first = time.perf_counter() - t0          # includes growing the library's device arena to this batch size
t0 = time.perf_counter()
res = fs.sweep(snrs, 2442, 100, seed=2026)
table("Config 2b: the same link, fresh 8-tap Rayleigh realisation per frame of 100 OFDM symbols (2 442 frames / point)", snrs, res,
      time.perf_counter() - t0, f"; first sweep of this size in the process: {t_first:.3f} s")
# 3. custom channel models, 64-QAM, ZF vs MMSE
for name in ("Lin-Phoong_P1", "Lin-Phoong_P2", "default_multipath", "two_ray", "rayleigh_fading", "severe_multipath"):
    taps = chan(name)
    for eq in ("ZF", "MMSE"):
        link_sweep(f"Config 3: {name}.npy ({len(taps)} taps), N=64 64-QAM {eq}",
                   LinkConfig(64, taps, 64, prefix_length=len(taps) - 1, equalizator_type=eq), [10.0, 20.0, 30.0], 10 ** 8 // 384)
# 4. water-filling + adaptive loading QPSK..256-QAM, fresh realisation per frame
snrs = [5.0, 10.0, 15.0, 20.0, 25.0, 30.0]
for n in (64, 1024):
    fs = FrameSweep(n, n_taps=8, equalizer="MMSE", waterfilling=True, min_order=4, max_order=256, ser=1e-3)
    fs.sweep(snrs[:1], 64, 10)
    t0 = time.perf_counter()
    res = fs.sweep(snrs, 20000 if n == 64 else 2000, 200 if n == 64 else 100, seed=2026)
    table(f"Config 4: N={n}, water-filling + gap-rule loading (target SER 1e-3), fresh 8-tap Rayleigh realisation per frame", snrs,
          res, time.perf_counter() - t0, "; bits per point grow with the SNR because the loading does")
# 5. 4096-subcarrier 256-QAM MMSE sweep to BER 1e-9: 1e11 bits per point, then 1e12 bits where the curve crosses 1e-9
cfg5 = LinkConfig(4096, chan("severe_multipath"), 256, prefix_length=7, equalizator_type="MMSE")
link_sweep("Config 5: N=4096 256-QAM MMSE severe_multipath.npy, 1e11 bits / point", cfg5,
           [20.0, 24.0, 28.0, 32.0, 34.0, 36.0, 38.0, 40.0], 10 ** 11 // 32768)
link_sweep("Config 5: the same link at the BER 1e-9 crossing, 1e12 bits / point", cfg5, [39.0, 39.5], 10 ** 12 // 32768)
