"""Tiny pass through every kernel variant for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
(On this round's pool compute-sanitizer was closed by the operators; the plain run is still a quick pass through every
instantiation, and tests/test_link_edges_gpu.py checks run-to-run determinism of every variant instead.)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ofdm-based-systems_b200"))
from ofdm_based_systems import _native as nat

kat = np.load(os.path.join(ROOT, "tests", "golden", "kat.npz"))
rng = np.random.default_rng(0)


def link(n, order, chan, prefix, P, eq, modulator="OFDM", scheme="QAM", orders=None, **kw):
    taps = kat["chan_" + chan]
    tn = taps / np.sqrt(np.sum(np.abs(taps) ** 2))
    return nat.Link(n, tn, np.fft.fft(taps, n), np.full(n, order) if orders is None else orders, prefix_type=prefix,
                    prefix_len=P, equalizer=eq, modulator=modulator, scheme=scheme, **kw)


def exercise(l, n, P, bps, n_sym=5):
    r, _ = l.run_fused(15.0, 0.1, n_sym, seed=1, dump=("y", "z", "rx_labels", "tx_labels", "noise"))
    l.run_fused(15.0, 0.1, 40, seed=2)
    if bps:
        bits = rng.integers(0, 256, n_sym * n * bps // 8, dtype=np.uint8)
        noise = ((rng.normal(size=n_sym * (n + P)) + 1j * rng.normal(size=n_sym * (n + P))) * 0.1).astype(np.complex64)
        l.run_replay(15.0, bits.tobytes(), noise, n_sym, dump=("z", "rx_labels"))
        l.run_replay(15.0, bits.tobytes(), noise.astype(np.complex128), n_sym)
    print("ok", n, l.uses_fast_kernel, r.bits, r.bit_errors)
    l.close()


for n in (64, 128, 256, 512, 1024, 2048, 4096):
    exercise(link(n, 16, "severe_multipath", "CYCLIC", 7, "MMSE"), n, 7, 4)
exercise(link(256, 64, "severe_multipath", "ZERO", 9, "ZF"), 256, 9, 6)
exercise(link(1024, 4, "Lin-Phoong_P1", "CYCLIC", 3, "ZF", modulator="SC-OFDM"), 1024, 3, 2)
exercise(link(128, 16, "severe_multipath", "CYCLIC", 2, "MMSE"), 128, 2, 4, n_sym=9)          # ISI
exercise(link(2048, 64, "severe_multipath", "NONE", 0, "MMSE"), 2048, 0, 6, n_sym=7)          # ISI, 2-warp teams
exercise(link(512, 8, "two_ray", "CYCLIC", 1, "MMSE", scheme="PSK"), 512, 1, 3, n_sym=8)
orders = rng.choice([0, 4, 16, 64, 256], size=1024)
exercise(link(1024, 0, "rayleigh_fading", "CYCLIC", 5, "MMSE", orders=orders), 1024, 5, 0)
exercise(link(64, 16, "severe_multipath", "CYCLIC", 40, "MMSE"), 64, 40, 4)                    # long prefix
exercise(link(32, 16, "two_ray", "ZERO", 1, "MMSE"), 32, 1, 4)                                 # general kernel
exercise(link(256, 64, "severe_multipath", "ZERO", 3, "MMSE"), 256, 3, 6, n_sym=9)             # zero padding shorter than the channel
exercise(link(128, 4, "Lin-Phoong_P2", "ZERO", 1, "MMSE", modulator="SC-OFDM"), 128, 1, 2, n_sym=9)
exercise(link(64, 4, "flat_fading", "CYCLIC", 16, "ZF"), 64, 16, 2)                            # one tap, no noise estimate
exercise(link(1024, 64, "severe_multipath", "CYCLIC", 7, "ZF"), 1024, 7, 6)
pl = link(64, 64, "Lin-Phoong_P2", "CYCLIC", 3, "MMSE", amp=np.full(64, 0.125), rx_gain=np.full(64, 8.0))   # post-equaliser stage
res = pl.run_fused_renormalised(20.0, 0.0, 50, noise_profile=np.linspace(1.0, 2.0, 64), seed=3)
print("post ok", res.bits, res.bit_errors)
pl.close()
w8 = nat.waterfill_bitload_batched(np.tile(kat["chan_severe_multipath"][None, :], (11, 1)), 64, 15.0)   # one warp per realisation
out = nat.run_frames(256, 6, 20, 18.0, n_taps=8, waterfilling=True)
out = nat.run_frames(4096, 3, 9, 18.0, n_taps=8, order=64)
w = nat.waterfill_bitload_batched(kat["chan_severe_multipath"][None, :], 256, 15.0)
print("frames ok", out["total"].bits)
