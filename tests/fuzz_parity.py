"""Randomised differential test on the GPU box: random link shapes through the product kernel (fused mode, dumps), the
same bits and noise replayed through the CPU oracle and through the OTHER CUDA kernel (general), decisions and counters
compared.      python tests/fuzz_parity.py [cases] [seed]
Test infrastructure (it imports oracle/); prints one line per case and a summary, exits non-zero on a mismatch."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ofdm-based-systems_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ofdm_oracle as oc
from ofdm_based_systems import _native as nat
from ofdm_based_systems.simulation.sweep import LinkConfig

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
fast_count = 0
for case in range(cases):
    n = int(rng.choice([64, 128, 256, 512, 1024, 2048, 4096, 8192], p=[.25, .1, .2, .1, .17, .05, .08, .05]))
    scheme = "PSK" if rng.random() < 0.2 else "QAM"
    order = int(rng.choice([2, 4, 8, 16, 32, 64]) if scheme == "PSK" else rng.choice([4, 16, 64, 256]))
    L = int(rng.integers(1, 9))
    taps = (rng.normal(size=L) + 1j * rng.normal(size=L)) * np.exp(-np.arange(L) / 3)
    if rng.random() < 0.5:
        taps /= np.sqrt(np.sum(np.abs(taps) ** 2))            # unit energy like the shipped files, else raw taps (quirk Q3)
    prefix = str(rng.choice(["CYCLIC", "CYCLIC", "ZERO", "NONE"]))
    P = 0 if prefix == "NONE" else int(rng.integers(0, min(n // 2, 24)))
    eq = str(rng.choice(["ZF", "MMSE", "MMSE", "NONE"]))
    modulator = "SC-OFDM" if rng.random() < 0.2 else "OFDM"
    adaptive = scheme == "QAM" and modulator == "OFDM" and rng.random() < 0.25
    snr = float(rng.uniform(5, 35))
    n_ofdm = int(rng.integers(3, 10)) if n >= 1024 else int(rng.integers(8, 40))
    orders = rng.choice([0, 4, 16, 64, 256], size=n).astype(np.int64) if adaptive else np.full(n, order, dtype=np.int64)
    if adaptive:
        orders[0] = 16
    bps = [oc.bits_per_symbol(int(o)) if o > 1 else 0 for o in orders]
    if (sum(bps) * n_ofdm) % 8:
        n_ofdm = 8
    cfg = LinkConfig(num_subcarriers=n, taps_raw=taps, constellation_order=order, constellation_scheme=scheme,
                     modulator_type=modulator, prefix_scheme=prefix, prefix_length=P, equalizator_type=eq,
                     orders=orders if adaptive else None)
    setup = oc.LinkSetup(n_sc=n, taps_raw=taps, snr_db=snr, order=order, scheme=scheme, modulator=modulator, prefix_type=prefix,
                         eq=eq, orders=orders if adaptive else None, prefix_len_override=P)
    tag = f"{case:3d} N={n:4d} {scheme}{'-adapt' if adaptive else order:>6} L={L} {prefix:6s} P={P:2d} {eq:4s} {modulator:7s} snr={snr:4.1f} S={n_ofdm:2d}"
    try:
        link = nat.Link(n, cfg.taps_chan, cfg.h_eq, orders, prefix_type=prefix, prefix_len=P, modulator=modulator, equalizer=eq,
                        scheme=scheme)
    except ValueError as e:
        print(tag, "unsupported:", e)
        continue
    fast = link.uses_fast_kernel
    fast_count += fast
    res, d = link.run_fused(snr, cfg.noise_sigma(snr), n_ofdm, seed=case, first_symbol=0, dump=("z", "rx_labels", "tx_labels", "noise"))
    link.close()
    # the byte stream the reference's encoder would have consumed
    bits = []
    for k, b in enumerate(bps):
        if b:
            bits.append((d["tx_labels"][:, k, None].astype(np.int64) >> np.arange(b - 1, -1, -1)) & 1)
    tx_bytes = oc.pack_bits(np.concatenate(bits, axis=1).reshape(-1))
    noise = d["noise"].astype(np.complex128).reshape(-1)
    ref = oc.run_link(setup, tx_bytes, n_ofdm * sum(bps), noise=noise)
    act = orders > 1
    z_ref = np.asarray(ref["received_symbols"]).reshape(n_ofdm, n)
    rx_ref = np.where(act, np.asarray(ref["rx_labels"]).reshape(n_ofdm, n), 0)
    finite = np.isfinite(z_ref) & act[None, :]
    scale = np.max(np.abs(z_ref[finite])) if finite.any() else 1.0
    zerr = float(np.max(np.abs(np.where(finite, d["z"] - z_ref, 0))) / scale)
    dist = np.full(z_ref.shape, np.inf)
    for k in np.nonzero(act)[0]:
        f = oc.qam_boundary_distance if scheme == "QAM" else oc.psk_boundary_distance
        dist[:, k] = f(z_ref[:, k], int(orders[k]))
    mism = (d["rx_labels"] != rx_ref) & act[None, :]
    far = int(np.sum(mism & (dist > 2e-4)))
    # the other CUDA kernel on the same recorded streams
    os.environ["OFDM_B200_FORCE_GENERAL"] = "1"
    other = nat.Link(n, cfg.taps_chan, cfg.h_eq, orders, prefix_type=prefix, prefix_len=P, modulator=modulator, equalizer=eq, scheme=scheme)
    os.environ.pop("OFDM_B200_FORCE_GENERAL")
    o = other.run_replay(snr, tx_bytes, d["noise"].reshape(-1), n_ofdm)
    other.close()
    ok = (zerr < 1e-5 or eq == "ZF" and zerr < 1e-4) and far == 0 and (mism.any() or (res.bit_errors == ref["bit_errors"] and res.symbol_errors == ref["symbol_errors"])) \
        and abs(o.bit_errors - res.bit_errors) <= 2 + int(mism.sum()) * 8 and abs(res.papr_db - ref["papr_db"]) < 5e-4
    bad += not ok
    print(tag, "fast" if fast else "gen ", f"zerr={zerr:.1e} borderline={int(mism.sum())} far={far} errs={res.bit_errors}/{ref['bit_errors']}/{o.bit_errors}",
          "OK" if ok else "MISMATCH")
print(f"{cases} cases, {fast_count} on the fast kernel, {bad} mismatches")
sys.exit(1 if bad else 0)
