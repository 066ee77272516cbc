"""Fused (Philox) mode on the GPU: the kernel dumps the bits and noise it generated in registers; the
oracle replays exactly those through the reference algorithm and must arrive at the same decisions
and error counts.  Plus distribution checks of the in-register generators and shard invariance."""
import numpy as np
import pytest

import ofdm_oracle as oc
from conftest import load_golden

pytestmark = pytest.mark.gpu

CASES = [
    # name, N, order, scheme, channel, prefix, P, eq, snr, modulator, n_ofdm
    ("headline", 1024, 64, "QAM", "severe_multipath", "CYCLIC", 7, "MMSE", 20.0, "OFDM", 6),
    ("c1", 64, 4, "QAM", "flat_fading", "CYCLIC", 16, "ZF", 6.0, "OFDM", 64),
    ("c2", 1024, 16, "QAM", "severe_multipath", "CYCLIC", 7, "MMSE", 16.0, "OFDM", 6),
    ("c5", 4096, 256, "QAM", "severe_multipath", "CYCLIC", 7, "MMSE", 30.0, "OFDM", 9),
    ("n2048", 2048, 64, "QAM", "rayleigh_fading", "CYCLIC", 5, "MMSE", 24.0, "OFDM", 11),
    ("n2048zf", 2048, 16, "QAM", "severe_multipath", "CYCLIC", 64, "ZF", 14.0, "OFDM", 5),
    ("n128", 128, 16, "QAM", "rayleigh_fading", "CYCLIC", 5, "MMSE", 17.0, "OFDM", 30),
    ("n512", 512, 256, "QAM", "severe_multipath", "CYCLIC", 9, "ZF", 30.0, "OFDM", 12),
    ("zp", 256, 16, "QAM", "rayleigh_fading", "ZERO", 5, "MMSE", 15.0, "OFDM", 16),
    ("zp1024", 1024, 64, "QAM", "severe_multipath", "ZERO", 40, "ZF", 26.0, "OFDM", 6),
    ("isi", 128, 64, "QAM", "severe_multipath", "CYCLIC", 2, "ZF", 24.0, "OFDM", 40),
    ("isi1024", 1024, 16, "QAM", "severe_multipath", "CYCLIC", 3, "MMSE", 19.0, "OFDM", 37),
    ("none", 64, 16, "QAM", "Lin-Phoong_P2", "NONE", 0, "MMSE", 22.0, "OFDM", 48),
    ("sc", 512, 4, "QAM", "Lin-Phoong_P1", "CYCLIC", 3, "ZF", 8.0, "SC-OFDM", 10),
    ("sc4096", 4096, 64, "QAM", "severe_multipath", "CYCLIC", 7, "MMSE", 21.0, "SC-OFDM", 5),
    ("psk", 2048, 8, "PSK", "two_ray", "CYCLIC", 1, "MMSE", 15.0, "OFDM", 4),
    ("psk64", 256, 64, "PSK", "rayleigh_fading", "ZERO", 6, "ZF", 30.0, "OFDM", 16),
    ("bpsk", 64, 2, "PSK", "Lin-Phoong_P1", "CYCLIC", 3, "MMSE", 4.0, "OFDM", 64),
    ("n8192", 8192, 16, "QAM", "default_multipath", "CYCLIC", 3, "MMSE", 18.0, "OFDM", 2),
    ("n16", 16, 4, "QAM", "two_ray", "CYCLIC", 1, "ZF", 10.0, "OFDM", 200),
    ("n32", 32, 16, "QAM", "two_ray", "ZERO", 1, "MMSE", 18.0, "OFDM", 100),
    ("n8", 8, 4, "PSK", "flat_fading", "NONE", 0, "NONE", 8.0, "OFDM", 300),
]


def pack_labels(labels, bps_per_sc):
    """labels[S, N] -> the byte stream the reference's encode() would have consumed."""
    bits = []
    for k, b in enumerate(bps_per_sc):
        if b:
            sh = np.arange(b - 1, -1, -1)
            bits.append(((labels[:, k, None].astype(np.int64) >> sh) & 1))
    return oc.pack_bits(np.concatenate(bits, axis=1).reshape(-1))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_fused_dump_replays_through_oracle(case, kat):
    from ofdm_based_systems._native import Link
    name, n, order, scheme, chan, prefix, P, eq, snr, modulator, n_ofdm = case
    taps_raw = kat["chan_" + chan]
    setup = oc.LinkSetup(n_sc=n, taps_raw=taps_raw, snr_db=snr, order=order, scheme=scheme, modulator=modulator,
                         prefix_type=prefix, eq=eq, prefix_len_override=P)
    sigma = float(np.sqrt(1.0 / 10 ** (snr / 10) / 2))
    link = Link(n, setup.taps_chan, setup.H_eq, np.full(n, order), prefix_type=prefix, prefix_len=P,
                modulator=modulator, equalizer=eq, scheme=scheme)
    assert link.uses_fast_kernel == (name in ("headline", "c1", "c2", "c5", "n2048", "n2048zf", "n128", "n512", "sc", "sc4096", "zp", "zp1024", "isi", "none", "isi1024", "psk", "psk64", "bpsk", "n8192"))
    # with inter-symbol interference the oracle's stream must start where the kernel's does (zero history)
    first = 0 if len(taps_raw) - 1 > P else 1000
    res, d = link.run_fused(snr, sigma, n_ofdm, seed=1234, point=3, first_symbol=first,
                            dump=("z", "rx_labels", "tx_labels", "noise"))
    bps = oc.bits_per_symbol(order)
    tx_bytes = pack_labels(d["tx_labels"], [bps] * n)
    ref = oc.run_link(setup, tx_bytes, n_ofdm * n * bps, noise=d["noise"].astype(np.complex128).reshape(-1))
    z_ref = np.asarray(ref["received_symbols"]).reshape(n_ofdm, n)
    assert np.max(np.abs(d["z"] - z_ref)) / np.max(np.abs(z_ref)) < 1e-5
    f = oc.qam_boundary_distance if scheme == "QAM" else oc.psk_boundary_distance
    mismatch = d["rx_labels"] != np.asarray(ref["rx_labels"]).reshape(n_ofdm, n)
    assert not np.any(mismatch & (f(z_ref, order) > 2e-4))
    if not mismatch.any():
        assert res.bit_errors == ref["bit_errors"]
        assert res.symbol_errors == ref["symbol_errors"]
    assert res.bits == n_ofdm * n * bps
    assert abs(res.papr_db - ref["papr_db"]) < 2e-4
    link.close()


ADAPTIVE_CASES = [
    # name, N, channel, P, eq, snr, n_ofdm, expects the fast kernel
    ("a64", 64, "Lin-Phoong_P1", 3, "MMSE", 24.0, 48, True),
    ("a256", 256, "severe_multipath", 7, "ZF", 26.0, 24, True),
    ("a1024", 1024, "severe_multipath", 7, "MMSE", 22.0, 8, True),
    ("a4096", 4096, "rayleigh_fading", 5, "MMSE", 28.0, 8, True),
    ("a128", 128, "two_ray", 1, "MMSE", 20.0, 32, True),
    ("a512", 512, "Lin-Phoong_P2", 3, "ZF", 25.0, 16, True),
    ("a8192", 8192, "two_ray", 1, "MMSE", 20.0, 2, True),
    ("a32", 32, "two_ray", 1, "MMSE", 20.0, 40, False),
]


@pytest.mark.parametrize("case", ADAPTIVE_CASES, ids=[c[0] for c in ADAPTIVE_CASES])
def test_fused_adaptive_loading_replays_through_oracle(case, kat):
    """Per-subcarrier QAM orders (0 / 4 / 16 / 64 / 256, constellation/adaptive.py:52-201): the kernel's own bits
    and noise, replayed through the reference algorithm, give the same decisions and counts."""
    from ofdm_based_systems._native import Link
    name, n, chan, P, eq, snr, n_ofdm, fast = case
    rng = np.random.default_rng(n)
    orders = rng.choice([0, 4, 16, 64, 256], size=n, p=[0.15, 0.25, 0.25, 0.2, 0.15]).astype(np.int64)
    orders[:5] = [256, 0, 4, 64, 16]
    taps_raw = kat["chan_" + chan]
    setup = oc.LinkSetup(n_sc=n, taps_raw=taps_raw, snr_db=snr, order=16, eq=eq, orders=orders, prefix_len_override=P)
    active_frac = float(np.mean(orders > 1))
    sigma = float(np.sqrt(active_frac / 10 ** (snr / 10) / 2))
    link = Link(n, setup.taps_chan, setup.H_eq, orders, prefix_type="CYCLIC", prefix_len=P, equalizer=eq)
    assert link.uses_fast_kernel == fast
    bps = [oc.bits_per_symbol(int(o)) if o > 1 else 0 for o in orders]
    assert link.bits_per_ofdm_symbol == sum(bps)
    res, d = link.run_fused(snr, sigma, n_ofdm, seed=77, point=2, first_symbol=123456789012,
                            dump=("z", "rx_labels", "tx_labels", "noise"))
    tx_bytes = pack_labels(d["tx_labels"], bps)
    ref = oc.run_link(setup, tx_bytes, n_ofdm * sum(bps), noise=d["noise"].astype(np.complex128).reshape(-1))
    act = orders > 1
    assert np.all(d["tx_labels"][:, ~act] == 0) and np.all(d["rx_labels"][:, ~act] == 0)
    z_ref = np.asarray(ref["received_symbols"]).reshape(n_ofdm, n)
    assert np.max(np.abs(d["z"][:, act] - z_ref[:, act])) / np.max(np.abs(z_ref[:, act])) < 1e-5
    rx_ref = np.asarray(ref["rx_labels"]).reshape(n_ofdm, n)
    dist = np.full(z_ref.shape, np.inf)
    for k in np.nonzero(act)[0]:
        dist[:, k] = oc.qam_boundary_distance(z_ref[:, k], int(orders[k]))
    mismatch = (d["rx_labels"] != np.where(act, rx_ref, 0)) & act
    assert not np.any(mismatch & (dist > 2e-4))
    if not mismatch.any():
        assert res.bit_errors == ref["bit_errors"]
        assert res.symbol_errors == ref["symbol_errors"]
    assert res.bits == n_ofdm * sum(bps) and res.symbols == n_ofdm * n
    assert abs(res.papr_db - ref["papr_db"]) < 2e-4
    link.close()


def test_generators_are_well_distributed(kat):
    from ofdm_based_systems._native import Link
    n, n_ofdm = 1024, 64
    taps = oc.normalize_taps(kat["chan_severe_multipath"])
    link = Link(n, taps, np.fft.fft(taps, n), np.full(n, 64), prefix_type="CYCLIC", prefix_len=7)
    _, d = link.run_fused(10.0, 0.5, n_ofdm, seed=7, dump=("tx_labels", "noise"))
    w = d["noise"][:, 7:].reshape(-1).astype(np.complex128)      # the prefix samples are never generated
    m = w.size
    assert abs(w.real.mean()) < 5 * 0.5 / np.sqrt(m) and abs(w.imag.mean()) < 5 * 0.5 / np.sqrt(m)
    assert abs(w.real.var() / 0.25 - 1) < 0.02 and abs(w.imag.var() / 0.25 - 1) < 0.02
    assert abs(np.mean(w.real * w.imag)) < 5 * 0.25 / np.sqrt(m)
    kurt = np.mean(w.real ** 4) / w.real.var() ** 2
    assert abs(kurt - 3) < 0.1
    assert np.all(d["noise"][:, :7] == 0)
    counts = np.bincount(d["tx_labels"].reshape(-1), minlength=64)
    expected = n * n_ofdm / 64
    assert np.all(np.abs(counts - expected) < 6 * np.sqrt(expected))
    link.close()


def test_sharding_is_invariant(kat):
    """Counters of [0, S) equal the sum over any partition of the symbol range."""
    from ofdm_based_systems._native import Link
    n = 256
    taps = oc.normalize_taps(kat["chan_severe_multipath"])
    for prefix, P in (("CYCLIC", 7), ("CYCLIC", 2), ("ZERO", 3), ("NONE", 0)):
        link = Link(n, taps, np.fft.fft(taps, n), np.full(n, 16), prefix_type=prefix, prefix_len=P)
        sigma = float(np.sqrt(1 / 10 ** 1.4 / 2))
        whole = link.run_fused(14.0, sigma, 900, seed=99, point=1)
        parts = [link.run_fused(14.0, sigma, cnt, seed=99, point=1, first_symbol=start)
                 for start, cnt in ((0, 301), (301, 299), (600, 300))]
        assert whole.bit_errors == sum(p.bit_errors for p in parts) and whole.bit_errors > 0
        assert whole.symbol_errors == sum(p.symbol_errors for p in parts)
        assert whole.bits == sum(p.bits for p in parts)
        assert whole.tx_power_max == max(p.tx_power_max for p in parts)
        assert abs(whole.tx_power_sum - sum(p.tx_power_sum for p in parts)) < 1e-6 * whole.tx_power_sum
        other = link.run_fused(14.0, sigma, 900, seed=100, point=1)
        assert other.bit_errors != whole.bit_errors
        link.close()


@pytest.mark.parametrize("n,chan,P,eq,fast", [(64, "Lin-Phoong_P2", 3, "ZF", True), (1024, "severe_multipath", 7, "MMSE", True),
                                              (128, "rayleigh_fading", 5, "MMSE", True), (8192, "two_ray", 1, "ZF", True), (32, "two_ray", 1, "ZF", False)])
def test_applied_power_loading_replays_through_oracle(n, chan, P, eq, fast, kat):
    """water-filling power APPLIED at the transmitter (sqrt(P_k) on every subcarrier) and compensated at
    the receiver (1/sqrt(P_k)), as examples/waterfilling_noise_bump_experiment.py:148,165-169 does around the
    reference's components; Simulation.run() itself never applies the allocation (simulation/models.py:508)."""
    from ofdm_based_systems._native import Link
    from ofdm_based_systems.simulation.sweep import LinkConfig
    taps_raw, snr, order, n_ofdm = kat["chan_" + chan], 16.0, 16, 8
    gains = np.abs(np.fft.fft(taps_raw, n)) ** 2
    power = np.maximum(oc.waterfilling(1.0, gains, 10 ** (-snr / 10)) * n, 1e-4)      # mean power 1, floored
    amp = np.sqrt(power)
    rx_gain = 1.0 / amp
    setup = oc.LinkSetup(n_sc=n, taps_raw=taps_raw, snr_db=snr, order=order, eq=eq, prefix_len_override=P, amp=amp,
                         rx_gain=rx_gain)
    cfg = LinkConfig(num_subcarriers=n, taps_raw=taps_raw, constellation_order=order, prefix_length=P, equalizator_type=eq,
                     amp=amp, rx_gain=rx_gain)
    link = Link(n, cfg.taps_chan, cfg.h_eq, cfg.orders, prefix_type="CYCLIC", prefix_len=P, equalizer=eq, amp=amp,
                rx_gain=rx_gain)
    assert link.uses_fast_kernel == fast
    res, d = link.run_fused(snr, cfg.noise_sigma(snr), n_ofdm, seed=5, dump=("z", "rx_labels", "tx_labels", "noise"))
    tx_bytes = pack_labels(d["tx_labels"], [4] * n)
    ref = oc.run_link(setup, tx_bytes, n_ofdm * n * 4, noise=d["noise"].astype(np.complex128).reshape(-1))
    z_ref = np.asarray(ref["received_symbols"]).reshape(n_ofdm, n)
    assert np.max(np.abs(d["z"] - z_ref)) / np.max(np.abs(z_ref)) < 1e-5
    mismatch = d["rx_labels"] != np.asarray(ref["rx_labels"]).reshape(n_ofdm, n)
    assert not np.any(mismatch & (oc.qam_boundary_distance(z_ref, order) > 2e-4))
    if not mismatch.any():
        assert res.bit_errors == ref["bit_errors"] and res.symbol_errors == ref["symbol_errors"]
    assert abs(res.papr_db - ref["papr_db"]) < 2e-4
    link.close()


def test_noise_tails_follow_the_gaussian(kat):
    """The in-register Box-Muller generator (32-bit radius word, tail to 6.6 sigma) against the normal tail
    probabilities on 4e7 real samples: a BER sweep to 1e-9 lives on these tails."""
    from scipy.stats import norm
    from ofdm_based_systems._native import Link
    n, P, n_ofdm, sigma = 1024, 7, 19532, 0.37
    taps = oc.normalize_taps(kat["chan_severe_multipath"])
    link = Link(n, taps, np.fft.fft(taps, n), np.full(n, 16), prefix_type="CYCLIC", prefix_len=P)
    _, d = link.run_fused(12.0, sigma, n_ofdm, seed=2024, dump=("noise",))
    link.close()
    w = d["noise"][:, P:].reshape(-1)
    x = np.concatenate([w.real, w.imag]).astype(np.float64) / sigma
    m = x.size
    assert m >= 4e7
    assert abs(x.mean()) < 5 / np.sqrt(m) and abs(x.var() - 1) < 5 * np.sqrt(2 / m)
    for k in (2.0, 3.0, 4.0, 4.5, 5.0):
        p = 2 * norm.sf(k)
        count = int(np.sum(np.abs(x) > k))
        assert abs(count - m * p) < 5 * np.sqrt(m * p) + 1, (k, count, m * p)
    assert np.max(np.abs(x)) < 6.7           # 32-bit radius word: the radius stops at sqrt(2 ln 2^33) = 6.76
    # independence of neighbouring samples and of the two components
    assert abs(np.mean(x[:-1] * x[1:])) < 5 / np.sqrt(m)
    assert abs(np.mean(w.real * w.imag)) / sigma ** 2 < 5 / np.sqrt(m / 2)


def test_noise_stream_is_white_across_calls_lanes_symbols_and_points(kat):
    """The fast kernel draws its noise from Philox4x32 with SEVEN rounds.  Beyond the marginal distribution
    (test above): no correlation between samples of the same Philox call, of neighbouring calls, of neighbouring lanes
    (32 samples apart), of neighbouring OFDM symbols and of two SNR points, fourth and sixth moments of a Gaussian, squares
    uncorrelated too (a weak generator shows up in the squares first), noise independent of the data labels."""
    from ofdm_based_systems._native import Link
    n, P, n_ofdm, sigma = 1024, 7, 6000, 0.21
    taps = oc.normalize_taps(kat["chan_severe_multipath"])
    link = Link(n, taps, np.fft.fft(taps, n), np.full(n, 64), prefix_type="CYCLIC", prefix_len=P)
    _, d0 = link.run_fused(15.0, sigma, n_ofdm, seed=77, point=0, dump=("noise", "tx_labels"))
    _, d1 = link.run_fused(15.0, sigma, n_ofdm, seed=77, point=1, dump=("noise",))
    link.close()
    w0 = d0["noise"][:, P:].astype(np.complex128) / sigma           # [symbol, sample]: sample i of a symbol is word i % 4 of call i // 4
    w1 = d1["noise"][:, P:].astype(np.complex128) / sigma
    m = w0.size
    tol = 5.0 / np.sqrt(m)

    def corr(a, b):
        return abs(np.mean(a * np.conj(b)))

    for lag in (1, 2, 3, 4, 5, 8, 32, 33, 64, 512):                  # inside a call, across calls, across lanes
        assert corr(w0[:, lag:], w0[:, :-lag]) < 1.5 * tol, lag
        assert abs(np.mean(w0[:, lag:] * w0[:, :-lag])) < 1.5 * tol, lag           # non-circular part
    assert corr(w0[1:], w0[:-1]) < 1.5 * tol                          # same sample index, neighbouring OFDM symbols
    assert corr(w0, w1) < 1.5 * tol                                   # same counters, neighbouring SNR points
    x = np.concatenate([w0.real.ravel(), w0.imag.ravel()])
    assert abs(np.mean(x ** 4) - 3.0) < 5 * np.sqrt(96.0 / x.size)    # var(x^4) = 96
    assert abs(np.mean(x ** 6) - 15.0) < 5 * np.sqrt(10170.0 / x.size)
    p0 = np.abs(w0) ** 2 - 2.0                                         # centred powers: E = 0, var = 4
    for lag in (1, 4, 32):
        assert abs(np.mean(p0[:, lag:] * p0[:, :-lag])) < 5 * 4.0 / np.sqrt(m), lag
    assert abs(np.mean(p0 * (np.abs(w1) ** 2 - 2.0))) < 5 * 4.0 / np.sqrt(m)
    lab = d0["tx_labels"].astype(np.float64)
    lab -= lab.mean()
    assert abs(np.mean(lab * w0.real)) < 5 * lab.std() / np.sqrt(m) and abs(np.mean(lab * p0)) < 5 * 2 * lab.std() / np.sqrt(m)
