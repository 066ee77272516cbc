"""Edge cases of the C ABI on the GPU: empty ranges, ragged recorded streams, argument errors, 64-bit symbol
indices, and size-independent properties at BASELINE.json's full sizes (linearity of the counters in the symbol
range, noiseless links decode without error, counters scale with the bits processed)."""
import numpy as np
import pytest

import ofdm_oracle as oc

pytestmark = pytest.mark.gpu


def headline_link(order=64, n=1024, eq="MMSE", **kw):
    from ofdm_based_systems._native import Link
    taps = oc.normalize_taps(np.load(__import__("os").path.join(__import__("conftest").ROOT, "config", "channel_models",
                                                                    "severe_multipath.npy")))
    return Link(n, taps, np.fft.fft(taps, n), np.full(n, order), prefix_type="CYCLIC", prefix_len=7, equalizer=eq, **kw)


def test_empty_symbol_range_and_zero_noise():
    link = headline_link()
    r = link.run_fused(20.0, 0.1, 0)
    assert (r.bits, r.bit_errors, r.symbols, r.ofdm_symbols, r.tx_samples) == (0, 0, 0, 0, 0)
    assert r.papr_db == float("inf")
    r = link.run_fused(200.0, 0.0, 500, seed=3)                     # noiseless: every decision is right
    assert r.bits == 500 * 6144 and r.bit_errors == 0 and r.symbol_errors == 0
    assert 8.0 < r.papr_db < 16.0
    link.close()


@pytest.mark.parametrize("kernel", ["auto", "general"])
def test_ragged_recorded_stream(kernel, monkeypatch):
    """A recorded byte stream that ends inside the last OFDM symbol: positions past compare_limit_bits are not compared
    (zip() truncation, simulation/models.py:597); the missing bits transmit as zeros."""
    if kernel == "general":
        monkeypatch.setenv("OFDM_B200_FORCE_GENERAL", "1")
    link = headline_link(order=16)
    rng = np.random.default_rng(5)
    n_sym, bits_per = 5, 4096
    full = rng.integers(0, 256, n_sym * bits_per // 8, dtype=np.uint8)
    cut = full[: len(full) - 100]                                     # 800 bits short
    noise = (rng.normal(size=n_sym * 1031) + 1j * rng.normal(size=n_sym * 1031)) * 0.05
    from ofdm_based_systems import _native
    whole = link.run_replay(20.0, full.tobytes(), noise, n_sym)
    before = _native.launch_count()
    ragged = link.run_replay(20.0, cut.tobytes(), noise, n_sym, compare_limit_bits=8 * len(cut))
    # a fast-kernel link streams the whole symbols through the fast kernel and runs only the last one on the general kernel
    assert _native.launch_count() - before == (2 if kernel == "auto" else 1)
    assert whole.bits == n_sym * bits_per and ragged.bits == 8 * len(cut)
    assert ragged.bit_errors <= whole.bit_errors + 40                 # the zero-filled tail changes only the last symbol
    none = link.run_replay(20.0, full.tobytes(), None, n_sym)
    assert none.bit_errors == 0
    with pytest.raises(ValueError):
        link.run_replay(20.0, full.tobytes(), noise[:-1], n_sym)
    with pytest.raises(ValueError):
        link.run_replay(20.0, full.tobytes(), noise.real.copy(), n_sym)
    link.close()


def test_argument_errors_come_back_as_value_errors():
    from ofdm_based_systems._native import Link
    h = np.ones(1, complex)
    ok = dict(prefix_type="CYCLIC", prefix_len=0)
    with pytest.raises(ValueError, match="power of two"):
        Link(100, h, np.ones(100, complex), np.full(100, 4), **ok)
    with pytest.raises(ValueError, match="perfect square"):
        Link(64, h, np.ones(64, complex), np.full(64, 8), **ok)
    with pytest.raises(ValueError, match="n_taps"):
        Link(64, np.ones(33, complex), np.ones(64, complex), np.full(64, 4), **ok)
    with pytest.raises(ValueError, match="prefix NONE"):
        Link(64, h, np.ones(64, complex), np.full(64, 4), prefix_type="NONE", prefix_len=3)
    with pytest.raises(ValueError, match="one entry per subcarrier"):
        Link(64, h, np.ones(63, complex), np.full(64, 4), **ok)
    with pytest.raises(ValueError, match="No active subcarriers"):
        Link(64, h, np.ones(64, complex), np.zeros(64, int), **ok).run_fused(10.0, 0.1, 4)
    with pytest.raises(KeyError):
        Link(64, h, np.ones(64, complex), np.full(64, 4), prefix_type="BOGUS")


@pytest.mark.parametrize("n,order,per_launch", [(1024, 16, 244_141), (4096, 256, 30_518)])
def test_full_size_counters_are_additive_and_reproducible(n, order, per_launch):
    """BASELINE configs #2 / #5 at 1e9 bits per point: the counters of [0, S) equal the sum over a 3-way split
    (what sharding over GPUs relies on), a re-run reproduces them exactly, 64-bit symbol indices work."""
    link = headline_link(order=order, n=n)
    bps = oc.bits_per_symbol(order)
    sigma = float(np.sqrt(1 / 10 ** 2.4 / 2))
    whole = link.run_fused(24.0, sigma, per_launch, seed=42, point=3)
    assert whole.bits == per_launch * n * bps >= 10 ** 9
    again = link.run_fused(24.0, sigma, per_launch, seed=42, point=3)
    assert (again.bit_errors, again.symbol_errors, again.tx_power_max) == (whole.bit_errors, whole.symbol_errors, whole.tx_power_max)
    cuts = [0, per_launch // 3, per_launch // 3 * 2 + 1, per_launch]
    parts = [link.run_fused(24.0, sigma, b - a, seed=42, point=3, first_symbol=a) for a, b in zip(cuts, cuts[1:])]
    assert whole.bit_errors == sum(p.bit_errors for p in parts) > 0
    assert whole.symbol_errors == sum(p.symbol_errors for p in parts)
    clean = link.run_fused(300.0, 0.0, per_launch, seed=43)       # map -> IFFT -> FIR -> FFT -> MMSE -> demap round trip
    assert clean.bits == whole.bits and clean.bit_errors == 0 and clean.symbol_errors == 0
    far = link.run_fused(24.0, sigma, 2000, seed=42, point=3, first_symbol=(1 << 40) + 7)
    assert far.bits == 2000 * n * bps and 0 < far.bit_errors < far.bits // 4
    # BER of the two halves agree within their sampling noise (per-symbol correlated errors: generous 6 sigma)
    ber = [p.bit_errors / p.bits for p in parts]
    assert max(ber) - min(ber) < 6 * np.sqrt(max(ber) / (parts[0].bits / (n * bps))) + 1e-9
    link.close()


def test_no_noise_sample_poisons_an_ofdm_symbol():
    """1e12 bits of BASELINE config #5 (N = 4096, 256-QAM, MMSE) at an SNR where thermal errors are impossible: every noise
    word the generator can produce must give a finite sample.  A radius taken straight from the 32-bit word (lg2(0) for w = 0,
    once per 2^32 samples) turned ~29 OFDM symbols of such a run into NaNs - 16 384 bit errors each (simulation/models.py:597-606
    counts them like any other)."""
    link = headline_link(order=256, n=4096)
    sigma = float(np.sqrt(1 / 10 ** 5.5 / 2))
    r = link.run_fused(55.0, sigma, 30_518_000, seed=0x0FD3)
    link.close()
    assert r.bits >= 10 ** 12 and r.bit_errors == 0 and r.symbol_errors == 0


VARIANTS = [
    # name, N, order, scheme, channel, prefix, P, eq, modulator, per-subcarrier orders?
    ("qam", 1024, 64, "QAM", "severe_multipath", "CYCLIC", 7, "MMSE", "OFDM", False),
    ("wide", 4096, 16, "QAM", "severe_multipath", "CYCLIC", 7, "MMSE", "OFDM", False),
    ("narrow", 128, 16, "QAM", "rayleigh_fading", "CYCLIC", 5, "ZF", "OFDM", False),
    ("zp", 512, 64, "QAM", "severe_multipath", "ZERO", 9, "MMSE", "OFDM", False),
    ("sc", 2048, 16, "QAM", "Lin-Phoong_P1", "CYCLIC", 3, "MMSE", "SC-OFDM", False),
    ("isi", 256, 16, "QAM", "severe_multipath", "CYCLIC", 2, "MMSE", "OFDM", False),
    ("isi_wide", 2048, 64, "QAM", "severe_multipath", "NONE", 0, "ZF", "OFDM", False),
    ("psk", 1024, 8, "PSK", "two_ray", "CYCLIC", 1, "MMSE", "OFDM", False),
    ("adaptive", 1024, 0, "QAM", "severe_multipath", "CYCLIC", 7, "MMSE", "OFDM", True),
    ("general", 32, 16, "QAM", "two_ray", "CYCLIC", 1, "MMSE", "OFDM", False),        # N < 64: the general kernel
    ("widest", 8192, 16, "QAM", "two_ray", "CYCLIC", 1, "MMSE", "OFDM", False),        # two teams of eight warps, radix-8 third pass
    # the headline shape's short-channel instantiations (1 and 4 evaluated taps) and the small transforms of the shipped configs
    ("flat", 64, 16, "QAM", "flat_fading", "CYCLIC", 16, "ZF", "OFDM", False),
    ("four_taps", 64, 64, "QAM", "Lin-Phoong_P1", "CYCLIC", 3, "MMSE", "OFDM", False),
    ("four_taps_256", 256, 16, "QAM", "default_multipath", "CYCLIC", 3, "ZF", "OFDM", False),
    ("sc_isi", 512, 16, "QAM", "severe_multipath", "CYCLIC", 3, "MMSE", "SC-OFDM", False),
    # the instantiations without the per-symbol noise estimate (ZF / no equaliser) at 8 taps, the one-tap MMSE one, and
    # zero padding shorter than the channel (partial overlap-add + leak into the next symbol), OFDM and SC-OFDM
    ("zf8", 1024, 64, "QAM", "severe_multipath", "CYCLIC", 7, "ZF", "OFDM", False),
    ("flat_mmse", 256, 64, "QAM", "flat_fading", "CYCLIC", 4, "MMSE", "OFDM", False),
    ("isi_zp", 256, 64, "QAM", "severe_multipath", "ZERO", 3, "MMSE", "OFDM", False),
    ("isi_zp_sc", 128, 4, "QAM", "Lin-Phoong_P2", "ZERO", 1, "MMSE", "SC-OFDM", False),
    # the 4-tap instantiations of the loading and the chained-symbol kernels (the dump-capable twin evaluates 8 taps)
    ("adaptive_4taps", 64, 0, "QAM", "Lin-Phoong_P1", "CYCLIC", 3, "MMSE", "OFDM", True),
    ("isi_4taps", 64, 64, "QAM", "Lin-Phoong_P2", "CYCLIC", 1, "MMSE", "OFDM", False),
    # combinations that ran on the general kernel until the end of round 2: PSK on single-carrier symbols, PSK or loading
    # tables with a prefix shorter than the channel memory
    ("psk_sc", 256, 8, "PSK", "Lin-Phoong_P1", "CYCLIC", 3, "MMSE", "SC-OFDM", False),
    ("psk_isi", 128, 16, "PSK", "severe_multipath", "CYCLIC", 2, "ZF", "OFDM", False),
    ("psk_sc_isi", 64, 16, "PSK", "rayleigh_fading", "NONE", 0, "MMSE", "SC-OFDM", False),
    ("adaptive_isi", 512, 0, "QAM", "severe_multipath", "CYCLIC", 3, "MMSE", "OFDM", True),
    ("adaptive_isi_zp", 64, 0, "QAM", "rayleigh_fading", "ZERO", 2, "MMSE", "OFDM", True),
]


def variant_link(v, kat):
    from ofdm_based_systems._native import Link
    name, n, order, scheme, chan, prefix, P, eq, modulator, adaptive = v
    taps = kat["chan_" + chan]
    tn = oc.normalize_taps(taps)
    orders = np.random.default_rng(1).choice([0, 4, 16, 64, 256], size=n) if adaptive else np.full(n, order)
    return Link(n, tn, np.fft.fft(taps, n), orders, prefix_type=prefix, prefix_len=P, equalizer=eq, modulator=modulator,
                scheme=scheme)


@pytest.mark.parametrize("v", VARIANTS, ids=[v[0] for v in VARIANTS])
def test_counter_only_kernel_equals_the_dump_kernel(v, kat):
    """The instantiation bench.py times (DUMP = false, for short channels also fewer evaluated taps) is a different
    compilation of the same source than the one the oracle comparisons run (DUMP = true).  Same seed, same symbol range:
    every counter - bit errors, symbol errors, power sum, power maximum - must be IDENTICAL, which pins the benchmarked
    kernels to the oracle through tests/test_link_fused_gpu.py (simulation/models.py:597-606)."""
    link = variant_link(v, kat)
    n = v[1]
    S = max(64, 400_000 // n)
    for seed, first in ((8, 0), (0x0FD3, 12_345)):
        plain = link.run_fused(17.0, 0.09, S, seed=seed, point=2, first_symbol=first)
        dumped, _ = link.run_fused(17.0, 0.09, S, seed=seed, point=2, first_symbol=first, dump=("z",))
        assert same_counters(plain, dumped) and plain.bit_errors > 0
    link.close()


def same_counters(a, b) -> bool:
    """Every integer counter and the power maximum identical; the power SUM is a double accumulated by atomics in
    launch-dependent order, so it agrees to rounding only."""
    ints = ("bit_errors", "bits", "symbol_errors", "symbols", "ofdm_symbols", "tx_samples", "tx_power_max")
    return all(getattr(a, k) == getattr(b, k) for k in ints) and abs(a.tx_power_sum - b.tx_power_sum) <= 1e-12 * abs(a.tx_power_sum)


@pytest.mark.parametrize("v", VARIANTS, ids=[v[0] for v in VARIANTS])
def test_sweep_launch_equals_per_point_launches(v, kat):
    """ofdm_link_run_sweep (ONE launch, the SNR point is the grid's second dimension; main.py:234-240 as a single call)
    returns, for point i, exactly the counters of ofdm_link_run_fused(point = first_point + i) over the same symbols."""
    link = variant_link(v, kat)
    n = v[1]
    S = max(48, 200_000 // n)
    snrs = [6.0, 12.0, 18.0, 24.0, 30.0]
    sigmas = [0.30, 0.20, 0.12, 0.07, 0.04]
    swept = link.run_sweep(snrs, sigmas, S, seed=77, first_point=3, first_symbol=1000)
    assert len(swept) == 5
    for i, got in enumerate(swept):
        single = link.run_fused(snrs[i], sigmas[i], S, seed=77, point=3 + i, first_symbol=1000)
        assert same_counters(got, single), (i, got, single)
    assert swept[0].bit_errors > swept[-1].bit_errors
    link.close()


def test_sweep_of_more_points_than_one_launch_carries(kat):
    """40 SNR points: queued as launches of 32 + 8 points into one block of counters; empty and invalid sweeps."""
    link = headline_link(order=16, n=256)
    snrs = list(np.linspace(0.0, 39.0, 40))
    sigmas = [float(np.sqrt(10 ** (-x / 10) / 2)) for x in snrs]
    swept = link.run_sweep(snrs, sigmas, 400, seed=5)
    for i in (0, 31, 32, 39):
        assert same_counters(swept[i], link.run_fused(snrs[i], sigmas[i], 400, seed=5, point=i))
    empty = link.run_sweep(snrs[:3], sigmas[:3], 0)
    assert all(r.bits == 0 and r.bit_errors == 0 for r in empty)
    with pytest.raises(ValueError):
        link.run_sweep([], [], 10)
    with pytest.raises(ValueError):
        link.run_sweep([1.0, 2.0], [0.1], 10)
    link.close()


@pytest.mark.parametrize("v", VARIANTS, ids=[v[0] for v in VARIANTS])
def test_every_variant_is_deterministic_and_additive(v, kat):
    """Same seed -> identical counters run to run (a shared-memory race or an uninitialised read would show up here), and
    the counters of [0, S) equal the sum over a split of the symbol range for every kernel variant."""
    name, n = v[0], v[1]
    link = variant_link(v, kat)
    assert link.uses_fast_kernel == (name != "general")
    S = max(300, 2_000_000 // n)
    runs = [link.run_fused(16.0, 0.11, S, seed=8, point=1, first_symbol=77) for _ in range(3)]
    for r in runs[1:]:
        assert (r.bit_errors, r.symbol_errors, r.bits, r.tx_power_max) == (runs[0].bit_errors, runs[0].symbol_errors,
                                                                           runs[0].bits, runs[0].tx_power_max)
    assert 0 < runs[0].bit_errors < runs[0].bits // 3
    a = link.run_fused(16.0, 0.11, S // 3, seed=8, point=1, first_symbol=77)
    b = link.run_fused(16.0, 0.11, S - S // 3, seed=8, point=1, first_symbol=77 + S // 3)
    assert a.bit_errors + b.bit_errors == runs[0].bit_errors and a.symbol_errors + b.symbol_errors == runs[0].symbol_errors
    assert max(a.tx_power_max, b.tx_power_max) == runs[0].tx_power_max
    link.close()


def test_two_devices_in_one_process(kat):
    """The dynamic-shared-memory opt-in and the occupancy are per device: links on device 0 and device 1 of ONE process run
    the same shapes (N = 1024 needs 164 KB of shared memory per block) and return identical counters."""
    from ofdm_based_systems import _native
    if _native.device_count() < 2:
        pytest.skip("needs two GPUs")
    taps = kat["chan_severe_multipath"]
    taps_n = taps / np.sqrt(np.sum(np.abs(taps) ** 2))
    out = []
    for dev in (0, 1, 0):
        for n, order in ((1024, 64), (256, 16), (8192, 16)):
            link = _native.Link(n, taps_n, np.fft.fft(taps, n), np.full(n, order), prefix_type="CYCLIC", prefix_len=7,
                                equalizer="MMSE", device=dev)
            r = link.run_fused(18.0, 0.089, 300, seed=9)
            link.close()
            out.append((dev, n, r.bit_errors, r.bits))
    per_shape = {}
    for dev, n, be, bits in out:
        per_shape.setdefault(n, set()).add((be, bits))
    assert all(len(v) == 1 for v in per_shape.values()), out
