"""GPU parity tests proper: the CUDA link kernel, called through the C ABI, replays the exact bits,
noise and taps that were fed to the LIVE reference (tests/golden/link_*.npz) and must reproduce

* the transmitted labels and the decided labels bit-exactly away from decision boundaries,
* the reference's bit / symbol error counts,
* Y (FFT output) and Z (equaliser output) within 1e-5 relative error (fp32 kernel vs fp64 reference),
* PAPR.

The oracle (oracle/ofdm_oracle.py) supplies per-symbol labels that the fixtures do not store."""
import numpy as np
import pytest

import ofdm_oracle as oc
from conftest import golden_link_names, golden_loaded_names, golden_noisebump_names, load_golden

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5          # north_star: FFT / equaliser outputs within 1e-5 relative (fp32 vs fp64)
BOUNDARY_TAU = 2e-4     # "away from a decision boundary": farther than this (constellation units)


def make_link(g):
    from ofdm_based_systems._native import Link
    n = int(g["n_sc"])
    orders = g["orders"] if g["orders"].size else np.full(n, int(g["order"]), dtype=np.int64)
    return Link(n, g["taps_chan"], g["H_eq"], orders, prefix_type=str(g["prefix_type"]), prefix_len=int(g["prefix_len"]),
                modulator=str(g["modulator"]), equalizer=str(g["eq"]), scheme=str(g["scheme"]))


def oracle_run(g):
    orders = g["orders"] if g["orders"].size else None
    setup = oc.LinkSetup(n_sc=int(g["n_sc"]), taps_raw=g["taps_raw"], snr_db=float(g["snr_db"]), order=int(g["order"]),
                         scheme=str(g["scheme"]), modulator=str(g["modulator"]), prefix_type=str(g["prefix_type"]),
                         eq=str(g["eq"]), awgn=bool(g["awgn"]), orders=orders, prefix_len_override=int(g["prefix_len"]))
    return setup, oc.run_link(setup, g["tx_bytes"].tobytes(), int(g["total_bits"]),
                              noise=(g["noise"] if bool(g["awgn"]) else None))


def rel_err(a, ref):
    return float(np.max(np.abs(a - ref)) / np.max(np.abs(ref)))


# golden cases whose link shape the register-resident fast kernel covers (csrc/link_fast.cuh)
FAST_CASES = {"headline_n1024_64qam_mmse", "c2_n1024_16qam_mmse", "c3_n64_64qam_mmse_p2", "c3_n64_64qam_zf_p2",
              "c5_n4096_256qam_mmse", "c1_n64_qpsk_cp16_awgn_zf", "cp10_n64_16qam_mmse", "rawtaps_n64_16qam_mmse", "n512_256qam_zf", "sc_n256_16qam_mmse",
              "sc_n64_qpsk_zf_p1", "zp_n256_64qam_zf", "zp_n64_16qam_mmse", "isi_zp2_n64_qpsk_mmse", "isi_cp3_n256_64qam_mmse",
              "isi_none_n128_16qam_zf", "psk2_n64_zf_flat", "psk8_n128_mmse_two_ray", "psk16_n64_none_eq_flat", "adaptive_p1_n64_mmse",
              "adaptive_severe_n256_mmse"}


@pytest.mark.parametrize("kernel", ["auto", "general"])
@pytest.mark.parametrize("noise_dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("name", golden_link_names())
def test_replay_matches_reference(name, noise_dtype, kernel, monkeypatch):
    g = load_golden("link", name)
    if kernel == "general":
        monkeypatch.setenv("OFDM_B200_FORCE_GENERAL", "1")     # read by ofdm_link_create
    n, n_ofdm = int(g["n_sc"]), int(g["n_ofdm"])
    setup, ref = oracle_run(g)
    assert ref["bit_errors"] == int(g["bit_errors"])          # the oracle itself is pinned to the reference
    link = make_link(g)
    if kernel == "general":
        assert not link.uses_fast_kernel
    elif name in FAST_CASES:
        assert link.uses_fast_kernel, "this link shape must run on the fast kernel"
    elif not link.uses_fast_kernel:
        pytest.skip("general kernel only: covered by kernel=general")
    noise = g["noise"].astype(noise_dtype) if bool(g["awgn"]) else None
    res, d = link.run_replay(float(g["snr_db"]), g["tx_bytes"].tobytes(), noise, n_ofdm,
                             compare_limit_bits=8 * g["tx_bytes"].size, dump=("y", "z", "rx_labels", "tx_labels"))
    link.close()

    active = (np.asarray(ref["tx_labels"]).reshape(n_ofdm, n) >= 0)
    tx_ref = np.where(active, np.asarray(ref["tx_labels"]).reshape(n_ofdm, n), 0)
    rx_ref = np.where(active, np.asarray(ref["rx_labels"]).reshape(n_ofdm, n), 0)
    np.testing.assert_array_equal(d["tx_labels"], tx_ref)

    # Y and Z: fp32 vs fp64
    assert rel_err(d["y"].astype(np.complex128), ref["Y"]) < REL_TOL
    z_ref = g["received_symbols"].reshape(n_ofdm, n)
    finite = np.isfinite(z_ref)
    assert rel_err(np.where(finite, d["z"].astype(np.complex128), 0), np.where(finite, z_ref, 0)) < REL_TOL

    # decisions: bit-exact away from decision boundaries
    if setup.adaptive:
        dist = np.full(z_ref.shape, np.inf)
        for k, o in enumerate(setup.orders):
            if o > 1:
                f = oc.qam_boundary_distance if setup.scheme == oc.QAM else oc.psk_boundary_distance
                dist[:, k] = f(z_ref[:, k], int(o))
    else:
        f = oc.qam_boundary_distance if setup.scheme == oc.QAM else oc.psk_boundary_distance
        dist = f(z_ref, setup.order)
    mismatch = (d["rx_labels"] != rx_ref) & active
    assert not np.any(mismatch & (dist > BOUNDARY_TAU)), "decision differs away from a boundary"
    n_boundary = int(np.sum(mismatch))
    if n_boundary == 0:
        assert res.bit_errors == int(g["bit_errors"])
        assert res.symbol_errors == int(g["symbol_errors"])
    else:  # a sample sat within fp32 rounding of a threshold: report, and bound the effect
        assert n_boundary <= 2
        assert abs(res.bit_errors - int(g["bit_errors"])) <= 16 * n_boundary
    assert res.ofdm_symbols == n_ofdm
    assert res.symbols == n_ofdm * n
    assert res.bits == min(8 * g["tx_bytes"].size, n_ofdm * link.bits_per_ofdm_symbol)
    assert abs(res.papr_db - float(g["papr_db"])) < 2e-4


@pytest.mark.parametrize("kernel", ["auto", "general"])
@pytest.mark.parametrize("n,n_ofdm", [(64, 48), (1024, 8), (4096, 8)])
def test_adaptive_replay_matches_oracle(n, n_ofdm, kernel, monkeypatch, kat):
    """Per-subcarrier orders 0 / 4 / 16 / 64 / 256 with recorded bits and noise: symbols start at arbitrary bit offsets of the
    byte stream (constellation/adaptive.py:178-198).  The oracle is the checker (pinned to the reference by the fixtures)."""
    from ofdm_based_systems._native import Link
    if kernel == "general":
        monkeypatch.setenv("OFDM_B200_FORCE_GENERAL", "1")
    rng = np.random.default_rng(n + 1)
    orders = rng.choice([0, 4, 16, 64, 256], size=n, p=[0.1, 0.3, 0.25, 0.2, 0.15]).astype(np.int64)
    orders[:4] = [256, 0, 4, 16]
    taps_raw, snr, P = kat["chan_severe_multipath"], 27.0, 7
    setup = oc.LinkSetup(n_sc=n, taps_raw=taps_raw, snr_db=snr, order=16, eq="MMSE", orders=orders, prefix_len_override=P)
    bits_per = int(sum(oc.bits_per_symbol(int(o)) for o in orders if o > 1))
    assert (bits_per * n_ofdm) % 8 == 0 and bits_per % 32 != 0          # whole bytes in total, unaligned symbols
    tx = rng.integers(0, 256, bits_per * n_ofdm // 8, dtype=np.uint8)
    sigma = np.sqrt(np.mean(orders > 1) / 10 ** (snr / 10) / 2)
    noise = (rng.normal(size=n_ofdm * (n + P)) + 1j * rng.normal(size=n_ofdm * (n + P))) * sigma
    ref = oc.run_link(setup, tx.tobytes(), bits_per * n_ofdm, noise=noise)
    link = Link(n, setup.taps_chan, setup.H_eq, orders, prefix_type="CYCLIC", prefix_len=P, equalizer="MMSE")
    assert link.uses_fast_kernel == (kernel == "auto")
    res, d = link.run_replay(snr, tx.tobytes(), noise, n_ofdm, dump=("z", "rx_labels", "tx_labels"))
    link.close()
    act = orders > 1
    tx_ref = np.where(act, np.asarray(ref["tx_labels"]).reshape(n_ofdm, n), 0)
    rx_ref = np.where(act, np.asarray(ref["rx_labels"]).reshape(n_ofdm, n), 0)
    np.testing.assert_array_equal(d["tx_labels"], tx_ref)
    z_ref = np.asarray(ref["received_symbols"]).reshape(n_ofdm, n)
    assert rel_err(d["z"][:, act].astype(np.complex128), z_ref[:, act]) < REL_TOL
    dist = np.full(z_ref.shape, np.inf)
    for k in np.nonzero(act)[0]:
        dist[:, k] = oc.qam_boundary_distance(z_ref[:, k], int(orders[k]))
    mismatch = (d["rx_labels"] != rx_ref) & act
    assert not np.any(mismatch & (dist > BOUNDARY_TAU))
    if not mismatch.any():
        assert res.bit_errors == ref["bit_errors"] and res.symbol_errors == ref["symbol_errors"]
    assert res.bits == bits_per * n_ofdm


@pytest.mark.parametrize("kernel", ["auto", "general"])
@pytest.mark.parametrize("name", golden_loaded_names())
def test_applied_power_loading_replay_matches_reference(name, kernel, monkeypatch):
    """Applied power loading (tx amplitudes sqrt(P_k), receiver gains 1/sqrt(P_k)) replayed from the live reference's
    recorded bits and noise (tests/golden/loaded_*.npz), on both kernels."""
    from ofdm_based_systems._native import Link
    if kernel == "general":
        monkeypatch.setenv("OFDM_B200_FORCE_GENERAL", "1")
    g = load_golden("loaded", name)
    n, n_ofdm, order = int(g["n_sc"]), int(g["n_ofdm"]), int(g["order"])
    link = Link(n, g["taps_chan"], g["H_eq"], np.full(n, order), prefix_type="CYCLIC", prefix_len=int(g["prefix_len"]),
                equalizer=str(g["eq"]), amp=g["amp"], rx_gain=g["rx_gain"])
    assert link.uses_fast_kernel == (kernel == "auto")
    res, d = link.run_replay(float(g["snr_db"]), g["tx_bytes"].tobytes(), g["noise"], n_ofdm, dump=("z", "rx_labels"))
    link.close()
    z_ref = g["received_symbols"].reshape(n_ofdm, n)
    assert rel_err(d["z"].astype(np.complex128), z_ref) < REL_TOL
    rx_ref = oc.labels_from_bits(g["rx_bytes"].tobytes(), oc.bits_per_symbol(order)).reshape(n_ofdm, n)
    mismatch = d["rx_labels"] != rx_ref
    assert not np.any(mismatch & (oc.qam_boundary_distance(z_ref, order) > BOUNDARY_TAU))
    if not mismatch.any():
        assert res.bit_errors == int(g["bit_errors"]) and res.symbol_errors == int(g["symbol_errors"])
    assert abs(res.papr_db - float(g["papr_db"])) < 2e-4


@pytest.mark.parametrize("name", golden_noisebump_names())
def test_post_equaliser_stage_replay_matches_reference(name):
    """The post-equaliser stage of examples/waterfilling_noise_bump_experiment.py:163-183 (ofdm_link_set_post) replayed from
    the reference's recorded bits and noise matrix: first pass = mean power of the compensated values, second pass = the
    renormalised values through the demapper."""
    from ofdm_based_systems._native import Link
    g = load_golden("noisebump", name)
    n, n_ofdm, order = int(g["n_sc"]), int(g["n_ofdm"]), int(g["order"])
    link = Link(n, g["taps_chan"], g["H_eq"], np.full(n, order), prefix_type="CYCLIC", prefix_len=int(g["prefix_len"]),
                equalizer="MMSE", amp=g["amp"], rx_gain=g["rx_gain"])
    assert link.uses_fast_kernel
    link.set_post(g["noise_profile"], recorded_noise=g["post_noise"], measure_power=True)
    assert not link.uses_fast_kernel                      # the stage lives in the general kernel
    link.run_replay(float(g["snr_db"]), g["tx_bytes"].tobytes(), None, n_ofdm)
    total, count = link.read_z_power()
    assert count == n_ofdm * n
    avg = total / count
    assert abs(avg - float(g["avg_power"])) < 1e-5 * float(g["avg_power"])
    link.set_post(g["noise_profile"], recorded_noise=g["post_noise"], z_scale=1.0 / np.sqrt(avg))
    res, d = link.run_replay(float(g["snr_db"]), g["tx_bytes"].tobytes(), None, n_ofdm, dump=("z", "rx_labels"))
    link.set_post()
    assert link.uses_fast_kernel
    link.close()
    z_ref = g["received_symbols"].reshape(n_ofdm, n)
    # 1e-5 of the largest value on every subcarrier, scaled by how much more than the typical subcarrier the receiver gain
    # amplifies the fp32 equaliser output: the water-filling floor (1e-4 of the power) puts a gain of 100 on the
    # subcarrier in the channel's spectral dip, 12 x the gain of the others (measured: 1.7e-5 there, < 4e-7 elsewhere)
    err = np.abs(d["z"].astype(np.complex128) - z_ref).max(axis=0) / np.abs(z_ref).max()
    assert np.all(err < REL_TOL * np.maximum(1.0, g["rx_gain"] / np.median(g["rx_gain"])))
    rx_ref = oc.labels_from_bits(g["rx_bytes"].tobytes(), oc.bits_per_symbol(order)).reshape(n_ofdm, n)
    mismatch = d["rx_labels"] != rx_ref
    tau = BOUNDARY_TAU * np.maximum(1.0, g["rx_gain"] / np.median(g["rx_gain"]))[None, :]
    assert not np.any(mismatch & (oc.qam_boundary_distance(z_ref, order) > tau))
    if not mismatch.any():
        assert res.bit_errors == int(g["bit_errors"])
    assert abs(res.papr_db - float(g["papr_db"])) < 2e-4


def test_post_equaliser_stage_fused_statistics():
    """Fused mode of the same stage (Philox stream 3, two passes over the same seed): the mean power of the compensated
    values is what the noise profile and the loading predict, the renormalised block has unit power, and the BER agrees
    with the oracle run on its own random numbers."""
    from ofdm_based_systems._native import Link
    g = load_golden("noisebump", "wf_bump6_snr25")
    n, order, snr = int(g["n_sc"]), int(g["order"]), float(g["snr_db"])
    link = Link(n, g["taps_chan"], g["H_eq"], np.full(n, order), prefix_type="CYCLIC", prefix_len=int(g["prefix_len"]),
                equalizer="MMSE", amp=g["amp"], rx_gain=g["rx_gain"])
    n_ofdm = 20_000
    link.set_post(g["noise_profile"], measure_power=True)
    link.run_fused(snr, 0.0, n_ofdm, seed=5)
    total, count = link.read_z_power()
    avg = total / count
    link.set_post(g["noise_profile"], z_scale=1.0 / np.sqrt(avg), measure_power=True)
    res = link.run_fused(snr, 0.0, n_ofdm, seed=5)
    total2, _ = link.read_z_power()
    assert abs(total2 / count - avg) < 1e-6 * avg       # the power is measured before the scale, same streams
    res2 = link.run_fused_renormalised(snr, 0.0, n_ofdm, noise_profile=g["noise_profile"], seed=5)
    assert (res2.bit_errors, res2.bits) == (res.bit_errors, res.bits)
    link.close()
    # the oracle on NumPy random numbers: 400 OFDM symbols, the OFDM symbol as the unit of the standard error
    rng = np.random.default_rng(11)
    setup = oc.LinkSetup(n_sc=n, taps_raw=g["taps_raw"], snr_db=snr, order=order, eq="MMSE", awgn=False,
                         prefix_len_override=int(g["prefix_len"]), amp=g["amp"], rx_gain=g["rx_gain"])
    s_ref = 400
    bits = oc.generate_bits(s_ref * n * 6, rng)
    std = np.sqrt(10 ** (-snr / 10) * g["noise_profile"] / 2)[None, :]
    post = (rng.normal(size=(s_ref, n)) + 1j * rng.normal(size=(s_ref, n))) * std
    ref = oc.run_link(setup, bits, s_ref * n * 6, post_noise=post, renormalise=True)
    assert abs(avg - ref["z_avg_power"]) < 0.05 * ref["z_avg_power"]
    tb, rb = oc.unpack_bits(bits), oc.unpack_bits(ref["rx_bytes"])
    per_symbol = np.mean((tb != rb).reshape(s_ref, n * 6), axis=1)
    sem = per_symbol.std(ddof=1) / np.sqrt(s_ref)
    assert abs(res.bit_errors / res.bits - per_symbol.mean()) < 4 * sem + 1e-4
