"""Frame batches (ofdm_frames_run): many channel realisations in one launch.  Every frame must reproduce the
single-link path (ofdm_link_run_fused of a link built from the same taps and orders, same Philox counters), which
the other GPU tests pin to the reference through the oracle."""
import numpy as np
import pytest

import ofdm_oracle as oc

pytestmark = pytest.mark.gpu


def rayleigh(rng, f, l):
    """examples/generate_channel_models.py:70-78"""
    h = (rng.normal(size=(f, l)) + 1j * rng.normal(size=(f, l))) / np.sqrt(2) * np.sqrt(np.exp(-np.arange(l) / 2))
    return h / np.sqrt(np.sum(np.abs(h) ** 2, axis=1, keepdims=True))


def table_link(n, taps_raw, orders, P, eq):
    """The single link a frame corresponds to, built on the HOST by ofdm_link_create in the formulation frame batches use
    (per-subcarrier level / mask tables; the unit amplitudes select it even when every subcarrier has the same order)."""
    from ofdm_based_systems._native import Link
    from ofdm_based_systems.simulation.sweep import LinkConfig
    cfg = LinkConfig(num_subcarriers=n, taps_raw=taps_raw, prefix_length=P, equalizator_type=eq, orders=orders)
    return cfg, Link(n, cfg.taps_chan, cfg.h_eq, orders, prefix_type="CYCLIC", prefix_len=P, equalizer=eq, amp=np.ones(n))


def single_link(n, taps_raw, orders, P, eq, snr, S, seed, point, first_symbol):
    cfg, link = table_link(n, taps_raw, orders, P, eq)
    res = link.run_fused(snr, cfg.noise_sigma(snr), S, seed=seed, point=point, first_symbol=first_symbol)
    link.close()
    return res


def ulps(a, b):
    """Distance in units of the last place between two float32 arrays (0 = bit-identical)."""
    ia = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    ib = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    ia, ib = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia), np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return np.abs(ia - ib)


@pytest.mark.parametrize("n,order,eq", [(64, 16, "ZF"), (256, 64, "MMSE"), (1024, None, "MMSE"), (4096, 256, "NONE")])
def test_device_built_frame_tables_equal_the_host_built_link_tables(n, order, eq):
    """frame_tables_kernel (fp64 on the device, per frame) against ofdm_link_create (fp64 on the host) for the same raw
    taps and orders: the fp32 tables the link kernel reads - equaliser / decision table, level table, packed field masks,
    taps in both forms, sigma and the MMSE constant - agree to the last place (at most one ulp where the fp64
    intermediate sits on a rounding boundary), which is what makes a frame reproduce a single link exactly."""
    from ofdm_based_systems import _native
    rng = np.random.default_rng(n + 1)
    F, L, snr = 6, 8, 19.0
    taps = rayleigh(rng, F, L) * rng.uniform(0.5, 2.0, size=(F, 1))        # raw taps need not be unit energy
    if order is None:
        got = _native.frames_debug_tables(n, taps, snr, equalizer=eq, waterfilling=True, min_order=4, max_order=256)
        used = _native.run_frames(n, F, 1, snr, taps=taps, equalizer=eq, waterfilling=True, min_order=4, max_order=256)["orders"]
    else:
        got = _native.frames_debug_tables(n, taps, snr, equalizer=eq, order=order)
        used = np.full((F, n), order)
    worst = 0
    for f in range(F):
        cfg, link = table_link(n, taps[f], used[f], L - 1, eq)
        ref = link.debug_tables()
        link.close()
        np.testing.assert_array_equal(got["masks"][f], ref["masks"])
        np.testing.assert_array_equal(got["level"][f][:, 1], ref["level"][:, 1])
        active = used[f] > 1
        for name, a, b in (("eq", got["eq"][f][active], ref["eq"][active]), ("level", got["level"][f][active, 0], ref["level"][active, 0]),
                           ("taps", got["taps"][f], ref["taps"]), ("taps3", got["taps3"][f], ref["taps3"])):
            d = int(ulps(a, b).max())
            assert d <= 1, f"frame {f} {name}: {d} ulp"
            worst = max(worst, d)
        assert ulps(got["sigma"][f], np.float32(cfg.noise_sigma(snr))) <= 1
        if eq == "MMSE":
            mmse_ref = np.float32(1.0 / (float(n) ** 2 * 10 ** (snr / 10) * np.mean(np.abs(cfg.h_eq) ** 2)))
            assert ulps(got["mmse_c"][f], mmse_ref) <= 1
    print(f"n={n}: worst table difference {worst} ulp")


@pytest.mark.parametrize("n,order,eq,S", [(64, 16, "ZF", 100), (128, 4, "MMSE", 70), (256, 64, "MMSE", 37), (512, 16, "NONE", 33), (1024, 64, "MMSE", 20),
                                          (4096, 256, "MMSE", 9), (8192, 64, "MMSE", 5)])
def test_fixed_order_frames_reproduce_single_links(n, order, eq, S):
    from ofdm_based_systems._native import run_frames
    rng = np.random.default_rng(n)
    F, L, snr = 12, 8, 22.0
    taps = rayleigh(rng, F, L)
    out = run_frames(n, F, S, snr, taps=taps, equalizer=eq, order=order, seed=99, point=1, first_frame=5)
    bps = oc.bits_per_symbol(order)
    assert out["total"].bits == F * S * n * bps and out["total"].ofdm_symbols == F * S
    assert np.all(out["orders"] == order)
    tot_err = 0
    for f in range(F):
        ref = single_link(n, taps[f], np.full(n, order), L - 1, eq, snr, S, 99, 1, (5 + f) * S)
        got = out["frames"][f]
        assert got.bits == ref.bits and got.symbols == ref.symbols and got.ofdm_symbols == S
        # same Philox draws, same kernel formulation, tables equal to the last place (test above): identical decisions
        assert (got.bit_errors, got.symbol_errors) == (ref.bit_errors, ref.symbol_errors)
        assert abs(got.tx_power_sum - ref.tx_power_sum) < 1e-6 * ref.tx_power_sum
        assert got.tx_power_max == ref.tx_power_max
        tot_err += got.bit_errors
    assert out["total"].bit_errors == tot_err and tot_err > 0


@pytest.mark.parametrize("n,wf", [(64, True), (256, False), (1024, True)])
def test_adaptive_rayleigh_frames(n, wf):
    """config #4: fresh Rayleigh realisation per frame drawn on the device, water-filling + gap-rule orders bounded
    to QPSK .. 256-QAM, one launch."""
    from ofdm_based_systems._native import run_frames
    F, S, snr = 40, 32, 18.0
    out = run_frames(n, F, S, snr, n_taps=8, equalizer="MMSE", waterfilling=wf, min_order=4, max_order=256, seed=7)
    taps = out["taps"]
    np.testing.assert_allclose(np.sum(np.abs(taps) ** 2, axis=1), 1.0, rtol=1e-12)
    # exponential power delay profile of the draw (before the per-frame normalisation it is exp(-l/2))
    big = run_frames(64, 4000, 1, snr, n_taps=8, order=4, seed=3)["taps"]
    prof = np.mean(np.abs(big) ** 2, axis=0)
    assert np.all(np.abs(prof / prof[0] - np.exp(-np.arange(8) / 2)) < 0.12)
    mism = 0
    for f in range(0, F, 3):
        orders, _, _ = oc.adaptive_setup(n, taps[f], snr, 1e-3, oc.QAM, waterfill=wf)
        bounded = np.where(orders > 256, 256, np.where(orders < 4, 0, orders))
        mism += int(np.sum(out["orders"][f] != bounded))
        ref = single_link(n, taps[f], out["orders"][f], 7, "MMSE", snr, S, 7, 0, f * S)
        got = out["frames"][f]
        assert got.bits == ref.bits == S * sum(oc.bits_per_symbol(int(o)) for o in out["orders"][f] if o > 1)
        assert (got.bit_errors, got.symbol_errors) == (ref.bit_errors, ref.symbol_errors)
    assert mism <= 2


def test_frames_shard_like_symbols():
    from ofdm_based_systems._native import run_frames
    kw = dict(n_taps=6, equalizer="MMSE", order=16, seed=1234, point=2)
    whole = run_frames(256, 30, 50, 15.0, **kw)
    a = run_frames(256, 11, 50, 15.0, first_frame=0, **kw)
    b = run_frames(256, 19, 50, 15.0, first_frame=11, **kw)
    np.testing.assert_array_equal(whole["taps"], np.concatenate([a["taps"], b["taps"]]))
    assert whole["total"].bit_errors == a["total"].bit_errors + b["total"].bit_errors > 0
    assert whole["total"].bits == a["total"].bits + b["total"].bits
    assert [r.bit_errors for r in whole["frames"]] == [r.bit_errors for r in a["frames"] + b["frames"]]
    # many symbols per frame, few frames: the frame is split into chunks over the SMs, same counters
    few = run_frames(1024, 3, 4000, 20.0, n_taps=8, order=64, seed=5)
    again = run_frames(1024, 1, 4000, 20.0, n_taps=8, order=64, seed=5, first_frame=1)
    assert few["frames"][1].bit_errors == again["frames"][0].bit_errors > 0
    with pytest.raises(ValueError):
        run_frames(96, 2, 10, 20.0, order=16)
    with pytest.raises(ValueError):
        run_frames(256, 2, 10, 20.0, max_order=1024)


def test_frame_sweep_single_rank_matches_direct_calls():
    from ofdm_based_systems._native import run_frames
    from ofdm_based_systems.simulation.sweep import FrameSweep
    res = FrameSweep(256, n_taps=8, equalizer="MMSE", waterfilling=True, min_order=4, max_order=256).sweep(
        [10.0, 20.0], 24, 40, seed=21)
    for i, snr in enumerate((10.0, 20.0)):
        d = run_frames(256, 24, 40, snr, n_taps=8, equalizer="MMSE", waterfilling=True, min_order=4, max_order=256,
                       seed=21, point=i)["total"]
        assert res[i]["bit_errors"] == d.bit_errors and res[i]["total_bits"] == d.bits
        assert abs(res[i]["papr_db"] - d.papr_db) < 1e-9
    assert res[0]["bit_error_rate"] > 0 and res[1]["total_bits"] > res[0]["total_bits"]   # more bits loaded at 20 dB
