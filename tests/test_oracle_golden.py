"""Pins the CPU oracle (oracle/ofdm_oracle.py) against fixtures produced by the LIVE reference
(oracle/make_golden.py) and against the known-answer vectors of the reference's own tests."""
import numpy as np
import pytest

import ofdm_oracle as oc
from conftest import golden_loaded_names, golden_link_names, golden_noisebump_names, load_golden


def _setup(g):
    orders = g["orders"] if g["orders"].size else None
    return oc.LinkSetup(n_sc=int(g["n_sc"]), taps_raw=g["taps_raw"], snr_db=float(g["snr_db"]),
                        order=int(g["order"]), scheme=str(g["scheme"]), modulator=str(g["modulator"]),
                        prefix_type=str(g["prefix_type"]), eq=str(g["eq"]), awgn=bool(g["awgn"]),
                        orders=orders, prefix_len_override=int(g["prefix_len"]))


@pytest.mark.parametrize("name", golden_link_names())
def test_link_replay_matches_reference(name):
    g = load_golden("link", name)
    setup = _setup(g)
    noise = g["noise"] if bool(g["awgn"]) else None
    r = oc.run_link(setup, g["tx_bytes"].tobytes(), int(g["total_bits"]), noise=noise)
    assert r["bit_errors"] == int(g["bit_errors"])
    assert r["symbol_errors"] == int(g["symbol_errors"])
    assert r["rx_bytes"] == g["rx_bytes"].tobytes()
    np.testing.assert_allclose(r["received_symbols"], g["received_symbols"], rtol=0, atol=1e-12)
    assert abs(r["papr_db"] - float(g["papr_db"])) < 1e-10
    np.testing.assert_allclose(setup.taps_chan, g["taps_chan"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(setup.H_eq, g["H_eq"], rtol=0, atol=1e-13)
    for key, mine in (("symbols", "symbols"), ("rx", "rx"), ("Y", "Y")):
        if key in g.files:
            np.testing.assert_allclose(np.asarray(r[mine]).reshape(-1), g[key].reshape(-1), rtol=0, atol=1e-12)
    if "tx" in g.files:
        np.testing.assert_allclose(r["tx"], g["tx"], rtol=0, atol=1e-13)


@pytest.mark.parametrize("name", [n for n in golden_link_names() if "normal_re" in load_golden("link", n).files])
def test_awgn_scaling_from_normals(name):
    g = load_golden("link", name)
    setup = _setup(g)
    r = oc.run_link(setup, g["tx_bytes"].tobytes(), int(g["total_bits"]), normals=(g["normal_re"], g["normal_im"]))
    np.testing.assert_allclose(r["noise"], g["noise"], rtol=1e-14, atol=0)
    assert r["bit_errors"] == int(g["bit_errors"])


def test_constellation_tables(kat):
    for m in (4, 16, 64, 256, 1024):
        np.testing.assert_array_equal(oc.qam_constellation(m), kat[f"qam{m}"])
        np.testing.assert_allclose(oc.qam_constellation_closed_form(m), kat[f"qam{m}"], rtol=0, atol=1e-15)
    for m in (2, 4, 8, 16, 32):
        np.testing.assert_array_equal(oc.psk_constellation(m), kat[f"psk{m}"])


def test_gray_tables(kat):
    # reference KAT: tests/ofdm_based_systems/constellation/test_models.py:78-100
    assert [int(oc.gray(i)) for i in range(8)] == [0, 1, 3, 2, 6, 7, 5, 4]
    for b in (2, 3, 4):
        g = np.array([oc.gray(i) for i in range(1 << b)])
        np.testing.assert_array_equal(g, kat[f"gray{b}"])
        np.testing.assert_array_equal(np.argsort(g), kat[f"igray{b}"])


def test_bit_loading_tables(kat):
    snrs = kat["bl_snr"]
    assert [oc.bit_loading_qam(1e-3, s) for s in snrs] == kat["bl_qam_1e-3"].tolist() == [0, 0, 4, 4, 16, 64, 256, 1024, 16384]
    assert [oc.bit_loading_psk(1e-3, s) for s in snrs] == kat["bl_psk_1e-3"].tolist() == [0, 2, 4, 4, 8, 16, 32, 128, 256]
    assert [oc.bit_loading_qam(1e-2, s) for s in snrs] == kat["bl_qam_1e-2"].tolist()
    assert [oc.bit_loading_psk(1e-5, s) for s in snrs] == kat["bl_psk_1e-5"].tolist()


def test_waterfilling_kats(kat):
    p1, mu, it = oc.waterfilling(5.0, kat["wf1_gains"], 0.1, return_info=True)
    np.testing.assert_array_equal(p1, kat["wf1_power"])
    np.testing.assert_allclose(p1, [1.0256666667, 1.0206666667, 1.0123333333, 0.9956666667, 0.9456666667], atol=1e-9)
    assert it == 31 and abs(mu - 1.0456666666) < 1e-8
    p2 = oc.waterfilling(1.0, kat["wf2_gains"], 0.1)
    np.testing.assert_array_equal(p2, kat["wf2_power"])
    np.testing.assert_allclose(p2, [0.5013888889, 0.4986111111, 0, 0], atol=1e-9)
    np.testing.assert_array_equal(oc.uniform_power(5.0, 4), kat["uniform_5_4"])


def test_waterfilling_and_orders_on_shipped_channels(kat):
    for nm in kat["channel_names"]:
        h = kat["chan_" + str(nm)]
        for n_sc in (64, 1024):
            for snr in (5.0, 20.0):
                key = f"wf_{nm}_{n_sc}_{int(snr)}"
                gains = np.abs(np.fft.fft(h, n_sc)) ** 2
                n0 = 10 ** (-snr / 10)
                pw = oc.waterfilling(float(n_sc), gains, n0)
                np.testing.assert_array_equal(pw, kat[key + "_power"])
                np.testing.assert_array_equal(oc.bit_loading_orders(pw, gains, n0, 1e-3), kat[key + "_orders"])
                np.testing.assert_array_equal(oc.capacity_per_subcarrier(pw, gains, n0), kat[key + "_cap"])


def test_shannon_orders(kat):
    np.testing.assert_array_equal(oc.shannon_orders(kat["shannon_cap"], 4, 256, 1.0, oc.QAM), kat["shannon_qam"])
    np.testing.assert_array_equal(oc.shannon_orders(kat["shannon_cap"], 4, 256, 0.85, oc.PSK), kat["shannon_psk"])


def test_bit_order_and_tail_mask():
    # MSB-first (reference tests/ofdm_based_systems/simulation/test_models.py:27-102)
    assert oc.unpack_bits(bytes([0b10110001])).tolist() == [1, 0, 1, 1, 0, 0, 0, 1]
    # tail-bit masking (reference tests/ofdm_based_systems/bits_generation/test_models.py:392-404)
    class Fake:
        def bytes(self, n):
            return b"\xff" * n
    assert oc.generate_bits(11, Fake()) == bytes([0xFF, 0b11100000])
    assert oc.generate_bits(16, Fake()) == b"\xff\xff"
    assert oc.pack_bits(np.array([1, 0, 1])) == bytes([0b10100000])


def test_zp_overlap_add_kat():
    # reference tests/ofdm_based_systems/prefix/test_models.py:297-311: [1..6], P=2 -> [1+5, 2+6, 3, 4]
    rows = np.arange(1, 7, dtype=np.complex128)[None, :]
    np.testing.assert_array_equal(oc.remove_prefix(rows, 2, oc.PREFIX_ZERO)[0], [6, 8, 3, 4])
    x = np.arange(1, 5, dtype=np.complex128)[None, :]
    np.testing.assert_array_equal(oc.add_prefix(x, 2, oc.PREFIX_CYCLIC)[0], [3, 4, 1, 2, 3, 4])
    np.testing.assert_array_equal(oc.add_prefix(x, 2, oc.PREFIX_ZERO)[0], [1, 2, 3, 4, 0, 0])


def test_zf_exact_division_and_mmse_formula():
    # reference tests/ofdm_based_systems/equalization/test_models.py:67-88
    H = np.array([2 + 0j, 0 + 1j, 0.5 + 0.5j, 0j])
    Y = np.array([[4 + 2j, 1 + 1j, 1 + 0j, 1e-10 + 0j]])
    Z = oc.equalize_rows(Y, H, oc.EQ_ZF, None)
    np.testing.assert_allclose(Z[0, :3], Y[0, :3] / H[:3], rtol=1e-15)
    assert Z[0, 3] == 1.0
    sig = np.mean(np.abs(Y) ** 2)
    s2 = sig / 10 ** (1.0) / np.mean(np.abs(H) ** 2)
    np.testing.assert_allclose(oc.equalize_rows(Y, H, oc.EQ_MMSE, 10.0)[0], Y[0] * np.conj(H) / (np.abs(H) ** 2 + s2), rtol=1e-15)


def test_adaptive_setup_matches_simulation_run():
    g = load_golden("sim", "adaptive_wf_mmse")
    orders, power, level = oc.adaptive_setup(64, load_golden("link", "adaptive_p1_n64_mmse")["taps_raw"], 20.0, 1e-3)
    np.testing.assert_array_equal(orders, g["constellation_order_per_subcarrier"])
    np.testing.assert_array_equal(power, g["allocated_power"])
    assert level == float(g["water_level"])


from conftest import golden_sim_names  # noqa: E402


@pytest.mark.parametrize("name", golden_sim_names())
def test_simulation_run_matches_reference(name):
    """Seeds the oracle's explicit RNGs the way make_golden.py seeded the reference's implicit ones
    and compares the whole Simulation.run() result."""
    g = load_golden("sim", name)
    kw = {k[4:]: g[k].item() if g[k].ndim == 0 else g[k] for k in g.files if k.startswith("arg_")}
    seed = int(g["seed"])
    r = oc.simulate(**kw, bit_rng=np.random.Generator(np.random.PCG64(seed)), noise_rng=np.random.RandomState(seed))
    assert r["total_bits"] == int(g["total_bits"])
    assert r["bit_errors"] == int(g["bit_errors"])
    assert r["symbol_errors"] == int(g["symbol_errors"])
    assert abs(r["papr_db"] - float(g["papr_db"])) < 1e-9
    np.testing.assert_allclose(r["received_symbols"], g["received_symbols"], rtol=0, atol=1e-12)
    np.testing.assert_array_equal(r["constellation_order_per_subcarrier"], g["constellation_order_per_subcarrier"])
    np.testing.assert_array_equal(r["allocated_power"], g["allocated_power"])
    if np.isnan(g["water_level"]):
        # quirk Q10: FIXED + WATERFILLING reports None because the dict is filled before the value exists
        assert r["water_level"] is None or str(g["arg_adaptive_modulation_mode"]) == "FIXED"
    else:
        assert r["water_level"] == float(g["water_level"])


@pytest.mark.parametrize("name", golden_loaded_names())
def test_applied_power_loading_matches_reference_components(name):
    """the oracle's amp / rx_gain path against the live reference's components with the experiment's
    loading lines around them (oracle/make_golden.py::loaded_cases)."""
    g = load_golden("loaded", name)
    setup = oc.LinkSetup(n_sc=int(g["n_sc"]), taps_raw=g["taps_raw"], snr_db=float(g["snr_db"]), order=int(g["order"]),
                         eq=str(g["eq"]), prefix_len_override=int(g["prefix_len"]), amp=g["amp"], rx_gain=g["rx_gain"])
    r = oc.run_link(setup, g["tx_bytes"].tobytes(), int(g["total_bits"]), noise=g["noise"])
    assert r["bit_errors"] == int(g["bit_errors"]) and r["symbol_errors"] == int(g["symbol_errors"])
    assert r["rx_bytes"] == g["rx_bytes"].tobytes()
    np.testing.assert_allclose(r["received_symbols"], g["received_symbols"], rtol=0, atol=1e-12)
    assert abs(r["papr_db"] - float(g["papr_db"])) < 1e-10


@pytest.mark.parametrize("name", golden_noisebump_names())
def test_post_equaliser_stage_matches_reference_experiment(name):
    """coloured noise injected after the equaliser, receiver compensation and the block-wide renormalisation of
    examples/waterfilling_noise_bump_experiment.py:163-183, recorded from the reference's own components
    (oracle/make_golden.py::noise_bump_cases)."""
    g = load_golden("noisebump", name)
    setup = oc.LinkSetup(n_sc=int(g["n_sc"]), taps_raw=g["taps_raw"], snr_db=float(g["snr_db"]), order=int(g["order"]),
                         eq="MMSE", awgn=False, prefix_len_override=int(g["prefix_len"]), amp=g["amp"], rx_gain=g["rx_gain"])
    r = oc.run_link(setup, g["tx_bytes"].tobytes(), int(g["total_bits"]), post_noise=g["post_noise"], renormalise=True)
    assert r["bit_errors"] == int(g["bit_errors"])
    assert r["rx_bytes"] == g["rx_bytes"].tobytes()
    np.testing.assert_allclose(r["received_symbols"], g["received_symbols"], rtol=0, atol=1e-11)
    assert abs(r["z_avg_power"] - float(g["avg_power"])) < 1e-10 * float(g["avg_power"])
