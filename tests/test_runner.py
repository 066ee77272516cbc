"""Batch-driver facade (reference main.py): CSV upsert and the file-naming contract on the CPU, the whole
``SimulationRunner`` flow with the shipped JSON configs on the GPU."""
import os

import numpy as np
import pandas as pd
import pytest
from PIL import Image

from conftest import ROOT


def test_results_manager_contract(tmp_path):
    from ofdm_based_systems.main import ResultsManager
    m = ResultsManager(results_dir=str(tmp_path / "results"), images_dir=str(tmp_path / "images"),
                       channel_name="severe_multipath", doc_figures_dir=str(tmp_path / "docs"))
    m.update_ber_csv("CP-OFDM-ZF", 20.0, 1e-3)
    m.update_ber_csv("CP-OFDM-ZF", 30.0, 1e-5)
    m.update_ber_csv("CP-OFDM-ZF", 20.0, 2e-3)              # upsert, not append
    df = pd.read_csv(m.csv_path)
    assert len(df) == 2 and float(df[df.snr_db == 20.0].bit_error_rate.iloc[0]) == 2e-3
    p = m.save_constellation_plot(Image.new("RGB", (4, 4)), "CP", "OFDM", "ZF", 64, "QAM", "WF", 30.0)
    assert p.name == "CP-OFDM-ZF-64QAM-WF-SNR30_0dB.png" and p.exists()
    assert (tmp_path / "docs" / "severe_multipath" / p.name).exists()
    res = [dict(snr_db=s, bit_error_rate=b, prefix_acronym="CP", modulator_type="OFDM", equalizator_type="MMSE",
                constellation_order=16, constellation_scheme="QAM", power_allocation_acronym="UNIFORM")
           for s, b in ((0, 0.2), (10, 0.01), (20, 0.0))]
    q = m.plot_ber_vs_snr(res)
    assert q.name == "CP-OFDM-MMSE-16QAM-UNIFORM-BER_vs_SNR.png" and q.exists()


def test_shipped_configs_load_unchanged():
    from ofdm_based_systems.configuration.models import Settings, SimulationSettings
    cfg = os.path.join(ROOT, "config")
    assert Settings.from_json(os.path.join(cfg, "settings.json")).project_name
    for name in sorted(os.listdir(cfg)):
        if name.startswith("simulation_settings"):
            s = SimulationSettings.from_json(os.path.join(cfg, name))
            assert s.num_bands >= 8 and len(s.signal_noise_ratios) >= 1
    with pytest.raises(FileNotFoundError):
        Settings.from_json(os.path.join(cfg, "missing.json"))


@pytest.mark.gpu
@pytest.mark.parametrize("config", ["simulation_settings_test.json", "simulation_settings_adaptive.json",
                                    "simulation_settings.json"])
def test_runner_end_to_end(tmp_path, monkeypatch, config):
    from ofdm_based_systems.configuration.models import Settings, SimulationSettings
    from ofdm_based_systems.main import ResultsManager, SimulationRunner
    monkeypatch.chdir(ROOT)                                   # channel_model_path is CWD-relative
    sim_settings = SimulationSettings.from_json(os.path.join("config", config))
    runner = SimulationRunner(Settings.from_json("config/settings.json"), sim_settings,
                              ResultsManager(results_dir=str(tmp_path / "r"), images_dir=str(tmp_path / "i"),
                                             channel_name="t", doc_figures_dir=None))
    results = runner.run_all()
    runner.process_results(results)
    assert len(results) == len(sim_settings.signal_noise_ratios)
    bers = [r["bit_error_rate"] for r in results]
    if config.endswith("test.json"):                          # 5.4e-4 @20 dB, 0 @30 dB on 409 600 bits
        assert 2e-4 < bers[0] < 1.2e-3 and bers[1] < 3e-5
        assert results[0]["total_bits"] == 409600
    elif config == "simulation_settings.json":                # the reference's default: SC-OFDM, QPSK, ZF, 7 SNR points
        assert results[0]["modulator_type"] == "SC_OFDM" and results[0]["total_bits"] == 2_000_000
        assert all(a > b for a, b in zip(bers[:5], bers[1:5])) and 0.05 < bers[0] < 0.3 and bers[-1] < 1e-5
    else:                                                     # adaptive: orders per SNR as measured on the reference
        assert all(b < 5e-3 for b in bers)
        assert max(results[2]["constellation_order_per_subcarrier"]) == 64
    assert len(pd.read_csv(tmp_path / "r" / "ber_results.csv")) == len(results)
    assert len(list((tmp_path / "i" / "t").glob("*.png"))) == len(results) + 1
