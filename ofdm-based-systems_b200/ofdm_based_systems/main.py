"""Batch driver: ``python -m ofdm_based_systems.main`` (reference: main.py:19-393).

Same classes, file-naming contract and CSV upsert as the reference; the plots are drawn with Pillow
(``simulation/plotting.py``) because matplotlib is not part of this image.  ``run_all`` hands the list of
Simulations to ``Simulation.run_sweep``: the SNR values that share a link run as ONE CUDA launch, sharded over
the ranks of an initialised ``torch.distributed`` process group with one all-reduce per sweep."""
from __future__ import annotations

import shutil
from pathlib import Path
from typing import Any, Dict, List, Optional, Union

import numpy as np
import pandas as pd
from PIL import Image, ImageDraw

from ofdm_based_systems.configuration.models import Settings, SimulationSettings
from ofdm_based_systems.simulation.models import Simulation


def _config_stem(result: Dict[str, Any]) -> str:
    """'CP-OFDM-ZF-64QAM-WF' (main.py:129-139)."""
    return (f"{result.get('prefix_acronym', 'NONE')}-{result.get('modulator_type', 'OFDM')}-"
            f"{result.get('equalizator_type', 'NONE')}-{result.get('constellation_order', 16)}"
            f"{result.get('constellation_scheme', 'QAM')}-{result.get('power_allocation_acronym', 'UNIFORM')}")


class ResultsManager:
    """CSV storage (upsert on (simulation_name, snr_db)) and image files under images/<channel>/."""

    def __init__(self, results_dir: str = "results", images_dir: str = "images", channel_name: str = "default",
                 doc_figures_dir: Union[str, Path, None] = "docs/figures"):
        self.results_dir = Path(results_dir)
        self.channel_name = channel_name
        self.images_dir = Path(images_dir) / channel_name
        self.csv_path = self.results_dir / "ber_results.csv"
        self.doc_figures_dir: Optional[Path] = Path(doc_figures_dir) if doc_figures_dir else None
        self.doc_channel_dir: Optional[Path] = None
        self.results_dir.mkdir(parents=True, exist_ok=True)
        self.images_dir.mkdir(parents=True, exist_ok=True)
        if self.doc_figures_dir:
            self.doc_figures_dir.mkdir(parents=True, exist_ok=True)
            self.doc_channel_dir = self.doc_figures_dir / self.channel_name
            self.doc_channel_dir.mkdir(parents=True, exist_ok=True)

    def _mirror_to_docs(self, source_path: Path) -> Optional[Path]:
        if not self.doc_channel_dir or not source_path.exists():
            return None
        try:
            relative = source_path.relative_to(self.images_dir)
        except ValueError:
            relative = Path(source_path.name)
        destination = self.doc_channel_dir / relative
        destination.parent.mkdir(parents=True, exist_ok=True)
        shutil.copy2(source_path, destination)
        return destination

    def update_ber_csv(self, simulation_name: str, snr_db: float, bit_error_rate: float) -> None:
        columns = ["simulation_name", "snr_db", "bit_error_rate"]
        frame = pd.read_csv(self.csv_path) if self.csv_path.exists() else pd.DataFrame(columns=columns)
        hit = (frame["simulation_name"] == simulation_name) & (frame["snr_db"] == snr_db)
        if hit.any():
            frame.loc[hit, "bit_error_rate"] = bit_error_rate
        else:
            row = pd.DataFrame([{"simulation_name": simulation_name, "snr_db": snr_db, "bit_error_rate": bit_error_rate}])
            frame = row if frame.empty else pd.concat([frame, row], ignore_index=True)
        frame.to_csv(self.csv_path, index=False)

    def save_constellation_plot(self, image: Image.Image, prefix_type: str, modulation_type: str,
                                equalization_method: str, constellation_order: int, constellation_type: str,
                                power_allocation: str, snr_db: float) -> Path:
        """e.g. 'CP-OFDM-ZF-64QAM-WF-SNR30_0dB.png'."""
        snr = f"{snr_db:.1f}".replace(".", "_")
        path = self.images_dir / (f"{prefix_type}-{modulation_type}-{equalization_method}-{constellation_order}"
                                  f"{constellation_type}-{power_allocation}-SNR{snr}dB.png")
        image.save(path)
        self._mirror_to_docs(path)
        return path

    def plot_ber_vs_snr(self, results: List[Dict[str, Any]]) -> Path:
        bers = [r["bit_error_rate"] for r in results if "bit_error_rate" in r]
        snrs = [r["snr_db"] for r in results if "snr_db" in r]
        if not bers or not snrs:
            print("Warning: No BER or SNR data to plot")
            return self.images_dir / "ber_vs_snr.png"
        path = self.images_dir / (f"{_config_stem(results[0])}-BER_vs_SNR.png" if results else "ber_vs_snr.png")
        _draw_semilog(snrs, bers, "BER vs SNR Performance", "SNR (dB)", "Bit Error Rate (BER)").save(path)
        self._mirror_to_docs(path)
        return path


def _draw_semilog(xs, ys, title: str, xlabel: str, ylabel: str) -> Image.Image:
    width, height, margin = 1500, 900, 110
    img = Image.new("RGB", (width, height), "white")
    draw = ImageDraw.Draw(img)
    positive = [y for y in ys if y > 0]
    lo = np.floor(np.log10(min(positive))) if positive else -6.0
    hi = max(np.ceil(np.log10(max(positive))) if positive else 0.0, lo + 1)
    x0, x1 = min(xs), max(xs) if max(xs) > min(xs) else min(xs) + 1

    def px(x, y):
        ly = np.log10(y) if y > 0 else lo
        return (margin + (x - x0) / (x1 - x0) * (width - 2 * margin),
                height - margin - (ly - lo) / (hi - lo) * (height - 2 * margin))

    for dec in range(int(lo), int(hi) + 1):
        _, gy = px(x0, 10.0 ** dec)
        draw.line([(margin, gy), (width - margin, gy)], fill=(200, 200, 200))
        draw.text((margin - 60, gy - 6), f"1e{dec}", fill="black")
    for x in sorted(set(xs)):
        gx, _ = px(x, 10.0 ** lo)
        draw.line([(gx, margin), (gx, height - margin)], fill=(230, 230, 230))
        draw.text((gx - 10, height - margin + 8), f"{x:g}", fill="black")
    draw.rectangle([margin, margin, width - margin, height - margin], outline="black")
    pts = [px(x, y) for x, y in zip(xs, ys)]
    if len(pts) > 1:
        draw.line(pts, fill=(0, 0, 255), width=3)
    for cx, cy in pts:
        draw.ellipse([cx - 6, cy - 6, cx + 6, cy + 6], fill=(0, 0, 255))
    draw.text((width // 2 - 4 * len(title), 40), title, fill="black")
    draw.text((width // 2 - 30, height - 50), xlabel, fill="black")
    draw.text((10, height // 2), ylabel, fill="black")
    return img


class SimulationRunner:
    def __init__(self, settings: Settings, simulation_settings: SimulationSettings, results_manager: ResultsManager):
        self.settings = settings
        self.simulation_settings = simulation_settings
        self.results_manager = results_manager

    def run_all(self) -> List[Dict[str, Any]]:
        print("=" * 80)
        print(f"  {self.settings.project_name} v{self.settings.version}")
        print("=" * 80)
        print(f"\n{self.simulation_settings}\n")
        simulations = Simulation.create_from_simulation_settings(self.simulation_settings)
        print(f"Created {len(simulations)} simulation(s) to run\n")
        # the reference runs sim.run() one SNR after the other (main.py:234-240); here the SNR points that share a link
        # are one launch, and the per-simulation report blocks are printed as the results are unpacked
        results = Simulation.run_sweep(simulations)
        for i, (sim, result) in enumerate(zip(simulations, results), start=1):
            print(f"\n{'#' * 80}\n  Simulation {i}/{len(simulations)} (SNR = {sim.snr_db} dB)\n{'#' * 80}\n")
            print(f"\n  Simulation {i} completed")
            print(f"    BER: {result['bit_error_rate']:.6e}")
            print(f"    Bit Errors: {result['bit_errors']}/{result['total_bits']}")
            print(f"    PAPR: {result['papr_db']:.2f} dB")
        return results

    def process_results(self, results: List[Dict[str, Any]]) -> None:
        if not results:
            print("Warning: No results to process")
            return
        print(f"\n{'=' * 80}\n  Processing Results\n{'=' * 80}")
        saved = []
        for result in results:
            if "constellation_plot" in result:
                image: Image.Image = result["constellation_plot"]
                saved.append(self.results_manager.save_constellation_plot(
                    image=image, prefix_type=result.get("prefix_acronym", "NONE"),
                    modulation_type=result.get("modulator_type", "OFDM"),
                    equalization_method=result.get("equalizator_type", "NONE"),
                    constellation_order=result.get("constellation_order", 16),
                    constellation_type=result.get("constellation_scheme", "QAM"),
                    power_allocation=result.get("power_allocation_acronym", "UNIFORM"),
                    snr_db=result.get("snr_db", 0.0)))
                image.close()
        print(f"  Saved {len(saved)} constellation plot(s)")
        name = results[0].get("title", "unknown").replace(" ", "_")
        for result in results:
            if "bit_error_rate" in result and "snr_db" in result:
                self.results_manager.update_ber_csv(simulation_name=name, snr_db=result["snr_db"],
                                                    bit_error_rate=result["bit_error_rate"])
        print(f"  Updated BER results CSV: {self.results_manager.csv_path}")
        print(f"  Generated BER vs SNR plot: {self.results_manager.plot_ber_vs_snr(results)}")
        bers = [r["bit_error_rate"] for r in results]
        snrs = [r["snr_db"] for r in results]
        paprs = [r["papr_db"] for r in results]
        print(f"\n{'=' * 80}\n  Summary Statistics\n{'=' * 80}")
        print(f"  SNR Range: {min(snrs):.1f} dB to {max(snrs):.1f} dB")
        print(f"  BER Range: {min(bers):.6e} to {max(bers):.6e}")
        print(f"  Average PAPR: {sum(paprs) / len(paprs):.2f} dB")
        print("=" * 80)


def main():
    """Loads config/settings.json + config/simulation_settings.json relative to the working directory;
    swallows every exception and returns 1, like the reference (main.py:383-391)."""
    try:
        settings = Settings.from_json(file_path="config/settings.json")
        simulation_settings = SimulationSettings.from_json(file_path="config/simulation_settings.json")
        channel_name = "default"
        if simulation_settings.channel_type.value == "CUSTOM" and simulation_settings.channel_model_path:
            channel_name = Path(simulation_settings.channel_model_path).stem
        elif simulation_settings.channel_type.value == "FLAT":
            channel_name = "flat"
        manager = ResultsManager(results_dir="results", images_dir="images", channel_name=channel_name)
        runner = SimulationRunner(settings, simulation_settings, manager)
        runner.process_results(runner.run_all())
        print("\nAll simulations completed successfully!\n")
    except FileNotFoundError as exc:
        print(f"Error: Configuration file not found - {exc}")
        return 1
    except Exception as exc:  # noqa: BLE001 - reference behaviour
        print(f"Error during simulation: {exc}")
        import traceback
        traceback.print_exc()
        return 1
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
