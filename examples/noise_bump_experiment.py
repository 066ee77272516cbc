"""The reference's water-filling robustness experiment (examples/waterfilling_noise_bump_experiment.py) on the CUDA link:
CP-OFDM over Lin-Phoong P2, 64 subcarriers, 64-QAM, MMSE; coloured noise (a +3 / +6 dB bump on the top quarter band)
injected AFTER the equaliser, applied power loading with receiver compensation, block-wide renormalisation before the
demapper.  Same three scenarios, same SNR grid and the same allocator classes as the script; 2048 OFDM symbols per point
like the script, then 200 000 to show the curve without its sampling noise."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ofdm-based-systems_b200"))

from ofdm_based_systems._native import Link
from ofdm_based_systems.power_allocation.models import UniformPowerAllocation, WaterfillingPowerAllocation

taps = np.load(os.path.join(ROOT, "config", "channel_models", "Lin-Phoong_P2.npy"))
n, order = 64, 64
snrs = [0, 5, 10, 15, 20, 25, 30]
h_eq = np.fft.fft(taps, n)
gains = np.abs(h_eq) ** 2
taps_chan = taps / np.sqrt(np.sum(np.abs(taps) ** 2))


def noise_profile(bump_db):      # create_noise_profile of the script
    profile = np.ones(n)
    profile[int(0.75 * n):] = 10 ** (bump_db / 10)
    return profile


print("scenario                                symbols " + "".join(f"{s:>9d} dB" for s in snrs))
for name, allocation, bump in (("Baseline (uniform power, +3 dB bump)", "UNIFORM", 3.0),
                               ("Water-filling (+3 dB noise bump)", "WATERFILLING", 3.0),
                               ("Water-filling (+6 dB noise bump)", "WATERFILLING", 6.0)):
    profile = noise_profile(bump)
    for n_ofdm in (2048, 200_000):
        bers = []
        for snr in snrs:
            noise_power = 10 ** (-snr / 10)
            if allocation == "WATERFILLING":
                power = WaterfillingPowerAllocation(total_power=1.0, channel_gains=gains / profile, noise_power=noise_power).allocate()
                power = np.maximum(power, 1e-4)
                power = power / np.sum(power)
            else:
                power = UniformPowerAllocation(total_power=1.0, num_subcarriers=n).allocate()
            root = np.sqrt(power)
            safe = np.where(root < 1e-10, 1.0, root)
            link = Link(n, taps_chan, h_eq, np.full(n, order), prefix_type="CYCLIC", prefix_len=len(taps) - 1, equalizer="MMSE",
                        amp=root, rx_gain=1.0 / safe)
            # the script's channel is noise-free (NoNoiseModel): sigma = 0; the noise enters after the equaliser
            res = link.run_fused_renormalised(float(snr), 0.0, n_ofdm, noise_profile=profile, seed=42)
            link.close()
            bers.append(res.bit_errors / res.bits)
        print(f"{name:38s} {n_ofdm:8d} " + "".join(f"{b:12.4f}" for b in bers))
print("BER of the three scenarios (the reference's script prints the same table from 2048 OFDM symbols per point)")
