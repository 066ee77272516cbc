#!/usr/bin/env python
"""bench.py - headline benchmark of the B200-native OFDM link simulator.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Metric (BASELINE.json): simulated bits/s of the fused OFDM Monte-Carlo kernel at N=1024 subcarriers,
64-QAM, MMSE, 8-tap multipath (config/channel_models/severe_multipath.npy, CP = 7), and the fraction of
the FP32 roofline it reaches.  A "step" is one pass of the hot path over one batch of OFDM symbols
(1e9 simulated bits per GPU); bits and noise are generated in registers (Philox), so there is no
input tensor to keep resident - the tables the kernel reads are 40 KB.

One JSON line is printed by rank 0.  See DESIGN.md "Measurement" for the definition of every key.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "ofdm-based-systems_b200"))

N_SC, ORDER, SNR_DB, PREFIX = 1024, 64, 20.0, 7
BITS_PER_OFDM = N_SC * 6
BITS_PER_STEP_PER_GPU = 1_000_000_000
SYMBOLS_PER_STEP = -(-BITS_PER_STEP_PER_GPU // BITS_PER_OFDM)          # 162 761 OFDM symbols
# algorithmic flops per OFDM symbol (SURVEY 8d / BASELINE.md 4): 10 N log2 N + 8 L (N+P) + 4 (N+P) + N (2 + 14 + 8)
F_SYM = 10 * N_SC * 10 + 8 * 8 * (N_SC + PREFIX) + 4 * (N_SC + PREFIX) + N_SC * (2 + 14 + 8)   # 197 084
WORKLOAD = "ofdm_link_fused N=1024 64-QAM MMSE CP=7 severe_multipath(8 taps) SNR=20dB, 1e9 bits/GPU/step"
METRIC = "OFDM Monte-Carlo sim bits/sec (N=1024,64-QAM,MMSE)"


def headline_taps() -> np.ndarray:
    return np.load(os.path.join(ROOT, "config", "channel_models", "severe_multipath.npy"))


# ---------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self) -> dict:
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------- CPU arms
def _cpu_chunk(args):
    """One bounded sample of the hot path on one host core: the oracle port of the reference chain."""
    seed, n_ofdm = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ofdm_oracle as oc
    setup = oc.LinkSetup(n_sc=N_SC, taps_raw=headline_taps(), snr_db=SNR_DB, order=ORDER, eq="MMSE",
                         prefix_len_override=PREFIX)
    rng = np.random.default_rng(seed)
    bits = oc.generate_bits(n_ofdm * BITS_PER_OFDM, rng)
    shape = (n_ofdm * (N_SC + PREFIX),)
    r = oc.run_link(setup, bits, n_ofdm * BITS_PER_OFDM, normals=(rng.normal(size=shape), rng.normal(size=shape)))
    return n_ofdm * BITS_PER_OFDM, r["bit_errors"]


def cpu_dsp_only(n_ofdm: int = 400) -> float:
    """bits/s of the DSP stages alone (ortho IFFT + CP, FIR + AWGN, strip + FFT + MMSE) of the
    oracle port on one core - separates the arithmetic from the reference's per-symbol Python mapping / demapping."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ofdm_oracle as oc
    setup = oc.LinkSetup(n_sc=N_SC, taps_raw=headline_taps(), snr_db=SNR_DB, order=ORDER, eq="MMSE",
                         prefix_len_override=PREFIX)
    rng = np.random.default_rng(1)
    X = (rng.integers(0, 8, (n_ofdm, N_SC)) * 2 - 7 + 1j * (rng.integers(0, 8, (n_ofdm, N_SC)) * 2 - 7)) / np.sqrt(42.0)
    best = float("inf")
    for _ in range(3):
        t0 = time.perf_counter()
        tx = oc.modulate(X, PREFIX, "CYCLIC", "OFDM")
        conv = oc.channel_convolve(tx.reshape(-1), setup.taps_chan)
        rx = conv + oc.awgn_noise(conv, SNR_DB, rng.normal(size=conv.shape), rng.normal(size=conv.shape))
        oc.demodulate(rx.reshape(-1, N_SC + PREFIX), N_SC, PREFIX, "CYCLIC", "MMSE", setup.H_eq, SNR_DB, "OFDM")
        best = min(best, time.perf_counter() - t0)
    return n_ofdm * BITS_PER_OFDM / best


def cpu_baseline_single(budget_s: float = 12.0) -> dict:
    """oracle port, one core, bounded sample of the same workload (reported baseline, not the target)."""
    n_ofdm, bits, t0 = 200, 0, time.perf_counter()
    calls = 0
    while time.perf_counter() - t0 < budget_s:
        b, _ = _cpu_chunk((1000 + calls, n_ofdm))
        bits += b
        calls += 1
    dt = time.perf_counter() - t0
    return {"value": bits / dt, "unit": "bits/s", "cores": 1, "kind": "port",
            "sample": f"{calls} x {n_ofdm} OFDM symbols ({bits} bits) of the headline workload, NumPy fp64 oracle port, {dt:.1f} s",
            "dsp_only_value": cpu_dsp_only(),
            "dsp_only_what": "IFFT + CP, FIR + AWGN, strip + FFT + MMSE only (no mapping / demapping), same port, 1 core"}


def run_reference_arm(args) -> None:
    """--impl reference: the reference's CPU algorithm (oracle port; the Python reference itself does not
    travel to the GPU box) on all host cores, same metric / config; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    n_ofdm = 100
    per_step = cores * 2                       # chunks per step -> ~0.1 Gbit per step on 64 cores
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        for w in range(args.warmup):
            pool.map(_cpu_chunk, [(w * 1000 + i, n_ofdm) for i in range(cores)])
        t0 = time.perf_counter()
        bits = 0
        for s in range(args.steps):
            res = pool.map(_cpu_chunk, [(50_000 + s * 1000 + i, n_ofdm) for i in range(per_step)])
            bits += sum(b for b, _ in res)
        dt = time.perf_counter() - t0
    value = bits / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "bits/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "step": f"{per_step} x {n_ofdm} OFDM symbols on {cores} host processes"},
            "cpu_baseline": {"value": value, "unit": "bits/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} steps x {per_step} chunks x {n_ofdm} OFDM symbols, NumPy fp64 oracle port of the reference chain"},
            "e2e": {"value": value, "unit": "bits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------- GPU arm
def run_b200_arm(args) -> None:
    import torch
    import torch.distributed as dist
    from ofdm_based_systems import _native
    from ofdm_based_systems.simulation.sweep import LinkConfig, LinkSweep

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    _native.require_gpu()
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=90))
    dev = torch.device("cuda", local)

    cfg = LinkConfig(num_subcarriers=N_SC, taps_raw=headline_taps(), constellation_order=ORDER,
                     constellation_scheme="QAM", modulator_type="OFDM", prefix_scheme="CYCLIC", prefix_length=PREFIX,
                     equalizator_type="MMSE")
    sweep = LinkSweep(cfg)
    flush = torch.empty(144 << 20, dtype=torch.uint8, device=dev)          # 151 MB > 126 MB L2
    snrs = [SNR_DB]
    S = SYMBOLS_PER_STEP

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        # ---- warm-up: W steps (>= 3), then a FIXED number of extra steps (~0.6 s under load; the count must be
        #      identical on every rank because each step ends in a collective) so that nvidia-smi (100 ms
        #      period) samples the clocks this kernel actually runs at
        w = max(args.warmup, 3) + 500
        for i in range(w):
            flush.zero_()
            sweep.enqueue(snrs, S, seed=i, weak_scaling=True)
            if i % 16 == 15:
                torch.cuda.synchronize()
        barrier()
        # ---- timed region: EXACTLY K steps, device-timed, counters stay on the device
        launches0 = _native.launch_count()
        ev0.record()
        payload = None
        for i in range(args.steps):
            flush.zero_()                                                  # L2 flush between timed iterations
            payload = sweep.enqueue(snrs, S, seed=100 + i, weak_scaling=True, kernel_events=k_ev[i])
        ev1.record()
        barrier()
    launches = _native.launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    ms_kernel = float(np.mean([a.elapsed_time(b) for a, b in k_ev]))
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    result = sweep.finalize(snrs, payload)[0]
    bits_per_step = S * BITS_PER_OFDM * world
    value = bits_per_step * args.steps / (ms_total * 1e-3)

    # ---- end to end through the public host API: build the link from host arrays (tables H2D), run, read back
    barrier()
    h2d = d2h = 0
    t0 = time.perf_counter()
    for i in range(args.steps):
        s2 = LinkSweep(cfg)
        r2 = s2.sweep(snrs, S, seed=200 + i, weak_scaling=True)[0]
        h2d, d2h = s2.link.table_bytes, 8 * (9 + world)
        s2.close()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = bits_per_step * args.steps / float(t.item())

    if rank == 0:
        # roofline of the dominant kernel (ofdm_link_fast_kernel<32>): algorithmic flops / measured launch time
        peak = _native.measure_fp32_tflops(8192)
        achieved = F_SYM * S / (ms_kernel * 1e-3) / 1e12
        traffic = None
        prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get("dram_bytes_per_launch")
            except (ValueError, OSError):
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": "bits/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "symbols_per_step_per_gpu": S, "bits_per_step": bits_per_step,
                       "untimed_steps": w, "untimed_steps_why": "max(W, 3) warm-up steps + 500 so that nvidia-smi (100 ms period) samples the clocks under this load",
                       "parallelism": f"symbol-range shards x{world}, one NCCL all-reduce per step" if world > 1 else "1 GPU (no collective)",
                       "l2": "144 MiB (151 MB > 126 MB L2) memset between timed steps, inside the timed region; the kernel's inputs are generated in registers (86 KB of tables read per launch)"},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": "bits/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "what": "LinkSweep(cfg).sweep() per step: link built from host taps / orders (tables staged in pinned memory, one H2D copy), kernel, all-reduce when N > 1, counters D2H"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak if peak and peak > 0 else None, "traffic": traffic,
                         "kernel": "ofdm_link_fast_kernel<32,...> (one launch per step)", "kernel_ms": ms_kernel,
                         "flops_per_ofdm_symbol": F_SYM, "symbols_per_launch": S,
                         "peak_source": "FFMA-chain microbenchmark run in this process (MEASURED_PEAKS.json has no FP32 figure)"},
            "check": {"bit_error_rate": result["bit_error_rate"], "bits": result["total_bits"], "papr_db": result["papr_db"]},
        }
        if world == 1:
            # replay mode on device-resident recorded streams: achieved HBM GB/s, reported as a fraction
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            from bench_replay import measure as replay_measure
            rp = replay_measure(symbols=100_000, reps=5)
            line["replay"] = {"value": rp["bits_per_s"], "unit": "bits/s", "achieved_gbs": rp["achieved_gbs"],
                              "hbm_peak_gbs": rp["hbm_peak_gbs"], "hbm_peak_source": rp["hbm_peak_source"], "frac_hbm": rp["frac_hbm"],
                              "algorithmic_bytes_per_symbol": rp["algorithmic_bytes_per_symbol"],
                              "algorithmic_tflops": rp["algorithmic_tflops"], "symbols_per_launch": rp["symbols"],
                              "what": "ofdm_link_launch_replay: recorded bits (768 B) + complex64 noise (8 248 B) per OFDM "
                                      "symbol streamed from HBM, same chain; compute-bound (FP32), so GB/s is a fraction"}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_single()
        print(json.dumps(line), flush=True)
    sweep.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
