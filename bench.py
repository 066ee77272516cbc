#!/usr/bin/env python
"""bench.py - headline benchmark of the B200-native OFDM link simulator.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Metric (BASELINE.json): simulated bits/s of the fused OFDM Monte-Carlo kernel at N=1024 subcarriers,
64-QAM, MMSE, 8-tap multipath (config/channel_models/severe_multipath.npy, CP = 7), and the fraction of
the FP32 roofline it reaches.  A "step" is one pass of the hot path over one batch: the BER-vs-SNR sweep of
SURVEY 8(d) (0:2:30 dB, 16 points, 1e9 simulated bits per point and GPU) as ONE kernel launch; bits and noise
are generated in registers (Philox), so there is no input tensor to keep resident - the tables the kernel
reads are 40 KB.

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
BITS_PER_POINT_PER_GPU = 1_000_000_000
SYMBOLS_PER_POINT = -(-BITS_PER_POINT_PER_GPU // BITS_PER_OFDM)        # 162 761 OFDM symbols
SNR_GRID = [float(x) for x in range(0, 31, 2)]                         # 0:2:30 dB, 16 points
# algorithmic flops per OFDM symbol (SURVEY 8d / BASELINE.md 4): 10 N log2 N + 8 L (N+P) + 4 (N+P) + N (2 + 14 + 8)
F_SYM = 10 * N_SC * 10 + 8 * 8 * (N_SC + PREFIX) + 4 * (N_SC + PREFIX) + N_SC * (2 + 14 + 8)   # 197 084
FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12                   # 74.45: 148 SMs x 128 FFMA lanes x 2 x max boost
WORKLOAD = ("ofdm_link_fused N=1024 64-QAM MMSE CP=7 severe_multipath(8 taps), BER-vs-SNR sweep 0:2:30 dB (16 points), "
            "1e9 bits/point/GPU/step")
METRIC = "OFDM Monte-Carlo sim bits/sec (N=1024,64-QAM,MMSE)"
STRONG_SEED = 0x0FD3


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
def _port_chunk(args):
    """One bounded sample of the hot path on one host core: the NumPy oracle port of the reference chain."""
    seed, n_ofdm = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ofdm_oracle as oc
    setup = oc.LinkSetup(n_sc=N_SC, taps_raw=headline_taps(), snr_db=SNR_DB, order=ORDER, eq="MMSE",
                         prefix_len_override=PREFIX)
    rng = np.random.default_rng(seed)
    bits = oc.generate_bits(n_ofdm * BITS_PER_OFDM, rng)
    shape = (n_ofdm * (N_SC + PREFIX),)
    t0 = time.perf_counter()
    r = oc.run_link(setup, bits, n_ofdm * BITS_PER_OFDM, normals=(rng.normal(size=shape), rng.normal(size=shape)))
    return n_ofdm * BITS_PER_OFDM, r["bit_errors"], time.perf_counter() - t0


def _reference_chunk(args):
    """One bounded sample on one host core through the UNMODIFIED reference (oracle/_ref, see oracle/ref_pipeline.py).
    The SNR point cycles through the sweep grid so that the sample covers the same workload."""
    seed, n_ofdm = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_pipeline
    return ref_pipeline.run_chunk((seed, n_ofdm, N_SC, ORDER, SNR_GRID[seed % len(SNR_GRID)], headline_taps()))


def reference_available() -> bool:
    return os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "src", "ofdm_based_systems"))


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


def cpu_baseline_single(budget_s: float = 14.0) -> dict:
    """The reference itself (oracle/_ref) on ONE host core - it is single-threaded - over a bounded sample of the same
    workload; the NumPy port and its DSP-only figure beside it (reported baselines, not the target).  The reference runs
    in a child process because its package has the same import name as the product's."""
    import multiprocessing as mp
    out = {}
    n_port, bits, calls, t0 = 200, 0, 0, time.perf_counter()
    while time.perf_counter() - t0 < 4.0:
        b, _, _ = _port_chunk((1000 + calls, n_port))
        bits += b
        calls += 1
    port = bits / (time.perf_counter() - t0)
    if reference_available():
        n_ofdm, bits, calls, t0 = 100, 0, 0, time.perf_counter()
        with mp.get_context("spawn").Pool(1) as pool:
            while time.perf_counter() - t0 < budget_s:
                b, _, _ = pool.apply(_reference_chunk, ((calls, n_ofdm),))
                bits += b
                calls += 1
        dt = time.perf_counter() - t0
        out = {"value": bits / dt, "unit": "bits/s", "cores": 1, "kind": "reference",
               "sample": f"{calls} x {n_ofdm} OFDM symbols ({bits} bits) of the headline workload cycling through the SNR grid, "
                         f"unmodified reference component pipeline (oracle/_ref), {dt:.1f} s"}
    else:
        out = {"value": port, "unit": "bits/s", "cores": 1, "kind": "port",
               "sample": "oracle/_ref absent: NumPy fp64 oracle port, 4 s of 200-symbol chunks"}
    out.update({"port_value": port, "port_what": "NumPy fp64 oracle port of the same chain, 1 core, 4 s sample",
                "dsp_only_value": cpu_dsp_only(),
                "dsp_only_what": "IFFT + CP, FIR + AWGN, strip + FFT + MMSE only (no mapping / demapping), same port, 1 core"})
    return out


def arm_config(world: int, warmup: int) -> dict:
    """`config` of the JSON line - ONE definition for both arms, so that the driver sees the same object on the reference
    line (`--impl reference` runs the reference "on your arm's config"); what only the GPU arm does (L2 flush, untimed
    steps, collective) is said here once and is simply not applicable to the CPU run."""
    bits_per_step = SYMBOLS_PER_POINT * BITS_PER_OFDM * len(SNR_GRID) * world      # whole job
    return {"workload": WORKLOAD, "snr_grid_db": SNR_GRID, "symbols_per_point_per_gpu": SYMBOLS_PER_POINT, "bits_per_step": bits_per_step,
            "untimed_steps": max(warmup, 3) + 30,
            "untimed_steps_why": "max(W, 3) warm-up steps + 30 (~0.5 s) so that nvidia-smi (100 ms period) samples the clocks under this load",
            "parallelism": f"symbol-range shards x{world}, ONE NCCL all-reduce per sweep (= per step), issued on the process group's stream behind the step's kernel so that the next step's kernel does not wait for it" if world > 1 else "1 GPU (no collective)",
            "l2": "144 MiB (151 MB > 126 MB L2) memset between timed steps, inside the timed region; the kernel's inputs are generated in registers (86 KB of tables read per launch)"}


def run_reference_arm(args) -> None:
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref: the unmodified package, copied
    by oracle/build_ref.py; the NumPy port only if that copy is absent) on all host cores, one independent process per
    core as SURVEY 8d-ii prescribes, same metric / config; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    use_ref = reference_available()
    chunk, n_ofdm, kind = (_reference_chunk, 25, "reference") if use_ref else (_port_chunk, 100, "port")
    per_step = cores                              # one chunk per core and step: ~0.3 s per step for the reference
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        for w in range(args.warmup):
            pool.map(chunk, [(w * 1000 + i, n_ofdm) for i in range(cores)])
        t0 = time.perf_counter()
        bits = 0
        for s_ in range(args.steps):
            res = pool.map(chunk, [(50_000 + s_ * 1000 + i, n_ofdm) for i in range(per_step)])
            bits += sum(r[0] for r in res)
        dt = time.perf_counter() - t0
    value = bits / dt
    what = ("unmodified reference component pipeline (oracle/_ref)" if use_ref else "NumPy fp64 oracle port of the reference chain")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "bits/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": arm_config(args.gpus, args.warmup),   # the GPU arm's config, verbatim: same workload, same grid
            "step_sample": f"{per_step} x {n_ofdm} OFDM symbols on {cores} host processes per step (a bounded sample of the workload)",
            "cpu_baseline": {"value": value, "unit": "bits/s", "cores": cores, "kind": kind,
                             "sample": f"{args.steps} steps x {per_step} chunks x {n_ofdm} OFDM symbols cycling through the SNR grid, {what}"},
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
    snrs, K, S = SNR_GRID, len(SNR_GRID), SYMBOLS_PER_POINT

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        # ---- warm-up: W steps (>= 3) plus a FIXED number of extra steps (identical on every rank: each step ends in a
        #      collective), ~0.5 s under load so that nvidia-smi (100 ms period) samples the clocks of this kernel
        w = max(args.warmup, 3) + 30
        for i in range(w):
            flush.zero_()
            sweep.enqueue(snrs, S, seed=i, weak_scaling=True, overlap_collective=True)
        sweep.wait_collectives()
        barrier()
        # ---- timed region: EXACTLY K steps, device-timed, counters stay on the device
        launches0 = _native.launch_count()
        ev0.record()
        payload = None
        for i in range(args.steps):
            flush.zero_()                                                  # L2 flush between timed iterations
            payload = sweep.enqueue(snrs, S, seed=100 + i, weak_scaling=True, kernel_events=k_ev[i], overlap_collective=True)
        sweep.wait_collectives()                                           # every step's all-reduce is inside the timed region
        ev1.record()
        barrier()
    launches = _native.launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    ms_kernel = float(np.mean([a.elapsed_time(b) for a, b in k_ev]))
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    result = sweep.finalize(snrs, payload)
    bits_per_step = K * S * BITS_PER_OFDM * world
    value = bits_per_step * args.steps / (ms_total * 1e-3)

    # ---- end to end through the public host API: build the link from host arrays (tables H2D), run the sweep, read back
    barrier()
    h2d = d2h = 0
    t0 = time.perf_counter()
    for i in range(args.steps):
        s2 = LinkSweep(cfg)
        s2.sweep(snrs, S, seed=200 + i, weak_scaling=True)
        h2d, d2h = s2.link.table_bytes, (8 * (9 + world) if world > 1 else 80) * K
        s2.close()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = bits_per_step * args.steps / float(t.item())

    # ---- shard invariance on the hardware: a FIXED union of work (strong scaling, seed 0x0FD3, 162 761 OFDM symbols per
    #      point in total) split over the ranks - the bit-error counts must be the same integers for N = 1, 2, 4, 8
    strong = sweep.sweep(snrs, S, seed=STRONG_SEED, weak_scaling=False)
    # ---- BASELINE config #5 (N=4096, 256-QAM, MMSE, ~1e12 bits per point near the BER 1e-9 crossing), strong scaling
    c5 = None
    if not args.no_c5:
        cfg5 = LinkConfig(num_subcarriers=4096, taps_raw=headline_taps(), constellation_order=256, prefix_length=PREFIX,
                          equalizator_type="MMSE")
        s5 = LinkSweep(cfg5)
        n5 = 10 ** 12 // (4096 * 8)
        s5.sweep([39.0], 4000, seed=1)
        barrier()
        t0 = time.perf_counter()
        r5 = s5.sweep([39.0, 39.5], n5, seed=STRONG_SEED, weak_scaling=False)
        torch.cuda.synchronize()
        t5 = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t5, op=dist.ReduceOp.MAX)
        c5 = {"workload": "N=4096 256-QAM MMSE severe_multipath CP=7, 1e12 bits per point at 39 / 39.5 dB, symbol range split over the GPUs",
              "bits": int(sum(r["total_bits"] for r in r5)), "seconds": float(t5.item()),
              "bits_per_s": sum(r["total_bits"] for r in r5) / float(t5.item()),
              "bit_errors": [int(r["bit_errors"]) for r in r5], "ber": [r["bit_error_rate"] for r in r5]}
        s5.close()

    if rank == 0:
        # roofline of the dominant kernel (ofdm_link_fast_kernel<32, 32, ...>, one launch per step covering the 16 points):
        # algorithmic flops / measured launch time
        peak = _native.measure_fp32_tflops(8192)
        achieved = F_SYM * S * K / (ms_kernel * 1e-3) / 1e12
        traffic = None
        prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get("dram_bytes_per_launch")
            except (ValueError, OSError):
                traffic = None
        at20 = result[SNR_GRID.index(SNR_DB)]
        line = {
            "metric": METRIC, "value": value, "unit": "bits/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": arm_config(world, args.warmup),
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": "bits/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "what": "LinkSweep(cfg).sweep(grid) per step: link built from host taps / orders (tables staged in pinned memory, one H2D copy), ONE kernel launch for the 16 points, all-reduce when N > 1, counters D2H"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak if peak and peak > 0 else None,
                         "frac_nominal": achieved / FP32_NOMINAL_TFLOPS, "peak_nominal": FP32_NOMINAL_TFLOPS, "traffic": traffic,
                         "kernel": "ofdm_link_fast_kernel<32,32,...> (one launch per step, grid = SMs x 16 SNR points)", "kernel_ms": ms_kernel,
                         "flops_per_ofdm_symbol": F_SYM, "symbols_per_launch": S * K,
                         "peak_source": "FFMA-chain microbenchmark run in this process (MEASURED_PEAKS.json has no FP32 figure); "
                                        "frac_nominal uses 148 SMs x 128 lanes x 2 x 1.965 GHz = 74.45 TFLOP/s",
                         "rng": "bits and noise from Philox4x32-7 (the crush-resistant minimum, DESIGN.md 3.3); the library built with "
                                "OFDM_FAST_PHILOX_ROUNDS=10 measures frac 0.503 with this same command (profiles/r2_bench_line_philox10.json)"},
            "check": {"bit_error_rate_20db": at20["bit_error_rate"], "bits_20db": at20["total_bits"], "papr_db": at20["papr_db"],
                      "strong_bit_errors": int(sum(r["bit_errors"] for r in strong)),
                      "strong_bit_errors_per_point": [int(r["bit_errors"]) for r in strong],
                      "strong_what": f"seed {STRONG_SEED:#x}, {S} OFDM symbols per point in TOTAL split over the GPUs: identical integers for every N"},
        }
        if c5 is not None:
            line["c5"] = c5
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
    ap.add_argument("--no-c5", action="store_true", help="skip the BASELINE config #5 sub-record (2 x 1e12 bits)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
