"""Sharded BER-vs-SNR sweeps on top of the CUDA link (libofdm_b200.so).

The reference loops over SNR points strictly sequentially, one ``Simulation.run()`` each
(main.py:234-240).  Here one ``LinkSweep`` owns one configured link per GPU; the WHOLE sweep is one
kernel launch (the SNR point is the grid's second dimension) over this rank's contiguous share of the
OFDM-symbol range, and the per-rank counters of ALL points are combined by a single NCCL all-reduce at
the end of the sweep.  torch is used for the stream, the device tensors that hold the
counters and ``torch.distributed``; the arithmetic is in the CUDA library.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from ofdm_based_systems import _native


@dataclass
class LinkConfig:
    """What ``Simulation.run()`` fixes before its hot loop (simulation/models.py:226-410)."""
    num_subcarriers: int
    taps_raw: np.ndarray                      # as loaded from the .npy / the 4-tap default (NOT normalised)
    constellation_order: int = 16
    constellation_scheme: str = "QAM"
    modulator_type: str = "OFDM"
    prefix_scheme: str = "CYCLIC"
    prefix_length: int = 0
    equalizator_type: str = "MMSE"
    awgn: bool = True
    orders: Optional[np.ndarray] = None      # adaptive mode: per-subcarrier orders
    amp: Optional[np.ndarray] = None         # optional per-subcarrier tx amplitude (sqrt of allocated power)
    rx_gain: Optional[np.ndarray] = None     # optional receiver compensation per subcarrier (1 / amp)
    taps_chan: np.ndarray = field(init=False)
    h_eq: np.ndarray = field(init=False)

    def __post_init__(self):
        self.taps_raw = np.asarray(self.taps_raw, dtype=np.complex128)
        power = np.sum(np.abs(self.taps_raw) ** 2)
        if power == 0:
            raise ValueError("Impulse response cannot be all zeros.")          # channel/models.py:43
        self.taps_chan = self.taps_raw / np.sqrt(power)                        # channel/models.py:14-16
        self.h_eq = np.fft.fft(self.taps_raw, self.num_subcarriers)            # simulation/models.py:263-266
        if self.orders is None:
            self.orders = np.full(self.num_subcarriers, self.constellation_order, dtype=np.int64)
        self.orders = np.asarray(self.orders, dtype=np.int64)

    @property
    def bits_per_ofdm_symbol(self) -> int:
        return int(sum(int(np.log2(o)) for o in self.orders if o > 1))

    def stream_power(self) -> float:
        """Analytic mean |.|^2 of the serial stream after the channel, the quantity AWGNoiseModel
        measures (noise/models.py:14).  With unit-power constellations on the active subcarriers the
        in-symbol sample power is (1/N) sum_k a_k |H_k|^2 (H from the normalised taps); a zero-padded
        stream spreads one symbol's energy over N + P samples."""
        cached = getattr(self, "_stream_power", None)
        if cached is not None:
            return cached
        n = self.num_subcarriers
        a = (self.orders > 1).astype(np.float64)
        if self.amp is not None:
            a = a * np.asarray(self.amp, dtype=np.float64) ** 2
        if self.modulator_type == "OFDM":
            p = float(np.mean(a * np.abs(np.fft.fft(self.taps_chan, n)) ** 2))
        else:  # SC-OFDM: the constellation symbols are the time samples (white)
            p = float(np.mean(a) * np.sum(np.abs(self.taps_chan) ** 2))
        if self.prefix_scheme == "ZERO":
            p *= n / (n + self.prefix_length)
        self._stream_power = p        # the configuration is fixed after construction
        return p

    def noise_sigma(self, snr_db: float) -> float:
        """Per-component standard deviation sqrt(P / snr_lin / 2) (noise/models.py:15-20)."""
        if not self.awgn:
            return 0.0
        return float(np.sqrt(self.stream_power() / (10 ** (snr_db / 10)) / 2))


class _DeviceWords:
    """Zero-copy torch view of device memory owned by the CUDA library."""

    def __init__(self, ptr: int, n_words: int):
        self.__cuda_array_interface__ = {"shape": (n_words,), "typestr": "<i8", "data": (ptr, False), "version": 2}


N_WORDS = 10  # 8 x uint64 counters, double power sum, uint64 bits of the double power max


def combine_counters(rows, rank: int, world: int, group=None):
    """rows: int64 tensor [points, 10] holding each point's raw counter block of THIS rank (8 uint64
    counters, the bit pattern of the double power sum, the bit pattern of the double power max).
    Packs them so that ONE SUM all-reduce combines everything: counters and power sums add (counts
    below 2^53 are exact in float64), and every rank writes its maximum into its own slot so that the
    element-wise sum carries all the maxima.  Works on any device / backend (NCCL on GPUs, gloo in the
    CPU tests).  Returns the float64 tensor [points, 9 + world]."""
    import torch
    import torch.distributed as dist
    payload = torch.zeros((rows.shape[0], 9 + world), dtype=torch.float64, device=rows.device)
    payload[:, :8] = rows[:, :8].to(torch.float64)
    payload[:, 8] = rows[:, 8].contiguous().view(torch.float64)
    payload[:, 9 + rank] = rows[:, 9].contiguous().view(torch.float64)
    if world > 1:
        dist.all_reduce(payload, op=dist.ReduceOp.SUM, group=group)
    return payload


def decode_counters(snr_dbs: Sequence[float], payload, samples_per_ofdm_symbol: int) -> List[dict]:
    host = payload.cpu().numpy()
    out = []
    for i, snr in enumerate(snr_dbs):
        c = host[i]
        bit_errors, bits, sym_errors, syms, ofdm = (int(c[0]), int(c[1]), int(c[2]), int(c[3]), int(c[4]))
        mean_p = c[8] / max(ofdm * samples_per_ofdm_symbol, 1)
        out.append(dict(snr_db=float(snr), bit_errors=bit_errors, total_bits=bits, symbol_errors=sym_errors,
                        num_constellation_symbols=syms, num_ofdm_symbols=ofdm,
                        bit_error_rate=bit_errors / bits if bits else 0.0,
                        symbol_error_rate=sym_errors / syms if syms else 0.0,
                        papr_db=float(10 * np.log10(c[9:].max() / mean_p)) if mean_p > 0 else float("inf")))
    return out


class LinkSweep:
    def __init__(self, cfg: LinkConfig, device: Optional[int] = None):
        self.cfg = cfg
        self.link = _native.Link(cfg.num_subcarriers, cfg.taps_chan, cfg.h_eq, cfg.orders,
                                 prefix_type=cfg.prefix_scheme, prefix_len=cfg.prefix_length,
                                 modulator=cfg.modulator_type, equalizer=cfg.equalizator_type,
                                 scheme=cfg.constellation_scheme, amp=cfg.amp, rx_gain=cfg.rx_gain,
                                 device=-1 if device is None else device)
        self.device = device

    def close(self):
        self.link.close()

    # ---- one rank, one point, synchronous (host result): the plain C-ABI call
    def run_point(self, snr_db: float, n_symbols: int, *, seed: int = 0x0FD3, point: int = 0, first_symbol: int = 0):
        return self.link.run_fused(snr_db, self.cfg.noise_sigma(snr_db), n_symbols, seed=seed, point=point,
                                   first_symbol=first_symbol)

    # ---- whole sweep, sharded over the ranks of a torch.distributed group
    @staticmethod
    def shard(n_symbols: int, rank: int, world: int):
        """Contiguous share [first, first + count) of the symbol range (contiguity keeps the
        inter-symbol-interference chain intact inside a shard)."""
        base, rem = divmod(n_symbols, world)
        first = rank * base + min(rank, rem)
        return first, base + (1 if rank < rem else 0)

    def enqueue(self, snr_dbs: Sequence[float], n_symbols: int, *, seed: int = 0x0FD3, group=None,
                weak_scaling: bool = False, kernel_events=None, overlap_collective: bool = False):
        """Queue the whole sweep on the current CUDA stream - ONE launch of the link kernel with the SNR point as the
        grid's second dimension (ofdm_link_launch_sweep), one launch that packs the per-point counters into the
        all-reduce payload - then ONE all-reduce per sweep; nothing is read back.  Returns the device tensor
        [points, 9 + world] (float64) that ``finalize`` decodes.
        ``kernel_events``: optional (start, end) torch CUDA events recorded around the link kernel launch.
        ``overlap_collective``: the all-reduce is issued asynchronously (it runs on the process group's own stream behind
        this sweep's kernels), so the next sweep's kernel does not wait for the slowest rank of this one; call
        ``wait_collectives`` (or ``finalize``) before reading any payload."""
        import torch
        import torch.distributed as dist
        distributed = dist.is_available() and dist.is_initialized()
        rank = dist.get_rank(group) if distributed else 0
        world = dist.get_world_size(group) if distributed else 1
        if weak_scaling:
            first, count = rank * n_symbols, n_symbols
        else:
            first, count = self.shard(n_symbols, rank, world)
        dev = torch.device("cuda", torch.cuda.current_device())
        stream = torch.cuda.current_stream().cuda_stream
        snrs = [float(x) for x in snr_dbs]
        payload = torch.empty((len(snrs), 9 + world), dtype=torch.float64, device=dev)
        if kernel_events is not None:
            kernel_events[0].record()
        self.link.launch_sweep(snrs, [self.cfg.noise_sigma(x) for x in snrs], count, seed=seed, first_symbol=first,
                               stream=stream)
        if kernel_events is not None:
            kernel_events[1].record()
        self.link.pack_sweep(payload.data_ptr(), rank, world, stream)
        if world > 1:
            work = dist.all_reduce(payload, op=dist.ReduceOp.SUM, group=group if distributed else None,
                                   async_op=overlap_collective)
            if overlap_collective:
                self._pending = getattr(self, "_pending", [])
                self._pending.append(work)
        return payload

    def wait_collectives(self) -> None:
        """Makes the current stream wait for every all-reduce issued with ``overlap_collective``."""
        for work in getattr(self, "_pending", []):
            work.wait()
        self._pending = []

    def finalize(self, snr_dbs: Sequence[float], payload) -> List[dict]:
        """Device -> host read of the combined counters; result keys follow the reference's result dict
        (bit_errors / total_bits / bit_error_rate / symbol_errors / symbol_error_rate / papr_db)."""
        self.wait_collectives()
        return decode_counters(snr_dbs, payload, self.cfg.num_subcarriers + self.cfg.prefix_length)

    def sweep_counters(self, snr_dbs: Sequence[float], n_symbols: int, *, seed: int = 0x0FD3, group=None,
                       weak_scaling: bool = False) -> List[_native.LinkCounters]:
        """The sweep as LinkCounters per point (summed over the ranks of the process group, if there is one)."""
        import torch.distributed as dist
        snrs = [float(x) for x in snr_dbs]
        if not (dist.is_available() and dist.is_initialized()):
            # one GPU: the synchronous C-ABI call (counter reset, ONE launch, counters D2H), no torch in the loop
            return self.link.run_sweep(snrs, [self.cfg.noise_sigma(x) for x in snrs], n_symbols, seed=seed)
        host = self.enqueue(snrs, n_symbols, seed=seed, group=group, weak_scaling=weak_scaling).cpu().numpy()
        spo = self.cfg.num_subcarriers + self.cfg.prefix_length
        return [_native.LinkCounters(int(c[0]), int(c[1]), int(c[2]), int(c[3]), int(c[4]), int(c[4]) * spo, float(c[8]),
                                     float(c[9:].max())) for c in host]

    def sweep(self, snr_dbs: Sequence[float], n_symbols: int, *, seed: int = 0x0FD3, group=None,
              weak_scaling: bool = False, ci_blocks: int = 0) -> List[dict]:
        """``n_symbols`` is the global OFDM-symbol count per point (the per-rank count with
        ``weak_scaling``); the symbol range is sharded over the ranks of the process group.

        ``ci_blocks`` = B > 1 additionally runs the range as B contiguous blocks of OFDM symbols (B launches, each
        sharded like the whole range, still ONE all-reduce) and adds a confidence interval from the block-to-block
        spread of the BER - errors inside an OFDM symbol are correlated through the channel and the equaliser, so the
        block is the unit, as in the reference-side intervals of tests/test_simulation_gpu.py: keys ``ber_sem``
        (standard error of the mean) and ``ber_ci95`` (mean -+ 1.96 s.e.m.)."""
        if ci_blocks and ci_blocks > 1:
            return self._sweep_blocks(snr_dbs, n_symbols, seed, group, weak_scaling, int(ci_blocks))
        return [self._result(snr, r) for snr, r in
                zip(snr_dbs, self.sweep_counters(snr_dbs, n_symbols, seed=seed, group=group, weak_scaling=weak_scaling))]

    @staticmethod
    def _result(snr, r) -> dict:
        return dict(snr_db=float(snr), bit_errors=r.bit_errors, total_bits=r.bits, symbol_errors=r.symbol_errors,
                    num_constellation_symbols=r.symbols, num_ofdm_symbols=r.ofdm_symbols,
                    bit_error_rate=r.bit_errors / r.bits if r.bits else 0.0,
                    symbol_error_rate=r.symbol_errors / r.symbols if r.symbols else 0.0, papr_db=r.papr_db)

    def _sweep_blocks(self, snr_dbs, n_symbols, seed, group, weak_scaling, blocks) -> List[dict]:
        import torch.distributed as dist
        distributed = dist.is_available() and dist.is_initialized()
        rank = dist.get_rank(group) if distributed else 0
        world = dist.get_world_size(group) if distributed else 1
        snrs = [float(x) for x in snr_dbs]
        sig = [self.cfg.noise_sigma(x) for x in snrs]
        total = n_symbols * world if weak_scaling else n_symbols
        blocks = max(2, min(blocks, total))
        edges = [total * b // blocks for b in range(blocks + 1)]
        spo = self.cfg.num_subcarriers + self.cfg.prefix_length
        if not distributed:
            per_block = [self.link.run_sweep(snrs, sig, edges[b + 1] - edges[b], seed=seed, first_symbol=edges[b])
                         for b in range(blocks)]
        else:
            import torch
            dev = torch.device("cuda", torch.cuda.current_device())
            stream = torch.cuda.current_stream().cuda_stream
            payload = torch.zeros((blocks, len(snrs), 9 + world), dtype=torch.float64, device=dev)
            for b in range(blocks):
                first, count = self.shard(edges[b + 1] - edges[b], rank, world)
                if count == 0:
                    continue
                self.link.launch_sweep(snrs, sig, count, seed=seed, first_symbol=edges[b] + first, stream=stream)
                self.link.pack_sweep(payload[b].data_ptr(), rank, world, stream)
            dist.all_reduce(payload, op=dist.ReduceOp.SUM, group=group)
            host = payload.cpu().numpy()
            per_block = [[_native.LinkCounters(int(c[0]), int(c[1]), int(c[2]), int(c[3]), int(c[4]), int(c[4]) * spo,
                                               float(c[8]), float(c[9:].max())) for c in host[b]] for b in range(blocks)]
        out = []
        for k, snr in enumerate(snrs):
            rows = [per_block[b][k] for b in range(blocks)]
            tot = _native.LinkCounters(sum(r.bit_errors for r in rows), sum(r.bits for r in rows),
                                       sum(r.symbol_errors for r in rows), sum(r.symbols for r in rows),
                                       sum(r.ofdm_symbols for r in rows), sum(r.tx_samples for r in rows),
                                       sum(r.tx_power_sum for r in rows), max(r.tx_power_max for r in rows))
            res = self._result(snr, tot)
            ber = np.array([r.bit_errors / r.bits for r in rows if r.bits])
            # blocks have (almost) equal sizes: the unweighted spread of their BERs estimates the variance of a block mean
            sem = float(ber.std(ddof=1) / np.sqrt(len(ber))) if len(ber) > 1 else float("nan")
            res.update(ber_sem=sem, ber_ci95=(res["bit_error_rate"] - 1.96 * sem, res["bit_error_rate"] + 1.96 * sem),
                       ci_blocks=len(ber))
            out.append(res)
        return out


class FrameSweep:
    """BER-vs-SNR sweep over a batch of channel realisations ("fresh channel per frame", BASELINE configs #2 / #4):
    every SNR point is one ``ofdm_frames_run`` call over this rank's contiguous share of the frame range, and ONE
    all-reduce combines the per-rank totals of all points.  Frame f draws its taps and its OFDM symbols from Philox
    counters derived from the global frame index, so the union is the same for 1, 2, 4 or 8 GPUs.

    ``frame_kwargs`` are those of ``_native.run_frames`` (n_taps, equalizer, order or waterfilling / min_order /
    max_order / ser, taps, prefix_len)."""

    def __init__(self, num_subcarriers: int, run_frames=None, **frame_kwargs):
        self.num_subcarriers = int(num_subcarriers)
        self.frame_kwargs = frame_kwargs
        self._run = run_frames if run_frames is not None else _native.run_frames

    def sweep(self, snr_dbs: Sequence[float], n_frames: int, symbols_per_frame: int, *, seed: int = 0x0FD3, group=None,
              device=None) -> List[dict]:
        import struct
        import torch
        import torch.distributed as dist
        distributed = dist.is_available() and dist.is_initialized()
        rank = dist.get_rank(group) if distributed else 0
        world = dist.get_world_size(group) if distributed else 1
        first, count = LinkSweep.shard(n_frames, rank, world)
        kw = dict(self.frame_kwargs)
        taps = kw.pop("taps", None)
        if taps is not None:
            taps = np.atleast_2d(taps)[first:first + count]
        rows = torch.zeros((len(snr_dbs), N_WORDS), dtype=torch.int64)
        bits_of = lambda x: struct.unpack("<q", struct.pack("<d", float(x)))[0]
        for i, snr in enumerate(snr_dbs):
            if count == 0:
                continue
            r = self._run(self.num_subcarriers, count, symbols_per_frame, float(snr), taps=taps, seed=seed, point=i,
                          first_frame=first, per_frame=False, want_orders=False, want_taps=False, **kw)["total"]
            rows[i, :5] = torch.tensor([r.bit_errors, r.bits, r.symbol_errors, r.symbols, r.ofdm_symbols], dtype=torch.int64)
            rows[i, 8], rows[i, 9] = bits_of(r.tx_power_sum), bits_of(r.tx_power_max)
        if device is None and distributed and dist.get_backend(group) == "nccl":
            device = torch.device("cuda", torch.cuda.current_device())
        if device is not None:
            rows = rows.to(device)
        payload = combine_counters(rows, rank, world, group if distributed else None)
        n_taps = int(kw.get("n_taps", 8)) if taps is None else int(taps.shape[1])
        prefix = kw.get("prefix_len")
        return decode_counters(snr_dbs, payload, self.num_subcarriers + (n_taps - 1 if prefix is None else int(prefix)))
