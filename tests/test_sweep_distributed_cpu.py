"""world_size-2 gloo test of the multi-GPU host logic: symbol-range sharding and the single
all-reduce that combines per-rank counters, power sums and maxima.  The counter blocks
are fabricated on the CPU; the GPU part of the same code path is covered by test_sweep_gpu.py."""
import os
import socket
import struct

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ofdm_based_systems.simulation.sweep import FrameSweep, LinkSweep, combine_counters, decode_counters


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _bits(x: float) -> int:
    return struct.unpack("<q", struct.pack("<d", x))[0]


def _fake_rows(rank: int, first: int, count: int, points: int) -> torch.Tensor:
    """Deterministic stand-in for what the kernel would leave in the counter block of this shard."""
    rows = torch.zeros((points, 10), dtype=torch.int64)
    for p in range(points):
        sym = np.arange(first, first + count)
        errs = int(np.sum((sym * 7 + p) % 13 == 0))
        rows[p, 0] = errs
        rows[p, 1] = count * 6144
        rows[p, 2] = errs // 2
        rows[p, 3] = count * 1024
        rows[p, 4] = count
        rows[p, 8] = _bits(float(count) * 1031 * 1.0005)
        rows[p, 9] = _bits(5.0 + 0.25 * ((rank * 3 + p) % 4))
    return rows


def _worker(rank, world, port, total, points, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = LinkSweep.shard(total, rank, world)
    payload = combine_counters(_fake_rows(rank, first, count, points), rank, world)
    res = decode_counters([float(p) for p in range(points)], payload, 1031)
    np.save(os.path.join(out_dir, f"r{rank}.npy"),
            np.array([[r["bit_errors"], r["total_bits"], r["symbol_errors"], r["num_ofdm_symbols"], r["papr_db"]] for r in res]))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [1000, 1001])
def test_two_rank_combine_equals_single_rank(tmp_path, total):
    points, world = 3, 2
    mp.spawn(_worker, args=(world, _free_port(), total, points, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    np.testing.assert_array_equal(r0, r1)                       # every rank ends with the same totals
    single = decode_counters([0.0, 1.0, 2.0], combine_counters(_fake_rows(0, 0, total, points), 0, 1), 1031)
    for p in range(points):
        assert r0[p, 0] == single[p]["bit_errors"]
        assert r0[p, 1] == single[p]["total_bits"] == total * 6144
        assert r0[p, 3] == total
    # the maximum travels through the SUM all-reduce in per-rank slots
    for p in range(points):
        expect_max = max(5.0 + 0.25 * ((r * 3 + p) % 4) for r in range(world))
        assert abs(r0[p, 4] - 10 * np.log10(expect_max / 1.0005)) < 1e-9


def test_shards_partition_the_symbol_range():
    for total in (0, 1, 7, 162761, 10 ** 9 + 3):
        for world in (1, 2, 4, 8):
            spans = [LinkSweep.shard(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _fake_run_frames(n, count, sym_per_frame, snr, *, taps=None, seed=0, point=0, first_frame=0, **kw):
    """Deterministic stand-in for _native.run_frames: per-frame counters that depend only on the GLOBAL frame index."""
    from ofdm_based_systems._native import LinkCounters
    frames = np.arange(first_frame, first_frame + count)
    errs = int(np.sum((frames * 11 + point * 3 + seed) % 17))
    if taps is not None:
        assert taps.shape[0] == count
        errs += int(np.sum(np.round(np.abs(taps[:, 0]) * 100)))
    return dict(total=LinkCounters(errs, count * sym_per_frame * n * 4, errs // 3, count * sym_per_frame * n,
                                   count * sym_per_frame, count * sym_per_frame * (n + 7),
                                   float(count * sym_per_frame * (n + 7)) * 0.999, 4.0 + float(np.max(frames % 5)) if count else 0.0))


def _frame_worker(rank, world, port, n_frames, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    taps = np.linspace(0.1, 1.0, n_frames)[:, None] * np.ones((1, 8))
    res = FrameSweep(64, run_frames=_fake_run_frames, n_taps=8, order=16, taps=taps).sweep([10.0, 20.0], n_frames, 50, seed=5)
    np.save(os.path.join(out_dir, f"f{rank}.npy"),
            np.array([[r["bit_errors"], r["total_bits"], r["num_ofdm_symbols"], r["papr_db"]] for r in res]))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [9, 1])
def test_frame_sweep_shards_frames_across_ranks(tmp_path, n_frames):
    world = 2
    mp.spawn(_frame_worker, args=(world, _free_port(), n_frames, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "f0.npy"), np.load(tmp_path / "f1.npy")
    np.testing.assert_array_equal(r0, r1)
    taps = np.linspace(0.1, 1.0, n_frames)[:, None] * np.ones((1, 8))
    single = FrameSweep(64, run_frames=_fake_run_frames, n_taps=8, order=16, taps=taps).sweep([10.0, 20.0], n_frames, 50, seed=5)
    for p in range(2):
        assert r0[p, 0] == single[p]["bit_errors"] and r0[p, 1] == single[p]["total_bits"] == n_frames * 50 * 64 * 4
        assert r0[p, 2] == n_frames * 50
        assert abs(r0[p, 3] - single[p]["papr_db"]) < 1e-9
