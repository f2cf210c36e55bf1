"""Multi-process host logic on CPU (gloo, world_size 2): segment-aligned window ranges partition the contig, the
per-rank call lists merge into the single-process order, and the timing reduction is max / sum."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from popdel_b200 import api, sharding


def test_ranges_cover_and_align():
    for n_windows, wb, world in [(1557014, 200000, 8), (8659, 200000, 2), (6667, 200000, 4), (100, 960, 3), (0, 200000, 2)]:
        r = sharding.segment_aligned_ranges(n_windows, wb, world)
        assert len(r) == world and sum(c for _, c in r) == n_windows
        pos = 0
        for w0, c in r:
            if c:
                assert w0 == pos
                pos += c
                if pos < n_windows:                      # every cut is the first window of a segment
                    assert (30 * pos) // wb == (30 * (pos - 1)) // wb + 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_windows, wb, N = 20000, 200000, 3
    ranges = sharding.segment_aligned_ranges(n_windows, wb, world)
    rng = np.random.default_rng(7)
    win = np.sort(rng.choice(n_windows, size=200, replace=False))
    calls = np.zeros(win.size, dtype=api.CALL_DTYPE)
    calls["window_position"] = 30 * win + 29
    calls["segment"] = (30 * win) // wb
    ps = rng.integers(0, 50, size=(win.size, N, 13)).astype(np.uint32)
    w0, c = ranges[rank]
    mine = (win >= w0) & (win < w0 + c)
    got_c, got_p = sharding.gather_calls(calls[mine], ps[mine], rank, world, dist)
    t, e = sharding.reduce_timing(1.0 + rank, float(c * N), dist)
    if rank == 0:
        q.put((np.array_equal(got_c, calls), np.array_equal(got_p, ps), t, e, n_windows * N))
    dist.destroy_process_group()


def _worker_samples(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N = 7
    rng = np.random.default_rng(11)
    calls = np.zeros(50, dtype=api.CALL_DTYPE)
    calls["window_position"] = 30 * np.sort(rng.choice(5000, size=50, replace=False)) + 29
    ps = rng.integers(0, 50, size=(50, N, 13)).astype(np.uint32)
    s0, n = sharding.sample_blocks(N, world)[rank]
    got_c, got_p = sharding.merge_sample_shards((calls, ps[:, s0:s0 + n]), rank, world, dist)
    if rank == 0:
        q.put((np.array_equal(got_c, calls), np.array_equal(got_p, ps)))
    dist.destroy_process_group()


def test_contig_batches_keep_the_grid_and_partition_the_windows():
    """scripts/run_config.py: a contig too large for one scan is walked in batches that are scanned as contigs of their own."""
    for length, batch in [(248_956_422, 6_000_000), (3_100_000_000, 3_000_000), (1_234_567, 600_000), (500_000, 6_000_000), (600_000, 600_000)]:
        bs = sharding.contig_batches(length, batch)
        owned = 0
        for k, (start, end, first, n) in enumerate(bs):
            assert start % 200_000 == 0 and start % 30 == 0             # same segment borders, same window grid
            assert (first == 0) == (k == 0) and first * 30 == (600_000 if k else 0)
            assert (start + 30 * first) == owned * 30                   # owned windows are contiguous on the contig's grid
            assert n > 0 and end <= length and end - start <= batch + 600_000 + 600_000
            owned += n
        assert owned == (length + 29) // 30                             # ... and partition it


def test_sample_blocks_and_tail_combination():
    assert sharding.sample_blocks(7, 2) == [(0, 4), (4, 3)] and sharding.sample_blocks(50000, 8)[7] == (43750, 6250)
    assert [n for _, n in sharding.sample_blocks(10, 4)] == api.split_samples(10, 4)
    wb = 200000
    # one rank: stop = max(E + 2, ceil(S / 30)) capped at the last window of the final segment
    assert sharding.combine_tails([(0, 150000, 5010, -1)], wb) == 5013
    assert sharding.combine_tails([(0, 199990, 6660, 6700)], wb) == 6667              # capped by the segment border
    # a rank whose read pairs end one segment earlier only contributes its spill-over entries
    assert sharding.combine_tails([(1, 200100, 6690, -1), (0, 199000, 6650, 6700)], wb) == 6703
    assert sharding.combine_tails([(1, 200100, 6690, -1), (-1, -1, -1, -1)], wb) == 6693
    assert sharding.combine_tails([(-1, -1, -1, -1)] * 2, wb) == 0


def test_merge_sample_shards_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_samples, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    same_calls, same_ps = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same_calls and same_ps


def test_gather_and_reduce_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    same_calls, same_ps, t, e, total = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same_calls and same_ps
    assert t == 2.0 and e == total
