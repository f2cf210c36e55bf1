"""Window-range sharding of the scan across ranks: contiguous ranges cut at the reference's segment
borders, so every `processSegment` group of calls lives on exactly one rank and `unifyCalls` behaves as in a
single-process run. No data-path collective: ranks only exchange their (small) call lists and timing scalars."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def segment_last_window(j: int, window_buffer: int) -> int:
    """Last window index of segment j (largest w with 30*w < (j+1)*window_buffer); = pd_seg_last_window."""
    return ((j + 1) * window_buffer - 1) // 30


def segment_aligned_ranges(n_windows: int, window_buffer: int, world: int) -> List[Tuple[int, int]]:
    """Splits windows [0, n_windows) into `world` contiguous (first_window, count) ranges whose cuts fall on
    segment borders; segments are dealt out as evenly as possible (rank r gets segments [r*S/world, (r+1)*S/world))."""
    if n_windows <= 0:
        return [(0, 0)] * world
    n_seg = (30 * (n_windows - 1)) // window_buffer + 1
    out = []
    for r in range(world):
        s0, s1 = (r * n_seg) // world, ((r + 1) * n_seg) // world
        w0 = 0 if s0 == 0 else segment_last_window(s0 - 1, window_buffer) + 1
        w1 = n_windows if s1 >= n_seg else min(n_windows, segment_last_window(s1 - 1, window_buffer) + 1)
        out.append((w0, max(0, w1 - w0)) if s1 > s0 else (w0, 0))
    return out


BATCH_BLOCK = 600_000      # lcm(200 000-bp segments, 30-bp windows): a batch that starts here has the contig's grid and borders


def contig_batches(length: int, batch_bp: int, window_buffer: int = 200_000, window: int = 30):
    """Window-range batches of a contig that does not fit one scan (scripts/run_config.py): (start, end, first_window,
    n_windows) per batch. A batch is scanned as its own contig anchored at `start`, a multiple of lcm(window_buffer, window), so
    its segment borders anchor + k * window_buffer and its window grid anchor + 30 w are the whole contig's; it begins one
    block (3 segments) before the windows it owns -- the halo the segment artefacts and the longest read pairs need -- and
    owns windows [first_window, first_window + n_windows) of ITS grid. The owned windows of all batches partition the contig's."""
    block = int(np.lcm(window_buffer, window))
    per = max(1, int(batch_bp) // block)
    n_blocks = (int(length) + block - 1) // block
    out = []
    for b0 in range(0, n_blocks, per):
        b1 = min(b0 + per, n_blocks)
        halo = block if b0 > 0 else 0
        start, end = b0 * block - halo, min(b1 * block, int(length))
        first = halo // window
        out.append((start, end, first, (end - start + window - 1) // window - first))
    return out


def gather_calls(calls: np.ndarray, per_sample: np.ndarray, rank: int, world: int, dist=None):
    """Host-side merge of the per-rank call lists on rank 0, in genomic order (ranks hold ascending ranges)."""
    if world == 1 or dist is None:
        return calls, per_sample
    objs = [None] * world if rank == 0 else None
    dist.gather_object((calls, per_sample), objs, dst=0)
    if rank != 0:
        return None, None
    return np.concatenate([o[0] for o in objs]), np.concatenate([o[1] for o in objs])


def reduce_timing(elapsed_s: float, evaluations: float, dist=None, device=None):
    """(max elapsed over ranks, sum of evaluations over ranks) -- the bench contract's whole-job aggregate."""
    if dist is None:
        return elapsed_s, evaluations
    import torch
    t = torch.tensor([elapsed_s], dtype=torch.float64, device=device)
    e = torch.tensor([evaluations], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(e, op=dist.ReduceOp.SUM)
    return float(t.item()), float(e.item())


# ---- sample sharding (config 5): every rank holds a contiguous block of the samples of the SAME window range ------------
def sample_blocks(n_samples: int, world: int) -> List[Tuple[int, int]]:
    """(first sample, count) per rank; sizes differ by at most one (= api.split_samples)."""
    base, extra = divmod(int(n_samples), int(world))
    out, s0 = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((s0, n))
        s0 += n
    return out


def combine_tails(tails, window_buffer: int) -> int:
    """Number of windows the reference scans for the cohort's contig from the ranks' tail summaries (kf, S, E, E_spill)
    -- the host restatement of pd_shard_window_total (pd_shard.cu): kf = final segment of the rank's read pairs, S =
    largest start position in it, E = largest last window of end set kf, E_spill = largest spill-over last window."""
    kf = max(t[0] for t in tails)
    if kf < 0:
        return 0
    S = E = -1
    for t in tails:
        if t[0] == kf:
            S, E = max(S, t[1]), max(E, t[2])
        elif t[0] == kf - 1 and t[0] >= 0:
            E = max(E, t[3])
    stop = max(E + 2, (S + 29) // 30)
    return min(stop, segment_last_window(kf, window_buffer)) + 1


def merge_sample_shards(parts, rank: int, world: int, dist=None):
    """Every rank returns the same calls and the per-sample rows of ITS samples: rank 0 concatenates the rows along
    the sample axis in rank (= cohort) order. parts = (calls, per_sample) of this rank."""
    if world == 1 or dist is None:
        return parts
    objs = [None] * world if rank == 0 else None
    dist.gather_object(parts, objs, dst=0)
    if rank != 0:
        return None, None
    for o in objs[1:]:
        if not np.array_equal(o[0], objs[0][0]):
            raise RuntimeError("sample-sharded ranks disagree on the window calls")
    return objs[0][0], np.concatenate([o[1] for o in objs], axis=1)
