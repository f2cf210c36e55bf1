"""Seeded read-pair streams for the active-coverage cap (ChromosomeProfile::add): dense bursts that hold the cap active
across segment borders, sparse stretches, jumps over whole segments (the read group sits out segments), and a few read
pairs that span more than a segment. Shared by tests/golden/make_golden.py (which replays them through the reference's
own ChromosomeProfile, oracle/_ref/popdel_ref_harness) and tests/test_host_logic.py."""
import numpy as np

SEEDS = list(range(12))
WINDOW_BUFFER = 200_000


def cap_stream(seed):
    """(max_load, start uint32[n], end uint32[n]) -- positions relative to the position the tables were reset to."""
    rng = np.random.default_rng([seed, 4711])
    max_load = int(rng.integers(5, 110))
    n_seg = int(rng.integers(3, 8))
    pos, out = int(rng.integers(0, 1000)), []
    while pos < n_seg * WINDOW_BUFFER:
        mode = int(rng.integers(0, 10))
        if mode >= 8:                                                       # jump, sometimes over whole segments
            pos += int(rng.integers(0, 450_000)) if rng.integers(0, 3) == 0 else int(rng.integers(0, 3000))
            nb = (pos // WINDOW_BUFFER + 1) * WINDOW_BUFFER
            if rng.integers(0, 2) and nb - pos < 100_000:
                pos = nb - int(rng.integers(0, 400))                        # resume right before a border
            continue
        length = int(rng.integers(200, 3200)) if mode < 5 else int(rng.integers(1000, 31_000))
        density = rng.uniform(0.3, 2.3) if mode < 5 else 0.02
        k = max(1, int(length * density))
        p = pos + np.cumsum(rng.exponential(1.0 / density, size=k)).astype(np.int64)
        out.append(p)
        pos = int(p[-1])
    start = np.concatenate(out)
    n = start.size
    inner = 200 + rng.integers(-70, 71, size=n)
    inner += np.where(rng.random(n) < 0.05, rng.integers(500, 3500, size=n), 0)
    inner += np.where(rng.random(n) < 0.005, rng.integers(0, 196_000, size=n), 0)
    if seed % 3 == 0:                                                       # read pairs that span more than a segment
        inner += np.where(rng.random(n) < 0.003, rng.integers(200_000, 430_000, size=n), 0)
    return max_load, start.astype(np.uint32), (start + np.maximum(inner, 0)).astype(np.uint32)
