"""BASELINE-sized parity cases whose profiles are NOT committed: the seeded simulator regenerates them (a few seconds),
only what the unmodified reference wrote for them is kept under tests/golden/<case>/ (merged.vcf, win.vcf.gz,
harness.dump.gz; made by tests/golden/make_golden.py). SURVEY.md 8c / VERDICT r01 item 7: the rank-indexed thresholds
(genotype_deletion_popdel_call.h:58-86) and the 100-sample shape of the fused EM kernel are pinned on the reference itself."""
import os

import numpy as np

from popdel_b200 import profile_format as pf
from popdel_b200 import simulate as sim


def big100_samples():
    rng = np.random.default_rng(2024)
    dels = [sim.Deletion(150_000, 900, rng.binomial(2, 0.3, size=100)), sim.Deletion(380_000, 2600, rng.binomial(2, 0.05, size=100)),
            sim.Deletion(610_000, 420, rng.binomial(2, 0.5, size=100)), sim.Deletion(845_000, 6000, rng.binomial(2, 0.15, size=100))]
    samples, _ = sim.simulate_cohort(seed=101, n_samples=100, contig_len=1_000_000, n_dels=0, dels=dels)
    return samples


def big100(d):
    """100 samples, one read group each, 1 Mbp (5 segments), 4 planted deletions with allele frequencies 0.05 .. 0.5."""
    return pf.write_cohort(d, big100_samples(), ("chr21", 1_000_000))


def mixed200_samples():
    rng = np.random.default_rng(2025)
    dels = [sim.Deletion(90_000, 1400, rng.binomial(2, 0.25, size=200)), sim.Deletion(260_000, 3500, rng.binomial(2, 0.1, size=200)),
            sim.Deletion(410_000, 650, rng.binomial(2, 0.4, size=200))]
    specs = sim.mixed_rg_specs(31, 200)
    samples, _ = sim.simulate_cohort(seed=102, n_samples=200, contig_len=500_000, n_dels=0, dels=dels, rg_specs=specs)
    return samples


def mixed200(d):
    """200 samples with 1-3 read groups of different median / sigma, 0.5 Mbp, 3 planted deletions."""
    return pf.write_cohort(d, mixed200_samples(), ("chr1", 500_000))


BIG = dict(big100=big100, mixed200=mixed200)


def profile_list(case, d):
    """Writes the case's profiles and profiles.txt into d; returns the paths."""
    os.makedirs(d, exist_ok=True)
    paths = BIG[case](d)
    with open(os.path.join(d, "profiles.txt"), "w") as fh:
        fh.write("\n".join(os.path.basename(p) for p in paths) + "\n")
    return paths
