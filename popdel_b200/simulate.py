"""Synthetic cohort generator (read pairs with planted deletions).

Follows the measurement spec in SURVEY.md section 8(d): read length 150, coverage 30x
(read-pair start density 0.1/bp), insert size ~ round(N(mu, sigma^2)) clipped to
(2*readLen, 20000), deletions planted per donor haplotype and mapped back to reference
coordinates (pairs spanning the breakpoint get isize += L, pairs inside the deleted
segment on carrier haplotypes do not exist).

A read pair is represented exactly like a record of the reference's profile format
(`/root/reference/popdel_profile/window_podel.h:161-197`): `pos` = reference end of the
forward read, `isize` = outer distance; the profile stores `isize - median`.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np


@dataclass
class ReadGroupSpec:
    name: str
    mu: float = 500.0
    sigma: float = 50.0
    read_length: int = 150
    coverage: float = 30.0


@dataclass
class Deletion:
    start: int          # first deleted reference base (0-based)
    length: int
    genotypes: np.ndarray  # per sample: 0, 1 or 2 deleted haplotypes


@dataclass
class ReadGroupData:
    spec: ReadGroupSpec
    pos: np.ndarray      # uint32, sorted by (pos, isize)
    isize: np.ndarray    # int32
    median: int = 0
    stddev: float = 0.0
    hist_start: int = 0
    hist_end: int = 0
    hist_counts: np.ndarray = field(default_factory=lambda: np.zeros(0))

    @property
    def dev(self) -> np.ndarray:
        return (self.isize.astype(np.int64) - int(self.median)).astype(np.int32)


@dataclass
class SampleData:
    name: str
    read_groups: List[ReadGroupData]


def _simulate_haplotype(rng, contig_len, spec: ReadGroupSpec, dels, lo=0, hi=None):
    """Fragments of one haplotype. `dels` = sorted list of (start, length) carried by it."""
    hi = contig_len if hi is None else hi
    removed = sum(l for s, l in dels if lo <= s < hi)
    donor_len = (hi - lo) - removed
    density = spec.coverage / (2.0 * spec.read_length) / 2.0   # pairs per bp per haplotype
    n = rng.poisson(density * donor_len)
    f = rng.integers(0, max(donor_len, 1), size=n, dtype=np.int64)          # donor fragment start
    isz = np.rint(rng.normal(spec.mu, spec.sigma, size=n)).astype(np.int64)
    isz = np.clip(isz, 2 * spec.read_length + 1, 19999)
    p = f + spec.read_length - 1            # donor: last base of the forward read
    q = f + isz - spec.read_length          # donor: first base of the reverse read
    keep = (f + isz) <= donor_len
    shift_p = np.zeros(n, dtype=np.int64)
    shift_q = np.zeros(n, dtype=np.int64)
    # donor coordinate of each junction = ref start - (deleted bases before it)
    acc = 0
    for s, l in dels:
        if not (lo <= s < hi):
            continue
        j = (s - lo) - acc                   # donor coordinate of the junction
        acc += l
        # a junction inside a read (split read) is dropped, as an aligner would clip/filter it
        in_fwd = (f < j) & (j <= p)
        in_rev = (q < j) & (j < f + isz)
        keep &= ~(in_fwd | in_rev)
        shift_p += np.where(p >= j, l, 0)
        shift_q += np.where(q >= j, l, 0)
    ref_p = p + shift_p + lo
    ref_isz = isz + (shift_q - shift_p)
    keep &= ref_isz < 2 ** 26
    return ref_p[keep], ref_isz[keep]


def simulate_sample(rng, name: str, contig_len: int, rgs: Sequence[ReadGroupSpec],
                    dels: Sequence[Deletion], sample_idx: int, lo=0, hi=None) -> SampleData:
    out = []
    for spec in rgs:
        ps, iss = [], []
        for hap in (0, 1):
            carried = [(d.start, d.length) for d in dels if int(d.genotypes[sample_idx]) > hap]
            # per-RG coverage is split over RGs by the caller (spec.coverage is per RG)
            p, i = _simulate_haplotype(rng, contig_len, spec, sorted(carried), lo, hi)
            ps.append(p)
            iss.append(i)
        p = np.concatenate(ps)
        i = np.concatenate(iss)
        order = np.lexsort((i, p))
        rg = ReadGroupData(spec=spec, pos=p[order].astype(np.uint32), isize=i[order].astype(np.int32))
        finalize_histogram(rg)
        out.append(rg)
    return SampleData(name=name, read_groups=out)


def finalize_histogram(rg: ReadGroupData, nominal: bool = True) -> None:
    """Header histogram as `popdel profile` would write it: raw counts of the sample's own
    insert sizes restricted to [max(1, floor(median-3sd)), ceil(median+3sd)+1)
    (`/root/reference/insert_histogram_popdel.h:80-90,128-166`)."""
    spec = rg.spec
    isz = rg.isize.astype(np.int64)
    if isz.size:
        core = isz[np.abs(isz - spec.mu) <= 6 * spec.sigma]
        med = int(np.sort(core)[(core.size - 1) // 2]) if core.size else int(round(spec.mu))
        sd = float(core.std()) if core.size > 1 else float(spec.sigma)
    else:
        med, sd = int(round(spec.mu)), float(spec.sigma)
    if nominal or sd <= 0:
        sd = float(spec.sigma)
    rg.median = med
    rg.stddev = sd
    rg.hist_start = max(1, int(np.floor(med - 3 * sd)))
    rg.hist_end = min(19999, int(np.ceil(med + 3 * sd)) + 1)
    counts = np.bincount(isz[(isz >= rg.hist_start) & (isz < rg.hist_end)] - rg.hist_start,
                         minlength=rg.hist_end - rg.hist_start).astype(np.float64)
    rg.hist_counts = counts


def plant_deletions(rng, contig_len: int, n_samples: int, n_dels: int, margin: int = 20000,
                    min_len: int = 300, max_len: int = 10000, afs=(0.1, 0.3, 0.5)) -> List[Deletion]:
    dels: List[Deletion] = []
    if n_dels <= 0:
        return dels
    # non-overlapping slots, uniform positions inside each slot
    slot = (contig_len - 2 * margin) // n_dels
    for k in range(n_dels):
        length = int(np.exp(rng.uniform(np.log(min_len), np.log(max_len))))
        room = max(1, slot - length - 2000)
        start = margin + k * slot + int(rng.integers(0, room))
        af = float(afs[int(rng.integers(0, len(afs)))])
        g = rng.binomial(2, af, size=n_samples)
        if g.sum() == 0:
            g[int(rng.integers(0, n_samples))] = 1
        dels.append(Deletion(start=start, length=length, genotypes=g))
    return dels


def simulate_cohort(seed: int, n_samples: int, contig_len: int, n_dels: int,
                    rg_specs: Sequence[Sequence[ReadGroupSpec]] | None = None,
                    dels: Sequence[Deletion] | None = None, lo=0, hi=None):
    """Returns (samples, deletions). `rg_specs[s]` = read groups of sample s
    (default: one RG, N(500, 50^2), 30x, 150 bp)."""
    rng = np.random.default_rng(seed)
    if dels is None:
        dels = plant_deletions(rng, contig_len, n_samples, n_dels)
    samples = []
    for s in range(n_samples):
        specs = rg_specs[s] if rg_specs is not None else [ReadGroupSpec(name=f"rg{s}")]
        srng = np.random.default_rng([seed, s])
        samples.append(simulate_sample(srng, f"sample{s:05d}", contig_len, specs, dels, s, lo, hi))
    return samples, list(dels)


def mixed_rg_specs(seed: int, n_samples: int, coverage: float = 30.0):
    """cfg3-style cohort: 1-3 read groups per sample, mu in {350,450,550}, sigma in {30,50,80}."""
    rng = np.random.default_rng([seed, 7])
    out = []
    for s in range(n_samples):
        k = int(rng.integers(1, 4))
        specs = []
        for j in range(k):
            specs.append(ReadGroupSpec(name=f"s{s}_rg{j}", mu=float(rng.choice([350, 450, 550])),
                                       sigma=float(rng.choice([30, 50, 80])), coverage=coverage / k))
        out.append(specs)
    return out
