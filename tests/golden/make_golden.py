#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/ (run in the authoring container only).

For every case: synthetic profiles in the reference's binary format (popdel_b200.profile_format), then the
UNMODIFIED reference built by oracle/Makefile is run on them:
  * oracle/_ref/popdel_ref call         -> merged.vcf      (end-to-end output)
  * oracle/_ref/popdel_ref call -n      -> win.vcf         (window-wise output)
  * oracle/_ref/popdel_ref_harness      -> harness.dump.gz (raw window calls per segment, merged calls,
                                                            optionally per-window active-set checksums)
The committed outputs pin oracle/popdel_oracle.cpp (tests/test_oracle_vs_reference.py) and, through it,
the CUDA path. Usage: python tests/golden/make_golden.py [case ...]
"""
import gzip
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from popdel_b200 import profile_format as pf  # noqa: E402
from popdel_b200 import simulate as sim  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "popdel_ref")
HARNESS = os.path.join(ROOT, "oracle", "_ref", "popdel_ref_harness")


def _cut(rg, keep):
    rg.pos = rg.pos[keep]
    rg.isize = rg.isize[keep]


def case_basic(d):
    """6 samples, one RG each, 450 kbp (two segment borders), 3 planted deletions; the reads start near
    position 0 so the index-jump loss at segment borders (SURVEY App. D / DESIGN) is exercised."""
    samples, _ = sim.simulate_cohort(seed=11, n_samples=6, contig_len=450_000, n_dels=3)
    return pf.write_cohort(d, samples, ("chr21", 450_000)), {}


def case_mixedrg(d):
    """5 samples with 1-3 read groups of different median/sigma (rank-indexed thresholds, per-RG tables)."""
    specs = sim.mixed_rg_specs(5, 5)
    samples, _ = sim.simulate_cohort(seed=12, n_samples=5, contig_len=330_000, n_dels=3, rg_specs=specs)
    return pf.write_cohort(d, samples, ("chr1", 330_000)), {}


def case_gap(d):
    """4 samples, 900 kbp, a 430 kbp hole in every sample (segments without data, stale scan position) and
    an additional private hole in sample 1; one deletion on either side."""
    dels = [sim.Deletion(120_000, 1500, np.array([1, 2, 0, 1])), sim.Deletion(760_000, 2500, np.array([0, 1, 1, 2]))]
    samples, _ = sim.simulate_cohort(seed=13, n_samples=4, contig_len=900_000, n_dels=0, dels=dels)
    for k, s in enumerate(samples):
        for rg in s.read_groups:
            keep = ~((rg.pos >= 230_000) & (rg.pos < 660_000))
            if k == 1:
                keep &= ~((rg.pos >= 40_000) & (rg.pos < 75_000))
            _cut(rg, keep)
    return pf.write_cohort(d, samples, ("chr2", 900_000)), {}


def case_highcov(d):
    """3 samples, 260 kbp; sample 0 carries a 25 kbp region at ~12x the normal coverage so that the
    active-coverage cap (default 100 pairs per read group) drops reads at load time and flags the RG."""
    dels = [sim.Deletion(60_000, 1200, np.array([1, 1, 2])), sim.Deletion(150_000, 900, np.array([2, 0, 1]))]
    samples, _ = sim.simulate_cohort(seed=14, n_samples=3, contig_len=260_000, n_dels=0, dels=dels)
    rng = np.random.default_rng(99)
    spec = sim.ReadGroupSpec(name="burst", coverage=330.0)
    extra, _ = sim.simulate_cohort(seed=15, n_samples=1, contig_len=260_000, n_dels=0, dels=dels[:0], rg_specs=[[spec]])
    e = extra[0].read_groups[0]
    m = (e.pos >= 140_000) & (e.pos < 165_000)
    rg = samples[0].read_groups[0]
    p = np.concatenate([rg.pos, e.pos[m]])
    i = np.concatenate([rg.isize, e.isize[m]])
    o = np.lexsort((i, p))
    rg.pos, rg.isize = p[o], i[o]
    del rng
    return pf.write_cohort(d, samples, ("chr3", 260_000)), {}


def case_twocontigs(d):
    """3 samples on two contigs (210 kbp and 150 kbp) plus an empty third contig: ROI change, full reset."""
    lens = [("chrA", 210_000), ("chrB", 150_000), ("chrC", 50_000)]
    a, _ = sim.simulate_cohort(seed=16, n_samples=3, contig_len=lens[0][1], n_dels=2)
    b, _ = sim.simulate_cohort(seed=17, n_samples=3, contig_len=lens[1][1], n_dels=1)
    os.makedirs(d, exist_ok=True)
    paths = []
    for s in range(3):
        ra, rb = a[s].read_groups[0], b[s].read_groups[0]
        rb.median = ra.median                 # one header histogram per RG: reuse contig A's
        meta = [pf.rg_meta_from(ra)]
        recs = {0: [(ra.pos, ra.isize)], 1: [(rb.pos, rb.isize)]}
        p = os.path.join(d, a[s].name + ".profile")
        pf.write_profile(p, meta, lens, recs)
        paths.append(p)
    return paths, {}


def case_offset(d):
    """4 samples whose reads start at 7 kbp (anchor not near a 10 kbp index boundary: no index-jump loss),
    with a user-set minimum initial length (-l 150)."""
    samples, _ = sim.simulate_cohort(seed=18, n_samples=4, contig_len=420_000, n_dels=3)
    for s in samples:
        for rg in s.read_groups:
            _cut(rg, rg.pos >= 7_000)
    return pf.write_cohort(d, samples, ("chr5", 420_000)), {"args": ["-l", "150"]}


def case_capfar(d):
    """3 samples, 640 kbp; sample 0 carries three read pairs that span more than a segment (the cursor of getEndCount stays
    inside their end set, so the activeLoad counter of that read group stops seeing read pairs that close in later sets) and
    two 400x bursts in later segments that drive the stale counter to the cap; deletions inside and outside the bursts."""
    dels = [sim.Deletion(120_000, 1100, np.array([1, 2, 1])), sim.Deletion(356_000, 900, np.array([2, 1, 0])),
            sim.Deletion(470_000, 1500, np.array([1, 0, 2])), sim.Deletion(590_000, 700, np.array([1, 1, 1]))]
    samples, _ = sim.simulate_cohort(seed=19, n_samples=3, contig_len=640_000, n_dels=0, dels=dels)
    spec = sim.ReadGroupSpec(name="burst", coverage=400.0)
    extra, _ = sim.simulate_cohort(seed=20, n_samples=1, contig_len=640_000, n_dels=0, rg_specs=[[spec]])
    e = extra[0].read_groups[0]
    m = ((e.pos >= 350_000) & (e.pos < 362_000)) | ((e.pos >= 520_000) & (e.pos < 528_000))
    rg = samples[0].read_groups[0]
    far_pos = np.array([150_003, 150_950, 188_000], dtype=rg.pos.dtype)
    far_isz = np.array([262_000, 255_000, 420_000], dtype=rg.isize.dtype)
    p, i = np.concatenate([rg.pos, e.pos[m], far_pos]), np.concatenate([rg.isize, e.isize[m], far_isz])
    o = np.lexsort((i, p))
    rg.pos, rg.isize = p[o], i[o]
    return pf.write_cohort(d, samples, ("chr7", 640_000)), {}


CASES = dict(basic=case_basic, mixedrg=case_mixedrg, gap=case_gap, highcov=case_highcov,
             twocontigs=case_twocontigs, offset=case_offset, capfar=case_capfar)
WINDOW_DUMPS = {"basic", "gap", "highcov", "twocontigs", "mixedrg", "offset", "capfar"}


def run_case(name):
    d = os.path.join(HERE, name)
    shutil.rmtree(d, ignore_errors=True)
    paths, opts = CASES[name](d)
    args = opts.get("args", [])
    lst = os.path.join(d, "profiles.txt")
    with open(lst, "w") as fh:
        fh.write("\n".join(os.path.basename(p) for p in paths) + "\n")
    quiet = dict(stdout=subprocess.DEVNULL, cwd=d)
    subprocess.run([REF, "call", "profiles.txt", "-o", "merged.vcf"] + args, check=True, **quiet)
    subprocess.run([REF, "call", "profiles.txt", "-n", "-o", "win.vcf"] + args, check=True, **quiet)
    env = dict(os.environ, POPDEL_HARNESS_OUT="harness.dump")
    if name in WINDOW_DUMPS:
        env["POPDEL_HARNESS_WINDOWS"] = "1"
    subprocess.run([HARNESS, "profiles.txt", "-o", "harness.vcf"] + args, check=True, env=env, **quiet)
    os.remove(os.path.join(d, "harness.vcf"))
    with open(os.path.join(d, "args.txt"), "w") as fh:
        fh.write(" ".join(args) + "\n")
    for f in ("harness.dump", "win.vcf"):
        with open(os.path.join(d, f), "rb") as src, gzip.GzipFile(os.path.join(d, f + ".gz"), "wb", mtime=0) as dst:
            shutil.copyfileobj(src, dst)
        os.remove(os.path.join(d, f))
    n = sum(1 for line in open(os.path.join(d, "merged.vcf")) if not line.startswith("#"))
    size = sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d))
    print(f"{name}: {len(paths)} profiles, {n} merged calls, {size / 1e6:.2f} MB")


def run_big_case(name):
    """BASELINE-sized cases (tests/bigcases.py): the profiles live in a scratch directory and are regenerated by the tests;
    only the reference's outputs are committed."""
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bigcases
    d = os.path.join(HERE, name)
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    tmp = tempfile.mkdtemp(prefix="popdel_golden_")
    try:
        paths = bigcases.profile_list(name, tmp)
        quiet = dict(stdout=subprocess.DEVNULL, cwd=tmp)
        subprocess.run([REF, "call", "profiles.txt", "-o", "merged.vcf"], check=True, **quiet)
        subprocess.run([REF, "call", "profiles.txt", "-n", "-o", "win.vcf"], check=True, **quiet)
        subprocess.run([HARNESS, "profiles.txt", "-o", "harness.vcf"], check=True, env=dict(os.environ, POPDEL_HARNESS_OUT="harness.dump"), **quiet)
        shutil.copy(os.path.join(tmp, "merged.vcf"), os.path.join(d, "merged.vcf"))
        for f in ("harness.dump", "win.vcf"):
            with open(os.path.join(tmp, f), "rb") as src, gzip.GzipFile(os.path.join(d, f + ".gz"), "wb", mtime=0) as dst:
                shutil.copyfileobj(src, dst)
        with open(os.path.join(d, "args.txt"), "w") as fh:
            fh.write("\n")
        n = sum(1 for line in open(os.path.join(d, "merged.vcf")) if not line.startswith("#"))
        size = sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d))
        print(f"{name}: {len(paths)} profiles (not committed), {n} merged calls, {size / 1e6:.2f} MB")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_cap_replay():
    """tests/golden/cap_replay.npz: which read pairs of tests/capcases.py's streams the reference's own
    ChromosomeProfile::add stores (popdel_ref_harness in replay mode), bit-packed."""
    import tempfile
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import capcases
    out = {}
    with tempfile.TemporaryDirectory(prefix="popdel_golden_") as tmp:
        for seed in capcases.SEEDS:
            max_load, start, end = capcases.cap_stream(seed)
            inp, res = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
            with open(inp, "wb") as fh:
                fh.write(np.array([capcases.WINDOW_BUFFER, max_load], np.uint32).tobytes() + np.uint64(start.size).tobytes())
                fh.write(start.tobytes() + end.tobytes())
            subprocess.run([HARNESS], check=True, env=dict(os.environ, POPDEL_HARNESS_CAP_REPLAY=inp, POPDEL_HARNESS_OUT=res))
            stored = np.fromfile(res, dtype=np.uint8)
            assert stored.size == start.size
            out[f"n{seed}"] = np.array([start.size, int(stored.sum())], np.int64)
            out[f"stored{seed}"] = np.packbits(stored)
            print(f"cap_replay seed {seed}: max_load {max_load}, {start.size} read pairs, {int(stored.sum())} stored")
    np.savez_compressed(os.path.join(HERE, "cap_replay.npz"), **out)


BIG_CASES = ["big100", "mixed200"]

if __name__ == "__main__":
    for c in (sys.argv[1:] or list(CASES) + BIG_CASES + ["cap_replay"]):
        run_cap_replay() if c == "cap_replay" else run_big_case(c) if c in BIG_CASES else run_case(c)
