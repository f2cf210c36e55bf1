"""Pins the CPU oracle (oracle/popdel_oracle.cpp) against outputs of the UNMODIFIED reference
(tests/golden/*/harness.dump.gz, produced by tests/golden/make_golden.py with oracle/_ref).

Integer fields (window positions, active read-pair counts and checksums, candidate/final lengths, iterations,
PL/LAD/DAD/FL, filters, segment grouping, merged calls) must be identical; LR and allele frequency agree to
1e-12 relative (the reference sums in std::unordered_set iteration order, the oracle in position order).
"""
import os

import pytest

from dumpfmt import parse_dump

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["basic", "mixedrg", "gap", "highcov", "twocontigs", "offset"]
FLOAT_RTOL = 1e-12


def _files(case):
    d = os.path.join(GOLDEN, case)
    return [os.path.join(d, l.strip()) for l in open(os.path.join(d, "profiles.txt")) if l.strip()]


def _args(case):
    t = open(os.path.join(GOLDEN, case, "args.txt")).read().split()
    return dict(min_init_len=int(t[t.index("-l") + 1]) if "-l" in t else 0)


def compare_calls(got, ref, rtol=FLOAT_RTOL):
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        assert a["ints"] == b["ints"]
        assert a["samples"] == b["samples"]
        assert a["lr"] == pytest.approx(b["lr"], rel=rtol, abs=0)
        assert a["freq"] == pytest.approx(b["freq"], rel=rtol, abs=0)


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_dump(case, oracle_lib, tmp_path):
    ref = parse_dump(os.path.join(GOLDEN, case, "harness.dump.gz"))
    out = tmp_path / "oracle.dump"
    oracle_lib.call_files(_files(case), str(out), dump_windows=True, **_args(case))
    got = parse_dump(out)
    assert len(ref["windows"]) > 1000
    assert got["windows"] == ref["windows"]          # every scanned window: position, #calls, per-RG active sets
    assert len(got["segments"]) == len(ref["segments"])
    n_raw = 0
    for sg, sr in zip(got["segments"], ref["segments"]):
        compare_calls(sg["raw"], sr["raw"])
        compare_calls(sg["merged"], sr["merged"])
        n_raw += len(sr["raw"])
    assert n_raw > 50


@pytest.mark.parametrize("case", CASES)
def test_unify_hook_matches_reference_merge(case, oracle_lib):
    """orc_unify_segments -- the checker of the device-side unifyCalls (tests/test_scan_gpu.py) -- fed with the raw window
    calls the REFERENCE dumped per processSegment() must return the merged calls the reference dumped (positions, lengths,
    filters, PL / LAD / DAD / FL rows, significant windows identical; LR / AF 1e-12)."""
    import numpy as np
    from popdel_b200 import api
    from popdel_b200 import profile_format as pf
    ref = parse_dump(os.path.join(GOLDEN, case, "harness.dump.gz"))
    stddevs = [m["stddev"] for f in _files(case) for m in pf.read_profile(f)[0]]
    mean_sd = float(np.mean(stddevs))                                           # setMeanStddev: mean over all read groups
    raw, per = [], []
    for k, seg in enumerate(ref["segments"]):
        for c in seg["raw"]:
            i = c["ints"]
            raw.append((i[0], i[1], i[2], i[6], c["lr"], c["freq"], i[3], i[4], i[5], k))
            per.append(c["samples"])
    calls = np.array(raw, dtype=api.CALL_DTYPE)
    ps = np.array(per, dtype=np.uint32)
    got, got_ps, got_sig = oracle_lib.unify_segments(calls, ps, mean_sd, 0.5, False)
    want = [c for seg in ref["segments"] for c in seg["merged"]]
    assert len(got) == len(want) and len(want) > 0
    for g, gp, gs, w in zip(got, got_ps, got_sig, want):
        i = w["ints"]
        assert [int(g["initial_length"]), int(g["iterations"]), int(g["deletion_length"]), int(g["window_position"]), int(g["position"]),
                int(g["end_position"]), int(g["filter"]), gp.shape[0], int(gs)] == i
        assert gp.tolist() == w["samples"]
        assert float(g["lr"]) == pytest.approx(w["lr"], rel=FLOAT_RTOL, abs=0)
        assert float(g["frequency"]) == pytest.approx(w["freq"], rel=FLOAT_RTOL, abs=0)
