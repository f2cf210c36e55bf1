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
