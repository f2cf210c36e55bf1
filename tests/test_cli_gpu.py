"""End-to-end drop-in check: the host shell `popdel_b200_call` (C++ loader + CUDA scan + merge + VCF writer) against
the VCFs written by the unmodified reference on the same profile files (tests/golden/*/merged.vcf, win.vcf.gz).
All text must be identical except the file date; AF and LR (printed with 6 significant digits) may differ by one unit
in the last printed digit."""
import gzip
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
CLI = os.path.join(ROOT, "popdel_b200", "popdel_b200_call")
CASES = ["basic", "mixedrg", "gap", "highcov", "twocontigs", "offset"]


def _lines(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as fh:
        return [l.rstrip("\n") for l in fh if not l.startswith("##fileDate")]


def _compare(got, ref):
    assert len(got) == len(ref), f"{len(got)} lines, reference has {len(ref)}"
    for a, b in zip(got, ref):
        if a == b:
            continue
        fa, fb = a.split("\t"), b.split("\t")
        assert len(fa) == len(fb)
        for k, (x, y) in enumerate(zip(fa, fb)):
            if x == y:
                continue
            assert k == 7, f"column {k} differs: {x} vs {y}"
            ia, ib = x.split(";"), y.split(";")
            assert len(ia) == len(ib)
            for u, v in zip(ia, ib):
                if u == v:
                    continue
                ku, vu = u.split("=")
                kv, vv = v.split("=")
                assert ku == kv and ku in ("AF", "LR"), f"{u} vs {v}"
                assert abs(float(vu) - float(vv)) <= 2e-6 * abs(float(vv)) + 1e-12, f"{u} vs {v}"


def _run(case, extra, out):
    args = open(os.path.join(GOLDEN, case, "args.txt")).read().split()
    subprocess.run([CLI, "profiles.txt", "-o", out] + args + extra, check=True, cwd=os.path.join(GOLDEN, case),
                   stdout=subprocess.DEVNULL)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_merged_vcf_equals_reference(case, tmp_path):
    out = str(tmp_path / "merged.vcf")
    _run(case, [], out)
    _compare(_lines(out), _lines(os.path.join(GOLDEN, case, "merged.vcf")))


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_window_wise_vcf_equals_reference(case, tmp_path):
    out = str(tmp_path / "win.vcf")
    _run(case, ["-n"], out)
    _compare(_lines(out), _lines(os.path.join(GOLDEN, case, "win.vcf.gz")))


@pytest.mark.gpu
@pytest.mark.parametrize("devs", ["0,0", "0,0,0"])
@pytest.mark.parametrize("case", ["basic", "gap", "twocontigs", "highcov"])
def test_window_range_sharding_equals_reference(case, devs, tmp_path):
    """-g D0,D1,...: one scan context and host thread per entry (here all on GPU 0), each scanning a contiguous range of
    whole segments of every contig; the VCFs must still be the reference's (workflow_popdel.h:297-366 merge order)."""
    out = str(tmp_path / "merged.vcf")
    _run(case, ["-g", devs], out)
    _compare(_lines(out), _lines(os.path.join(GOLDEN, case, "merged.vcf")))
    out = str(tmp_path / "win.vcf")
    _run(case, ["-n", "-g", devs], out)
    _compare(_lines(out), _lines(os.path.join(GOLDEN, case, "win.vcf.gz")))


def test_cli_fails_loudly_without_gpu(tmp_path):
    """No CPU fallback: on a box without a CUDA device the tool must stop with an error."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    r = subprocess.run([CLI, "profiles.txt", "-o", str(tmp_path / "x.vcf")], cwd=os.path.join(GOLDEN, "basic"),
                       capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU scan path" in r.stderr


def _option_variants():
    import json
    return json.load(open(os.path.join(GOLDEN, "options.json")))


@pytest.mark.gpu
@pytest.mark.parametrize("v", _option_variants(), ids=lambda v: f"{v['case']}-{v['name']}")
def test_option_variants_equal_reference(v, tmp_path):
    """-A (per-read-group coverage caps), -e (conflicting read-group IDs), -F / -c, -u / -t / -p / -s / -f: the VCF equals the
    one the unmodified reference wrote with the same options (tests/golden/make_golden_options.py)."""
    out = str(tmp_path / "out.vcf")
    subprocess.run([CLI, v["list"], "-o", out] + v["args"], check=True, cwd=os.path.join(GOLDEN, v["case"]), stdout=subprocess.DEVNULL)
    ref = _lines(os.path.join(GOLDEN, v["case"], v["vcf"]))
    assert sum(1 for l in ref if not l.startswith("#")) == v["records"]
    _compare(_lines(out), ref)
