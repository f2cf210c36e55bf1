"""The reference's own `popdel call` translation unit with its window loop replaced by the scan library
(integration/popdel_call_gpu.cpp: unmodified reference headers + include/popdel_b200.h through the reference-side stub
integration/gpu_scan_popdel_call.h) must write the VCFs the unmodified reference wrote (tests/golden/*)."""
import os
import subprocess

import pytest

from test_cli_gpu import GOLDEN, _compare, _lines

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "integration", "_build", "popdel_call_gpu")


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["basic", "mixedrg", "gap", "highcov", "twocontigs", "offset", "capfar"])
def test_reference_tu_with_gpu_scan_writes_the_reference_vcfs(case, tmp_path):
    if not os.path.exists(BIN):
        pytest.skip("integration/_build/popdel_call_gpu not built (needs /root/reference at build time)")
    args = open(os.path.join(GOLDEN, case, "args.txt")).read().split()
    for extra, ref in ([], "merged.vcf"), (["-n"], "win.vcf.gz"):
        out = str(tmp_path / "out.vcf")
        subprocess.run([BIN, "profiles.txt", "-o", out] + args + extra, check=True, cwd=os.path.join(GOLDEN, case),
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        _compare(_lines(out), _lines(os.path.join(GOLDEN, case, ref)))
