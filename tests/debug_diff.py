"""Debug helper (GPU box): prints the differences between the CUDA scan and the oracle for one cohort kind."""
import sys
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import oracle_api
from popdel_b200 import api
from parity import run_oracle
from test_host_logic import _cohort

kind = sys.argv[1] if len(sys.argv) > 1 else "longspan"
samples, params = _cohort(kind)
orc = oracle_api.load("oracle/liboracle.so")
res, rgs = api.scan_cohort(samples, params)
ref_calls, ref_ps, nwin = run_oracle(samples, params, rgs, orc)
g = {(int(c["window_position"]), int(c["initial_length"])): c for c in res["calls"]}
r = {(int(c["window_position"]), int(c["initial_length"])): c for c in ref_calls}
print("gpu", len(g), "oracle", len(r), "flagged", res["n_flagged_windows"], "cands", res["n_candidates"], "windows", res["n_windows"], nwin)
for k in sorted(set(r) - set(g)):
    c = r[k]
    print("MISSING on gpu", k, "win", (k[0] + 1 - api.cohort_anchor(samples)) // 30, dict(it=int(c["iterations"]), len=int(c["deletion_length"]), lr=float(c["lr"]), f=float(c["frequency"]), pos=int(c["position"]), end=int(c["end_position"])))
for k in sorted(set(g) - set(r)):
    c = g[k]
    print("EXTRA on gpu", k, dict(it=int(c["iterations"]), len=int(c["deletion_length"]), lr=float(c["lr"]), f=float(c["frequency"])))
for k in sorted(set(g) & set(r)):
    a, b = g[k], r[k]
    for f in ["iterations", "deletion_length", "filter", "position", "end_position", "segment"]:
        if a[f] != b[f]:
            print("DIFF", k, f, a[f], b[f])
    if abs(a["lr"] - b["lr"]) > 1e-6 * abs(b["lr"]) or abs(a["frequency"] - b["frequency"]) > 1e-6 * abs(b["frequency"]):
        print("DIFF float", k, a["lr"], b["lr"], a["frequency"], b["frequency"])
