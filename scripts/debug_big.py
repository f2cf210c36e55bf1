import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bigcases, oracle_api
from parity import run_oracle, INT_FIELDS
from popdel_b200 import api
orc = oracle_api.load(os.path.join(ROOT, "oracle", "liboracle.so"))
samples = bigcases.mixed200_samples()
params = api.CallParameters()
res, rgs = api.scan_cohort(samples, params)
ref_calls, ref_ps, nwin = run_oracle(samples, params, rgs, orc)
print(len(res["calls"]), len(ref_calls), res["n_windows"], nwin)
n = min(len(res["calls"]), len(ref_calls))
bad = 0
for i in range(n):
    a, b = res["calls"][i], ref_calls[i]
    d = [f for f in INT_FIELDS if a[f] != b[f]]
    pd = np.argwhere(res["per_sample"][i] != ref_ps[i])
    if d or len(pd):
        print(i, "win", b["window_position"], "L0", b["initial_length"], d, "ps diffs", len(pd), pd[:4].tolist(),
              [(res["per_sample"][i][s].tolist(), ref_ps[i][s].tolist()) for s in sorted(set(pd[:, 0].tolist()))[:2]])
        bad += 1
        if bad > 6: break
print("bad", bad)
