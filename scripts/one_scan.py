#!/usr/bin/env python
"""Runs a few resident scans of the bench cohort (for ncu captures):  python scripts/one_scan.py [--scans K] [--length L] [--mixed]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from popdel_b200 import api

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=100)
ap.add_argument("--length", type=int, default=bench.CHR21_LEN)
ap.add_argument("--scans", type=int, default=2)
ap.add_argument("--mixed", action="store_true")
ap.add_argument("--unify", action="store_true")
a = ap.parse_args()
cohort, dels = bench.make_cohort(1, a.samples, a.length, 2.0, os.cpu_count() or 8, a.mixed)
params = api.CallParameters()
rgs = api.read_groups_from_headers([[c[3] for c in cohort if c[4] == s] for s in range(a.samples)], params)
sc = api.Scanner(params, rgs, a.samples, device=0)
anchor = (min(int(c[0][0]) for c in cohort) // 30) * 30
sc.begin_contig(anchor)
for g, c in enumerate(cohort):
    sc.push(g, c[0], c[2])
sc.upload()
if a.unify:
    import numpy as np
    sc.set_unify(float(np.mean([r.stddev for r in rgs])), 0.5, False)
for _ in range(a.scans):
    res = sc.scan(copy=False)
    print({k: (float(v) if isinstance(v, float) else v) for k, v in res.items() if k.startswith("ms_") or k.startswith("n_")}, len(res["calls"]))
