#!/usr/bin/env python
"""Phase statistics of the fused EM kernel (debug build with -DPD_EM_STATS, see popdel_b200/csrc/pd_em.cu):
   POPDEL_B200_LIB=build/libpopdel_stats.so python scripts/em_stats.py [--length L]"""
import argparse, ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from popdel_b200 import api

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=100)
ap.add_argument("--length", type=int, default=bench.CHR21_LEN)
ap.add_argument("--mixed", action="store_true")
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
res = sc.scan(copy=False)
lib = api.load_library()
st = (C.c_ulonglong * 32)()
assert lib.pd_debug_em_stats(st) == 0
res = sc.scan(copy=False)
assert lib.pd_debug_em_stats(st) == 0
v = list(st)
names = {0: "blocks", 1: "clk_total(accepted)", 2: "clk_sort", 3: "clk_init", 4: "clk_dl", 5: "dl_calls", 6: "clk_final_pass", 7: "clk_percentile",
         9: "sum_iterations", 10: "warp_passes", 11: "warp_pass_iters(max nl)", 12: "lane_iters(sum nl)", 13: "final_passes",
         14: "warp_passes_mode0", 15: "warp_passes_mode1", 16: "warp_passes_mode2", 17: "lane_iters_mode0", 18: "lane_iters_mode1",
         19: "lane_iters_mode2", 20: "clk_em_loop", 21: "sum_supp"}
out = {names.get(i, str(i)): v[i] for i in range(32) if v[i]}
out["pairs"] = int(res["n_candidates"]); out["calls"] = int(len(res["calls"])); out["ms_em"] = float(res["ms_em"]); out["ms_total"] = float(res["ms_total"])
print(json.dumps(out, indent=1))
