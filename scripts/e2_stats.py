#!/usr/bin/env python
"""Clock statistics of k_e2_reads (debug build -DPD_EM_STATS): POPDEL_B200_LIB=build/libpopdel_stats.so python scripts/e2_stats.py"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from popdel_b200 import api
cohort, dels = bench.make_cohort(1, 100, bench.CHR21_LEN, 2.0, os.cpu_count() or 8, False)
params = api.CallParameters()
rgs = api.read_groups_from_headers([[c[3] for c in cohort if c[4] == s] for s in range(100)], params)
sc = api.Scanner(params, rgs, 100, device=0)
sc.begin_contig((min(int(c[0][0]) for c in cohort) // 30) * 30)
for g, c in enumerate(cohort):
    sc.push(g, c[0], c[2])
sc.upload()
lib = api.load_library()
st = (C.c_ulonglong * 32)()
res = sc.scan(copy=False); lib.pd_debug_e2_stats(st)
res = sc.scan(copy=False); lib.pd_debug_e2_stats(st)
v = list(st)
names = ["blocks", "clk_stage", "warpblocks", "clk_wb_loads", "sum_jmax", "clk_loop", "clk_block_total", ""]
print(json.dumps({("A_" if i < 8 else "M_") + names[i % 8]: v[i] for i in range(16) if v[i]}, indent=1), float(res["ms_em"]))
