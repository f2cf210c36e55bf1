#!/usr/bin/env python
"""Drop-in CLI timing: popdel_b200_call versus the reference `popdel call` on the SAME gzip profile files.

    python scripts/cli_bench.py [--samples 100] [--length 4800000] [--no-reference] [--compare]

Writes synthetic profiles (bench.py's cohort generator) to a temporary directory, runs our host shell once cold and
once warm (PD_TIMING=1 prints the stage times), then the unmodified reference binary (oracle/_ref/popdel_ref) with one
process per core over contiguous -r regions (what bench.py --impl reference times) and, with --compare, once as a
single process so that the two VCFs can be compared record by record. Prints one JSON line."""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def records(path):
    return [l.rstrip("\n").split("\t") for l in open(path) if not l.startswith("#")]


def main():
    import bench
    from popdel_b200 import profile_format as pf
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=100)
    ap.add_argument("--length", type=int, default=4_800_000)
    ap.add_argument("--dels-per-mbp", type=float, default=2.0)
    ap.add_argument("--no-reference", action="store_true")
    ap.add_argument("--compare", action="store_true")
    args = ap.parse_args()
    cores = os.cpu_count() or 1
    N, L = args.samples, args.length
    cohort, _ = bench.make_cohort(1, N, L, args.dels_per_mbp, cores)
    tmp = tempfile.mkdtemp(prefix="popdel_cli_bench_")
    out = {"samples": N, "length_bp": L, "windows": L // 30, "cores": cores}
    try:
        contigs = [("chr21", bench.CHR21_LEN)]

        def write(s):
            p = os.path.join(tmp, f"s{s:05d}.profile")
            pf.write_profile_single_rg(p, cohort[s][3], contigs, 0, cohort[s][0], cohort[s][1], compressed=True)
            return p

        with ThreadPoolExecutor(cores) as ex:
            paths = list(ex.map(write, range(N)))
        out["profile_bytes"] = int(sum(os.path.getsize(p) for p in paths))
        lst = os.path.join(tmp, "profiles.txt")
        open(lst, "w").write("\n".join(paths) + "\n")
        ours = os.path.join(ROOT, "popdel_b200", "popdel_b200_call")
        env = dict(os.environ, PD_TIMING="1")
        runs = []
        for i in range(2):
            t0 = time.perf_counter()
            r = subprocess.run([ours, lst, "-o", os.path.join(tmp, "ours.vcf")], env=env, capture_output=True, text=True)
            dt = time.perf_counter() - t0
            assert r.returncode == 0, r.stderr
            stages = {l.split()[1]: float(l.split()[2]) for l in r.stderr.splitlines() if l.startswith("[popdel_b200]") and l.endswith(" s")}
            runs.append({"wall_s": dt, "stages_s": stages})
        out["ours"] = {"runs": runs, "evals_per_s": N * (L // 30) / runs[-1]["wall_s"], "records": len(records(os.path.join(tmp, "ours.vcf")))}
        ref = os.path.join(ROOT, "oracle", "_ref", "popdel_ref")
        if not args.no_reference and os.path.exists(ref):
            regions = np.linspace(0, L, cores + 1).astype(int)
            t0 = time.perf_counter()
            procs = [subprocess.Popen([ref, "call", lst, "-r", f"chr21:{regions[i] + 1}-{regions[i + 1]}", "-o", os.path.join(tmp, f"ref{i}.vcf")],
                                      stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for i in range(cores)]
            assert all(p.wait() == 0 for p in procs)
            dt = time.perf_counter() - t0
            out["reference"] = {"wall_s": dt, "processes": cores, "evals_per_s": N * (L // 30) / dt}
            out["speedup_same_files"] = dt / runs[-1]["wall_s"]
            if args.compare:
                t0 = time.perf_counter()
                subprocess.run([ref, "call", lst, "-o", os.path.join(tmp, "ref.vcf")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
                out["reference_single"] = {"wall_s": time.perf_counter() - t0}
                a, b = records(os.path.join(tmp, "ours.vcf")), records(os.path.join(tmp, "ref.vcf"))
                same = len(a) == len(b) and all(x[:7] == y[:7] and x[8:] == y[8:] for x, y in zip(a, b))
                out["vcf_identical_except_info_float_digits"] = bool(same)
                out["reference_single"]["records"] = len(b)
        print(json.dumps(out))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
