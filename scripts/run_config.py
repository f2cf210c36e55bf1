#!/usr/bin/env python
"""BASELINE.json configs[2] at its stated size: 1 000 synthetic profiles with mixed read groups x chr1 (248 956 422 bp,
8.3e6 windows, 2.5e10 read pairs), scanned in window-range batches on the GPUs of one box.

    python scripts/run_config.py [--samples 1000] [--length 248956422] [--batch-mbp 6]            (1 GPU)
    python -m torch.distributed.run --nproc-per-node N ... scripts/run_config.py ...               (N GPUs: batches dealt round-robin)

The packed read pairs of the whole contig (100 GB) neither fit one scan batch (2^32 words) nor the host, so the contig is
processed the way the window-range sharding does it: every batch is a run of whole 600-kbp blocks (= 3 segments, so that
both the segment borders anchor + k * 200 000 and the 30-bp window grid of the batch coincide with the contig's), generated
by the counter-based generator (on the device: libpdsynth_cuda.so, or with --host-gen libpdsynth.so; any range of any read
group is reproducible), preceded by a 600-kbp
halo whose windows are scanned but not counted. Per batch: generate -> pd_contig_push_device -> upload -> pd_contig_scan (merged
output). Reports evaluations / device time (resident-style) and evaluations / (upload + scan) wall time."""
import argparse, json, os, sys, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from popdel_b200 import api

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=1000)
ap.add_argument("--length", type=int, default=248_956_422)
ap.add_argument("--batch-mbp", type=float, default=6.0)
ap.add_argument("--max-batches", type=int, default=0)
ap.add_argument("--single-rg", action="store_true")
ap.add_argument("--host-gen", action="store_true", help="generate on the host (libpdsynth.so) and push through PCIe instead of the device generator")
ap.add_argument("--first-batch", type=int, default=0)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dist = None
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
threads = max(1, (os.cpu_count() or 8) // world)
N, L, mixed = a.samples, a.length, not a.single_rg
from popdel_b200 import sharding
BLOCK = sharding.BATCH_BLOCK
blocks_per_batch = max(1, int(a.batch_mbp * 1e6) // BLOCK)
ds, dl, gt = bench.plant(1, N, L, 2.0)
specs = bench.cohort_specs(1, N, mixed)
R = len(specs)
HDR_LEN = min(L, 3_000_000 if R <= 3000 else 1_000_000)
# read-group headers from a 3-Mbp sample of every read group (the histograms of a profile are a property of the library)
near = ds < HDR_LEN + 40_000                                 # (the host generator walks every deletion per read pair)
def header(spec):
    s, g, mu, sd, dens = spec
    pos, isz = api.synth_read_group(1, g, mu, sd, 150, dens, 0, HDR_LEN, ds[near], dl[near], gt[near, s])
    med = int(mu); lo, hi = max(1, int(np.floor(med - 3 * sd))), int(np.ceil(med + 3 * sd)) + 1
    sel = isz[(isz >= lo) & (isz < hi)]
    return dict(name=f"rg{g}", median=med, stddev=sd, read_length=150, hist_start=lo, hist_end=hi,
                hist_counts=np.bincount(sel - lo, minlength=hi - lo).astype(np.float64)), s
with ThreadPoolExecutor(threads) as ex:
    hdrs = list(ex.map(header, specs))
params = api.CallParameters()
rgs = api.read_groups_from_headers([[h for h, s in hdrs if s == smp] for smp in range(N)], params)
sc = api.Scanner(params, rgs, N, device=local)
sc.set_unify(float(np.mean([r.stddev for r in rgs])), 0.5, False)
batches = sharding.contig_batches(L, blocks_per_batch * BLOCK)[a.first_batch:]          # (start, end, first owned window, owned windows)
if a.max_batches:
    batches = batches[:a.max_batches]
gen_dev = None if a.host_gen else api.SynthDevice(local)
mine = batches[rank::world]
tot = dict(evals=0, windows=0, reads=0, ms_dev=0.0, s_wall=0.0, s_gen=0.0, calls=0, variants=0, flagged=0, screened=0)
for (start, end, w0, n_own) in mine:
    t0 = time.time()
    if gen_dev is not None:                                 # device generator -> pd_contig_push_device: nothing crosses PCIe
        n_reads, dp, dd, rg_start = gen_dev.generate(1, specs, start, end, ds, dl, gt, N)
        tot["s_gen"] += time.time() - t0
        t0 = time.time()
        sc.begin_contig(start)                              # start is a multiple of 600 000: same window grid and segment borders
        for g in range(R):
            sc.push_device(g, int(rg_start[g + 1] - rg_start[g]), dp + 4 * int(rg_start[g]), dd + 4 * int(rg_start[g]))
    else:
        def gen(spec):
            s, g, mu, sd, dens = spec
            pos, isz = api.synth_read_group(1, g, mu, sd, 150, dens, start, end, ds, dl, gt[:, s])
            return pos, (isz - np.int32(int(mu))).astype(np.int32)
        with ThreadPoolExecutor(threads) as ex:
            data = list(ex.map(gen, specs))
        n_reads = sum(d[0].size for d in data)
        tot["s_gen"] += time.time() - t0
        t0 = time.time()
        sc.begin_contig(start)
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda g: sc.push(g, data[g][0], data[g][1]), range(R)))
        del data
    sc.upload()
    t_up = time.time() - t0
    nw = min(sc.window_count(), w0 + n_own)                 # the halo's windows are not counted; windows at or after `end` belong to the next batch
    res = sc.scan(first_window=w0, n_windows=max(nw - w0, 0), copy=False)
    tot["s_wall"] += time.time() - t0
    tot["windows"] += int(res["n_windows"]); tot["evals"] += int(res["n_windows"]) * N; tot["reads"] += n_reads
    tot["ms_dev"] += float(res["ms_total"]); tot["calls"] += int(res["n_window_calls"]); tot["variants"] += len(res["calls"])
    tot["flagged"] += int(res["n_flagged_windows"]); tot["screened"] += int(res["n_screened_windows"])
    if rank == 0:
        print(f"[run_config] batch {start}-{end}: {n_reads} read pairs, {res['n_windows']} windows, scan {float(res['ms_total']):.1f} ms on the device "
              f"(EM {float(res['ms_em']):.1f}), push + upload {t_up:.2f} s, generation {tot['s_gen']:.1f} s so far", file=sys.stderr, flush=True)
if dist is not None:
    import torch
    keys = ["evals", "windows", "reads", "calls", "variants", "flagged", "screened"]
    t = torch.tensor([float(tot[k]) for k in keys], dtype=torch.float64, device="cuda")
    m = torch.tensor([tot["ms_dev"], tot["s_wall"], tot["s_gen"]], dtype=torch.float64, device="cuda")
    dist.all_reduce(t); dist.all_reduce(m, op=dist.ReduceOp.MAX)
    for k, v in zip(keys, t.tolist()): tot[k] = int(v)
    tot["ms_dev"], tot["s_wall"], tot["s_gen"] = m.tolist()
    dist.destroy_process_group()
if rank == 0:
    print(json.dumps({"config": f"{N} synthetic profiles ({R} read groups, {'mixed' if mixed else 'single'} insert-size histograms) x {L} bp "
                                f"in window-range batches of {blocks_per_batch * BLOCK} bp + 600-kbp halo, {len(batches)} batches over {world} GPU(s)",
                      "evals": tot["evals"], "windows": tot["windows"], "read_pairs": tot["reads"],
                      "evals_per_s_device": tot["evals"] / (tot["ms_dev"] * 1e-3), "evals_per_s_upload_plus_scan": tot["evals"] / tot["s_wall"],
                      "device_s": tot["ms_dev"] * 1e-3, "upload_plus_scan_s": tot["s_wall"], "generation_s": tot["s_gen"],
                      "generator": "host (libpdsynth.so, pd_contig_push)" if a.host_gen else "device (libpdsynth_cuda.so, pd_contig_push_device)",
                      "window_calls": tot["calls"], "variants": tot["variants"], "flagged_windows": tot["flagged"],
                      "windows_after_second_screen_stage": tot["screened"], "n_gpus": world}))
