"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        if d.get("Metric Unit", "ns") in ("us", "usecond"):
            v *= 1e3
        k = d["Kernel Name"][:64]
        agg[k][0] += 1
        agg[k][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:64s} {v[0]:6d} launches {v[1] / 1e6:10.3f} ms {100 * v[1] / tot:5.1f}%")
