#!/usr/bin/env python
"""One-line view of the EM v2 kernels of the LAST scan in an ncu launch list: python scripts/launch_summary.py launches.csv [nscans]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors='replace')))
nscans = int(sys.argv[2]) if len(sys.argv) > 2 else 2
hdr = None; out = []
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try: v = float(d['Metric Value'].replace(',', ''))
        except ValueError: continue
        out.append((d['Kernel Name'], v))
n = len(out) // nscans
last = out[-n:]
print(' '.join(f"{k.split('::')[-1].split('(')[0][:16]}:{v/1000:.0f}" for k, v in last if 'e2' in k))
print('e2 total us', sum(v for k, v in last if 'e2' in k) / 1e3, ' all us', sum(v for k, v in last) / 1e3)
