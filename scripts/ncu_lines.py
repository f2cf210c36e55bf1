#!/usr/bin/env python
"""Per-source-line summary of one kernel of an ncu report:  python scripts/ncu_lines.py report.ncu-rep kernel_regex [launch_skip] [top]
(ncu --page source --print-source cuda,sass --csv; the kernel must have been compiled with -lineinfo and captured with --import-source on)"""
import csv, subprocess, sys, io, os
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + rx,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
cur, lines = None, {}
for row in csv.reader(io.StringIO(out)):
    if not row: continue
    if row[0] == "File Path": cur = os.path.basename(row[1]); continue
    if row[0] in ("Function Name",): continue
    if row[0] == "Line No": hdr = row; si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed"); ti = hdr.index("Thread Instructions Executed"); continue
    if row[2] != "-": continue                                  # SASS rows repeat the counts of their source line
    try: s, i, t = int(row[si]), int(row[ii]), int(row[ti])
    except ValueError: continue
    k = (cur, int(row[0]))
    a = lines.setdefault(k, [0, 0, 0, row[1].strip()])
    a[0] += s; a[1] += i; a[2] += t
S = sum(a[0] for a in lines.values()); I = sum(a[1] for a in lines.values())
print(f"total samples {S}, warp instructions {I}")
for k, a in sorted(lines.items(), key=lambda kv: (-kv[1][0] if os.environ.get("BY_SAMPLES") else -kv[1][1]))[:top]:
    print(f"{100*a[0]/max(S,1):5.1f}% smp {100*a[1]/max(I,1):5.1f}% ins  thr/ins {a[2]/max(a[1],1):4.1f}  {k[0]}:{k[1]}  {a[3][:110]}")
