#!/usr/bin/env python
"""Per-source-line stall samples / executed instructions of one kernel in an ncu report (needs -lineinfo and
--import-source on):  python profiles/ncu_source_lines.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    lines, fname = [], ""
    tot_s = tot_i = 0
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        if len(r) > 7 and r[0].isdigit() and r[2] == "-":
            try:
                s, i = int(r[4] or 0), int(r[7] or 0)
            except ValueError:
                continue
            lines.append((s, i, fname, int(r[0]), r[1].strip()))
            tot_s += s
            tot_i += i
    print(f"total samples {tot_s}, warp instructions {tot_i}")
    for s, i, f, ln, src in sorted(lines, reverse=True)[:top]:
        print(f"{100.0 * s / max(tot_s, 1):5.1f}% smp {100.0 * i / max(tot_i, 1):5.1f}% ins  {f}:{ln:<4d} {src[:110]}")


if __name__ == "__main__":
    main()
