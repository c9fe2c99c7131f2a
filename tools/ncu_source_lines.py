#!/usr/bin/env python
"""Per-source-line totals (samples, instructions executed) of one kernel from an ncu report.
usage: ncu_source_lines.py REPORT.ncu-rep KERNEL_REGEX [TOP]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
lines = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 8:
        continue
    if r[0]:      # a source line summary row
        try:
            lines.append((int(r[0]), r[1].strip(), int(r[4] or 0), int(r[7] or 0)))
        except ValueError:
            pass
tot_s = sum(l[2] for l in lines) or 1
tot_i = sum(l[3] for l in lines) or 1
print("total samples %d, instructions %d" % (tot_s, tot_i))
for ln, src, s, i in sorted(lines, key=lambda l: -l[2])[:top]:
    print("%5d  samples %5.1f%%  inst %5.1f%%  %s" % (ln, 100.0 * s / tot_s, 100.0 * i / tot_i, src[:110]))
