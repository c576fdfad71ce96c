#!/usr/bin/env python
"""Stall samples and executed instructions per CUDA source line of one kernel in a .ncu-rep
(ncu -i rep --page source --csv --print-source cuda,sass): the line rows carry the sums of their SASS rows."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = None; out = []
for r in rows:
    if len(r) > 6 and r[0] == "Line No":
        hdr = r; isamp = r.index("# Samples"); iinst = r.index("Instructions Executed"); continue
    if hdr and len(r) > isamp and r[0].isdigit() and r[2] == "-":      # a source-line row (address column is '-')
        try:
            out.append((int(r[isamp] or 0), int(r[iinst] or 0), int(r[0]), r[1].strip()))
        except ValueError:
            pass
tot = sum(o[0] for o in out) or 1; toti = sum(o[1] for o in out) or 1
print(f"total samples {tot}, warp instructions {toti}")
for s, i, ln, src in sorted(out, key=lambda o: -o[0])[:top]:
    print(f"{s:7d} {100 * s / tot:5.1f}%  inst {100 * i / toti:5.1f}%  L{ln:<5d} {src[:110]}")
