#!/usr/bin/env python
"""Summarise `ncu --page source --csv --print-source sass` output: consecutive
SASS instructions with the same execution count are folded into one region."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
kern = 0
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Address":
        hdr = rows[i]; ix = {h: n for n, h in enumerate(hdr)}
        data = []
        i += 1
        while i < len(rows) and rows[i] and rows[i][0] not in ("Address", "Kernel Name"):
            if len(rows[i]) > 10:
                data.append(rows[i])
            i += 1
        kern += 1
        if len(sys.argv) > 2 and kern != int(sys.argv[2]):
            continue
        tot = sum(int(r[ix["Instructions Executed"]]) for r in data)
        tsm = sum(int(r[ix["# Samples"]]) for r in data)
        print(f"== kernel launch {kern}: {tot} warp instructions, {tsm} samples")
        out = [(n, r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]]),
                r[ix["Avg. Threads Executed"]]) for n, r in enumerate(data)]
        a = 0
        while a < len(out):
            b = a
            while b + 1 < len(out) and out[b + 1][2] == out[a][2]:
                b += 1
            ie = out[a][2]; cnt = b - a + 1; smp = sum(o[3] for o in out[a:b + 1])
            if cnt * ie * 200 > tot or smp * 200 > tsm:
                print(f"  sass {out[a][0]:4d}-{out[b][0]:4d}: {cnt:4d} x {ie:9d} = {100*cnt*ie/tot:5.1f}% instr, "
                      f"{100*smp/max(tsm,1):5.1f}% samples, thr {out[a][4]:>5s}  [{out[a][1][:34]} .. {out[b][1][:28]}]")
            a = b + 1
    else:
        i += 1
