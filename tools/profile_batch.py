#!/usr/bin/env python
"""A few device-resident batch steps (scan + device post-pass + compare) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import supersampler_b200 as S
from supersampler_b200 import synth

def main():
    S.build()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    k, s = 31, float(sys.argv[2]) if len(sys.argv) > 2 else 1000.0
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    m = int(sys.argv[4]) if len(sys.argv) > 4 else 11
    fam = synth.Family(5_000_000, 42)
    ws, ros = [], []
    for i in range(n):
        w, nb, ro = S.pack_fasta(fam.fasta(i), k)
        ws.append(w); ros.append(ro)
    packed, n_total, rb, re_, ri = S.batch_layout(ws, ros)
    d = torch.from_numpy(packed.view(np.int32)).cuda()
    ctx = S.DeviceContext(k, m, S.threshold(k, m, s))
    for _ in range(steps):
        info = {}
        ctx.sketch_batch(None, n_total, rb, re_, ri, n, s, device_ptr=d.data_ptr(), info=info)
        ctx.cmp_load_batch()
        ctx.cmp_run((0, n), (0, n), True)
    dinfo = {}
    tot, sel = ctx.dense_stats(None, n_total, rb, re_, ri, n, device_ptr=d.data_ptr(), info=dinfo)
    tot, sel = ctx.dense_stats(None, n_total, rb, re_, ri, n, device_ptr=d.data_ptr(), info=dinfo)
    print("dense_ms", dinfo["dense_ms"], "total_superkmers[0]", int(tot[0]), "selected[0]", int(sel[0]))
    print("scan_ms", info["scan_ms"], "post_ms", info["post_ms"], "hits", info["n_hits"], "elems", info["n_elems"])

if __name__ == "__main__":
    main()
