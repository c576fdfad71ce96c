#!/usr/bin/env python
"""The sketch half of bench.py's C3 extra for a few batches (device-synthesised family, s=100): per-batch
scan / post-pass times next to the wall time, and a target for ncu launch lists."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import supersampler_b200 as S
from supersampler_b200 import synth_device as SD
import bench

def main():
    S.build()
    nbatch = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    bsz = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    k, m, s, nb = 31, 11, 100.0, 5_000_000
    fam = SD.DeviceFamily(nb, seed=4242)
    ctx = S.DeviceContext(k, m, S.threshold(k, m, s))
    rs = bench.ResidentSet(ctx, k, m, s)
    b0 = fam.packed_batch(0, bsz)
    ctx.sketch_batch(None, *b0[1:], bsz, s, device_ptr=b0[0].data_ptr())
    del b0
    for i in range(nbatch):
        buf, n_total, rb, re_, ri = fam.packed_batch(i * bsz, bsz)
        torch.cuda.synchronize()
        p0, s0, w0 = rs.post_ms, rs.scan_ms, rs.sketch_s
        rs.add_batch(buf, n_total, rb, re_, ri, bsz)
        print(f"batch {i}: scan {rs.scan_ms - s0:.3f} ms post {rs.post_ms - p0:.3f} ms wall {(rs.sketch_s - w0) * 1e3:.3f} ms hits {rs.hits}")
        del buf

if __name__ == "__main__":
    main()
