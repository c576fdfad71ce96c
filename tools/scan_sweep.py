#!/usr/bin/env python
"""Kernel-only timing of the scan on a device-resident batch (CUDA events on
the launching stream, inputs rotated over > L2).  Used to compare kernel
variants; prints one JSON line per configuration."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import supersampler_b200 as S
from supersampler_b200 import synth

def run(k, m, s, mode, n_bases=320_000_000, reps=20):
    thr = S.threshold(k, m, s)
    rng = np.random.default_rng(1)
    words = rng.integers(0, 2**32, size=n_bases // 16 + 64, dtype=np.uint32)
    nrep = 3
    d = [torch.from_numpy(words.view(np.int32)).cuda() for _ in range(nrep)]
    cap = int(n_bases * thr / 2.0**64 * 1.5) + 65536
    d_hits = torch.empty(cap * 16, dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    ctx = S.DeviceContext(k, m, thr)
    ctx.config(mode)
    finfo = ctx.filter_info() if mode != 1 else {}
    ms = []
    for i in range(reps + 3):
        ctx.scan_device(d[i % nrep].data_ptr(), n_bases, d_hits.data_ptr(), cap, d_cnt.data_ptr())
        ctx.sync()
        if i >= 3:
            ms.append(ctx.scan_kernel_ms())
    n = int(d_cnt.item())
    t = float(np.median(ms))
    gbs = (n_bases / 4 + 16 * n) / (t * 1e-3) / 1e9
    print(json.dumps({"k": k, "m": m, "s": s, "mode": mode, "G": os.environ.get("SPSP_FILTER_G"),
                      "rep": os.environ.get("SPSP_FILTER_REP"), "ms": round(t, 4), "min_ms": round(min(ms), 4),
                      "tbp_s": round(n_bases / (t * 1e-3) / 1e12, 3), "gb_s": round(gbs, 1),
                      "frac_hbm": round(gbs / 6534.8, 4), "hits": n, "n_bases": n_bases,
                      "kind": os.environ.get("SPSP_FILTER_KIND"), "filter": finfo}), flush=True)
    ctx.close()

if __name__ == "__main__":
    k, m, s = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
    mode = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    nb = int(float(sys.argv[5])) if len(sys.argv) > 5 else 320_000_000
    run(k, m, s, mode, n_bases=nb)
