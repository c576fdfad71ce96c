#!/usr/bin/env python
"""Phase timings of the batch pipeline for several host thread counts (GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import supersampler_b200 as S
from supersampler_b200 import synth

def main():
    S.build()
    n, bases = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 5_000_000
    fam = synth.Family(bases, 42)
    fas = [fam.fasta(i) for i in range(n)]
    tl = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4, 8, 16, 32]
    for t in tl:
        pl = S.Pipeline(31, 11, 1000, threads=t)
        for _ in range(3):
            pl.sketch(fas); pl.compare()
        acc = {}
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            info = {}
            pl.sketch(fas, info=info); pl.compare()
            for k_, v in info.items(): acc[k_] = acc.get(k_, 0) + v / reps
        dt = (time.perf_counter() - t0) / reps
        print(f"threads {t:3d}: step {dt*1e3:7.2f} ms | prep {acc['prep_s']*1e3:6.2f} pack {acc['pack_s']*1e3:6.2f} "
              f"device {acc['device_s']*1e3:6.2f} (scan {acc['scan_ms']:.3f} post {acc['post_ms']:.3f}) assemble {acc['assemble_s']*1e3:5.2f}",
              flush=True)
        pl.close()

if __name__ == "__main__":
    main()
