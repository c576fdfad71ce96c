#!/usr/bin/env python
"""A few batches through the pipeline with the text cleaned + packed on the device (ingest kernels), for ncu and
for a kernel-time figure: 64 x 5 Mbp genomes (80-column FASTA) and one read set of 150 bp reads."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import supersampler_b200 as S
from supersampler_b200 import synth


def main():
    S.build()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    fam = synth.Family(5_000_000, 42)
    fas = [S.PinnedBuffer(fam.fasta(i)) for i in range(n)]
    g = synth.random_genome(5_000_000, 5)
    reads = S.PinnedBuffer(synth.reads_fasta_bytes(synth.read_set(400_000, 150, g, 6)))
    pl = S.Pipeline(31, 11, 1000.0, threads=8, ingest="device")
    for name, job in (("genomes", fas), ("reads", [reads])):
        text = sum(len(x) for x in job)
        for _ in range(steps):
            info = {}
            t0 = time.perf_counter()
            pl.sketch(job, info=info)
            dt = time.perf_counter() - t0
        print(f"{name}: {text / 1e6:.0f} MB of text, ingest kernels {info['ingest_ms']:.3f} ms = {text / info['ingest_ms'] / 1e6:.0f} GB/s of text, "
              f"scan {info['scan_ms']:.3f} ms, post-pass {info['post_ms']:.3f} ms, call {dt * 1e3:.2f} ms ({text / dt / 1e9:.1f} GB/s of text end to end)")
    pl.close()


if __name__ == "__main__":
    main()
