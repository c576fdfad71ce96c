#!/usr/bin/env python
"""All-vs-all compare at C3 shape (N genomes of 5 Mbp, s=100) on one GPU: sketch through the batch pipeline,
compare from host element arrays (N x N tiled grid), pairs/s; the reference comparator (oracle/_ref,
single-threaded by construction) on the first M sketches for scale.
    python tools/compare_probe.py [N=256] [s=100] [M=128]
"""
import gzip, os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import supersampler_b200 as S
from supersampler_b200 import synth


def main():
    S.build()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    s = float(sys.argv[2]) if len(sys.argv) > 2 else 100.0
    m_ref = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    k, m = 31, 11
    fam = synth.Family(5_000_000, 42)
    pl = S.Pipeline(k, m, s, threads=16)
    sks = []
    t0 = time.perf_counter()
    for b in range(0, n, 64):
        sks += pl.sketch([fam.fasta(i) for i in range(b, min(n, b + 64))])
    t_sk = time.perf_counter() - t0
    el = [S.decode_sketch(x) for x in sks]
    off = np.concatenate([[0], np.cumsum([e[2].size for e in el])]).astype(np.uint64)
    mn = np.concatenate([e[2] for e in el]); lo = np.concatenate([e[3] for e in el])
    ctx = pl.device_context()
    ctx.cmp_load(off, mn, lo)
    for _ in range(2):
        ctx.cmp_run((0, n), (0, n), True)
    reps = 3
    t0 = time.perf_counter()
    kms = []
    for _ in range(reps):
        inter = ctx.cmp_run((0, n), (0, n), True)
        kms.append(ctx.cmp_kernel_ms())
    dt = (time.perf_counter() - t0) / reps
    pairs = n * (n - 1) // 2
    keycmp = float(sum(int(off[i + 1] - off[i]) for i in range(n))) * (n - 1)      # sum over pairs of |Ki| + |Kj|
    print(f"N={n} s={s:g}: {int(off[-1])} elements ({int(off[-1]) / n:.0f} per sketch), sketching {t_sk:.2f} s (incl. synth)")
    print(f"compare: kernel {np.mean(kms):.3f} ms, call {dt * 1e3:.3f} ms -> {pairs / (np.mean(kms) * 1e-3):.3e} pairs/s kernel, "
          f"{pairs / dt:.3e} pairs/s call; {keycmp / (np.mean(kms) * 1e-3):.3e} key comparisons/s, "
          f"{8 * keycmp / (np.mean(kms) * 1e-3) / 1e9:.1f} GB/s of merge-equivalent traffic")
    from oracle import oracle as O
    if O.have_ref() and m_ref > 1:
        wd = tempfile.mkdtemp(prefix="cmp_probe_")
        names = []
        for i in range(min(m_ref, n)):
            p = os.path.join(wd, f"g{i:05d}.gz")
            with gzip.open(p, "wb", compresslevel=1) as f:
                f.write(sks[i])
            names.append(p)
        fof = os.path.join(wd, "sk.txt")
        open(fof, "w").write("\n".join(names) + "\n")
        r = subprocess.run([os.path.join(O.REF_DIR, "comparator"), "-f", fof, "-o", os.path.join(wd, "res")], cwd=wd,
                           stdin=subprocess.DEVNULL, capture_output=True, text=True)
        for ln in r.stdout.splitlines():
            if ln.startswith("Comparisons lasted"):
                sec = float(ln.split()[2]); mm = len(names)
                print(f"reference comparator on {mm} of these sketches: {sec:.2f} s -> {mm * (mm - 1) // 2 / sec:.3e} pairs/s (1 thread)")
        # parity of the counts on that subset (CSV text at -p 6)
        with gzip.open(os.path.join(wd, "res_jaccard.csv.gz"), "rb") as f:
            ref_jac = f.read()
        mm = len(names)
        sub = np.zeros((mm, mm), np.uint32)
        sub[:, :] = inter[:mm, :mm]
        sizes = np.diff(off)[:mm].astype(np.uint64)
        ours = S.format_csv(names, mm, sub, False, sizes, True, 6, 0.0)
        print("jaccard CSV identical to the reference on the subset:", ours == ref_jac)


if __name__ == "__main__":
    main()
