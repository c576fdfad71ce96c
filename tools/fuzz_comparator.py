#!/usr/bin/env python
"""Differential fuzzing of the compare path on the CPU (build container, needs oracle/_ref): random families of
related messy genomes are sketched by the reference sub_sampler, then compared by (1) the unmodified reference
comparator (all-vs-all or -q, random -p / -m), (2) the oracle's restatement (counts + CSV text), (3) the product's
host layer fed with the oracle's counts (decode_sketch sizes + format_csv + the streamed write_csv_gz).
    python tools/fuzz_comparator.py [cases=60] [seed=1]"""
import gzip, os, shutil, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import supersampler_b200 as S
from oracle import oracle as O


def family(rng, n, length):
    anc = rng.choice(list(b"ACGT"), size=length).astype(np.uint8)
    out = []
    for i in range(n):
        g = anc.copy()
        if rng.integers(0, 6) == 0:                         # an unrelated genome
            g = rng.choice(list(b"ACGT"), size=length).astype(np.uint8)
        else:
            rate = 10.0 ** rng.uniform(-3, -0.7)
            hit = rng.random(length) < rate
            g[hit] = rng.choice(list(b"ACGT"), size=int(hit.sum())).astype(np.uint8)
        recs = []
        cuts = sorted(rng.integers(0, length, size=int(rng.integers(0, 4))).tolist()) + [length]
        a = 0
        for c in cuts:                                      # several records per file, some empty
            recs.append(b">r\n" + g[a:c].tobytes() + b"\n")
            a = c
        if rng.integers(0, 10) == 0:
            recs = [b">only header\n"]                      # a sketch without buckets
        out.append(b"".join(recs))
    return out


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    O.build()
    assert O.have_ref()
    bad = 0
    for c in range(cases):
        wd = tempfile.mkdtemp(prefix="fuzzc_")
        m = int(rng.choice([7, 9, 11, 13]))
        k = int(rng.choice([x for x in (21, 31, 33, 63) if x > m + 1]))
        s = float(rng.choice([2, 5, 20, 100]))
        n = int(rng.integers(2, 40))
        nq = int(rng.integers(0, n))                         # 0 = all-vs-all
        prec = int(rng.choice([3, 6, 8]))
        thr = float(rng.choice([0.0, 0.0, 0.05, 0.3]))
        fas = family(rng, n, int(rng.integers(2000, 30000)))
        paths = []
        for i, fa in enumerate(fas):
            p = os.path.join(wd, f"g{i:03d}.fa")
            open(p, "wb").write(fa)
            paths.append(p)
        sks = O.ref_sketch_files(paths, k, m, s, workdir=wd)
        rel = [f"g{i:03d}.gz" for i in range(n)]
        for r, sk in zip(rel, sks):
            with gzip.open(os.path.join(wd, r), "wb") as f:
                f.write(sk)
        cont, jac, _ = O.ref_compare_files(rel[nq:] if nq else rel, rel[:nq] if nq else (), prec, thr, workdir=wd)
        inter, sizes, _, _ = O.compare(sks, nq if nq else None)
        q = nq if nq else n
        ok = True
        full = inter + inter.T if nq else inter
        for jacc, want in ((False, cont), (True, jac)):
            o_csv = O.csv(rel, q, inter, sizes, jacc, prec, thr)
            h_csv = S.format_csv(rel, q, full[:q] if nq else inter, bool(nq), sizes, jacc, prec, thr)
            pz = os.path.join(wd, "s.csv.gz")
            S.write_csv_gz(pz, rel, q, full[:q] if nq else inter, bool(nq), sizes, jacc, prec, thr, threads=3)
            with gzip.open(pz, "rb") as f:
                z_csv = f.read()
            ok = ok and o_csv == want and h_csv == want and z_csv == want
        dec = [S.decode_sketch(sk)[2].size for sk in sks]
        ok = ok and dec == [int(x) for x in sizes]
        if not ok:
            bad += 1
            print(f"case {c}: k{k} m{m} s{s} n{n} q{nq} p{prec} thr{thr} DISAGREE (kept {wd})")
        else:
            shutil.rmtree(wd, ignore_errors=True)
    print(f"{cases} cases, {bad} disagreements")


if __name__ == "__main__":
    main()
