#!/usr/bin/env python
"""One-off parity runs at larger-than-unit-test sizes against the UNMODIFIED reference binaries (oracle/_ref)
on the GPU box, through the drop-in CLIs: C1 (one 5 Mbp genome, -i), C4 shape (a 100 Mbp set of 150 bp reads),
C5 shape (query mode, k31 m13 s200, 20 queries x 300 references of 1 Mbp).  Prints one line per check."""
import gzip, os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import supersampler_b200 as S
from supersampler_b200 import synth, capi
from oracle import oracle as O


def run(cmd, cwd):
    t0 = time.perf_counter()
    r = subprocess.run(cmd, cwd=cwd, stdin=subprocess.DEVNULL, capture_output=True, text=True)
    assert r.returncode == 0, (cmd, r.stderr[-500:])
    return time.perf_counter() - t0, r.stdout


def gunzip(p):
    with gzip.open(p, "rb") as f:
        return f.read()


def main():
    S.build(); O.build(with_ref=False)
    assert O.have_ref(), "oracle/_ref missing"
    ours_s, ours_c = os.path.join(capi.BIN_DIR, "sub_sampler"), os.path.join(capi.BIN_DIR, "comparator")
    ref_s, ref_c = os.path.join(O.REF_DIR, "sub_sampler"), os.path.join(O.REF_DIR, "comparator")
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    wd = tempfile.mkdtemp(prefix="spsp_val_", dir=base)
    a, b = os.path.join(wd, "ours"), os.path.join(wd, "ref")
    os.makedirs(a); os.makedirs(b)
    # ---- C1
    p = os.path.join(wd, "c1.fa")
    open(p, "wb").write(synth.fasta_bytes([("c1", synth.random_genome(5_000_000, 1))]))
    t1, _ = run([ours_s, "-i", p, "-v", "0"], a); t2, _ = run([ref_s, "-i", p, "-v", "0"], b)
    print(f"C1  5 Mbp -i            : identical={gunzip(a + '/subsampled_c1.gz') == gunzip(b + '/subsampled_c1.gz')}  ours {t1:.2f} s (process start + CUDA init included), reference {t2:.2f} s")
    # ---- C4 shape
    g = synth.random_genome(5_000_000, 7)
    p = os.path.join(wd, "reads.fa")
    open(p, "wb").write(synth.reads_fasta_bytes(synth.read_set(666_667, 150, g, 8)))
    t1, _ = run([ours_s, "-i", p, "-v", "1"], a); t2, out2 = run([ref_s, "-i", p, "-v", "1"], b)
    same = gunzip(a + '/subsampled_reads.gz') == gunzip(b + '/subsampled_reads.gz')
    print(f"C4  100 Mbp of 150 bp reads: identical={same}  ours {t1:.2f} s, reference {t2:.2f} s")
    # ---- C5 shape
    k, m, s = 31, 13, 200
    fam = list(synth.genome_family(320, 1_000_000, seed=77))
    paths = []
    for nm, gg in fam:
        pp = os.path.join(wd, nm + ".fa")
        open(pp, "wb").write(synth.fasta_bytes([(nm, gg)]))
        paths.append(pp)
    fof = os.path.join(wd, "all.txt"); open(fof, "w").write("\n".join(paths) + "\n")
    t1, _ = run([ours_s, "-f", fof, "-k", str(k), "-m", str(m), "-s", str(s), "-t", "16", "-v", "0"], a)
    t2, _ = run([ref_s, "-f", fof, "-k", str(k), "-m", str(m), "-s", str(s), "-t", "16", "-v", "0"], b)
    names = ["subsampled_" + nm + ".gz" for nm, _ in fam]
    same = sum(gunzip(os.path.join(a, n_)) == gunzip(os.path.join(b, n_)) for n_ in names)
    print(f"C5  sketch 320 x 1 Mbp k31 m13 s200: {same}/{len(names)} identical  ours {t1:.2f} s, reference {t2:.2f} s")
    for d in (a, b):
        open(os.path.join(d, "q.txt"), "w").write("\n".join(names[:20]) + "\n")
        open(os.path.join(d, "r.txt"), "w").write("\n".join(names[20:]) + "\n")
    t1, _ = run([ours_c, "-f", "r.txt", "-q", "q.txt", "-o", "res"], a)
    t2, _ = run([ref_c, "-f", "r.txt", "-q", "q.txt", "-o", "res"], b)
    okc = gunzip(a + "/res_containment.csv.gz") == gunzip(b + "/res_containment.csv.gz")
    okj = gunzip(a + "/res_jaccard.csv.gz") == gunzip(b + "/res_jaccard.csv.gz")
    print(f"C5  comparator -q 20 x 300: containment identical={okc} jaccard identical={okj}  ours {t1:.2f} s, reference {t2:.2f} s")


if __name__ == "__main__":
    main()
