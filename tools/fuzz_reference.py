#!/usr/bin/env python
"""Differential fuzzing on the CPU (build container, needs oracle/_ref): random messy FASTA bytes -- binary junk,
'>' and 0xFF in odd places, CR/LF mixes, empty lines, no final newline -- through (1) the unmodified reference
binary, (2) the oracle's C restatement, (3) the product's host route (packer + closed-form hits + exact host
post-pass).  Prints every case on which they disagree.
    python tools/fuzz_reference.py [cases=300] [seed=1]"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import supersampler_b200 as S
from oracle import oracle as O
from tests.test_host_logic import oracle_hits


def messy(rng):
    out = bytearray()
    junk = list(b"NnRYKM-*> \t\r\x00\x01\x7f\x80\xfe\xff>>\xff")
    n_rec = int(rng.integers(0, 12))
    if rng.integers(0, 4) == 0:                       # headerless first line
        out += bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(0, 80))).astype(np.uint8)) + b"\n"
    for r in range(n_rec):
        hdr = rng.integers(0, 8)
        out += b">" + (b"r%d ACGT" % r if hdr else b"") + (b"\xff" if hdr == 1 else b"") + (b"\r\n" if rng.integers(0, 5) == 0 else b"\n")
        L = int(rng.integers(0, 900))
        seq = bytearray(bytes(rng.choice(list(b"ACGTacgt"), size=L).astype(np.uint8)))
        kind = int(rng.integers(0, 8))
        if kind == 0 and L > 30:                         # tandem repeat / low complexity
            u = seq[: int(rng.integers(1, 15))]
            seq = bytearray((bytes(u) * (L // len(u) + 1))[:L])
        elif kind == 1 and out.count(b"\n") > 3 and rng.integers(0, 2):      # a copy of earlier text (duplicate k-mers, count wrap)
            prev = bytes(out)
            a0 = int(rng.integers(0, max(1, len(prev) - 200)))
            seq = bytearray(prev[a0:a0 + int(rng.integers(50, 400))].replace(b">", b"N"))
        elif kind == 2 and L > 40:                       # inverted repeat
            comp = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")
            u = bytes(seq[: L // 3])
            seq = bytearray(u + u.translate(comp)[::-1] + u)
        for _ in range(int(rng.integers(0, 6))):
            p = int(rng.integers(0, len(seq) + 1))
            seq[p:p] = bytes(rng.choice(junk, size=int(rng.integers(1, 6))).astype(np.uint8))
        width = int(rng.integers(1, 100))
        for i in range(0, len(seq), width):
            out += seq[i:i + width] + (b"\r\n" if rng.integers(0, 6) == 0 else b"\n")
            if rng.integers(0, 25) == 0:
                out += b"\n"
    if rng.integers(0, 2) and out.endswith(b"\n"):
        out = out[:-1]
    return bytes(out)


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    O.build()
    assert O.have_ref()
    wd = tempfile.mkdtemp(prefix="fuzz_")
    bad = 0
    for c in range(cases):
        m = int(rng.choice([5, 7, 9, 11, 13, 15]))
        k = int(rng.choice([x for x in (15, 17, 21, 31, 33, 45, 63) if x > m + 1]))
        s = float(rng.choice([1, 1.5, 2, 4, 10, 50]))
        a = int(rng.choice([1, 1, 1, 2, 3]))
        fa = messy(rng)
        if rng.integers(0, 12) == 0:                      # many copies of one short record: uint8 count wrap
            fa = fa + (b">dup\n" + bytes(rng.choice(list(b"ACGT"), size=70).astype(np.uint8)) + b"\n") * int(rng.integers(250, 600))
        p = os.path.join(wd, f"c{c}.fa")
        open(p, "wb").write(fa)
        ref = O.ref_sketch_files([p], k, m, s, workdir=wd, abundance=a)[0]
        ora = O.sketch(fa, k, m, s, a)[0]
        words, nb, offs = S.pack_fasta(fa, k)
        hits = oracle_hits(O, words, nb, m, S.threshold(k, m, s))
        host = S.postpass(words, offs, hits, k, m, s, a)[0]
        os.remove(p)
        if not (ref == ora == host):
            bad += 1
            keep = os.path.join(wd, f"bad{c}_k{k}_m{m}_s{s}.fa")
            open(keep, "wb").write(fa)
            print(f"case {c} k{k} m{m} s{s} a{a}: reference == oracle {ref == ora}, reference == host route {ref == host}, kept {keep}")
    print(f"{cases} cases, {bad} disagreements")


if __name__ == "__main__":
    main()
