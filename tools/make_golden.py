#!/usr/bin/env python
"""Generate tests/golden/golden.json from the UNMODIFIED reference binaries.

Run in the build container only (needs oracle/_ref, built by oracle/Makefile
from /root/reference).  Inputs are the seeded generators of
supersampler_b200.synth, so the GPU box can rebuild the same inputs and check
its outputs against the hashes stored here without the reference.

    python tools/make_golden.py
"""
import gzip
import hashlib
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O            # noqa: E402
from tests.golden_inputs import SKETCH_CASES, COMPARE_CASES, build_input   # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sha(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()


def main():
    O.build()
    assert O.have_ref(), "oracle/_ref missing: run make -C oracle"
    gold = {"sketch": {}, "compare": {}, "hash": {}, "threshold": {}}
    wd = tempfile.mkdtemp(prefix="golden_")
    # 1. sketches
    for name, (inp, k, m, s, a) in SKETCH_CASES.items():
        fa = build_input(inp)
        p = os.path.join(wd, inp + ".fa")
        if not os.path.exists(p):
            with open(p, "wb") as f:
                f.write(fa)
        sk = O.ref_sketch_files([p], k, m, s, abundance=a)[0]
        gold["sketch"][name] = {"sha256": sha(sk), "len": len(sk), "header": sk.split(b"\n", 1)[0].decode()}
        if len(sk) <= 4096:
            with open(os.path.join(OUT, name + ".sketch"), "wb") as f:
                f.write(sk)
        print(name, len(sk))
    # 2. compare (sketches made by the reference too)
    for name, (inputs, k, m, s, nq, prec, thr) in COMPARE_CASES.items():
        cd = os.path.join(wd, name)
        os.makedirs(cd)
        paths = []
        for inp in inputs:
            p = os.path.join(cd, inp + ".fa")
            with open(p, "wb") as f:
                f.write(build_input(inp))
            paths.append(p)
        sks = O.ref_sketch_files(paths, k, m, s, workdir=cd)
        rel = []
        for inp, sk in zip(inputs, sks):
            with gzip.open(os.path.join(cd, inp + ".gz"), "wb") as f:
                f.write(sk)
            rel.append(inp + ".gz")
        cont, jac, _ = O.ref_compare_files(rel[nq:] if nq else rel, rel[:nq] if nq else (), prec, thr, workdir=cd)
        gold["compare"][name] = {"containment_sha256": sha(cont), "jaccard_sha256": sha(jac),
                                 "containment_len": len(cont), "jaccard_len": len(jac),
                                 "sketch_sha256": [sha(x) for x in sks]}
        if len(cont) <= 8192:
            with open(os.path.join(OUT, name + ".containment.csv"), "wb") as f:
                f.write(cont)
            with open(os.path.join(OUT, name + ".jaccard.csv"), "wb") as f:
                f.write(jac)
        print(name, len(cont), len(jac))
    # 3. scalar known answers computed by the restatement and cross-checked
    #    through the sketches above (a wrong hash/threshold cannot reproduce them)
    for x in (0, 1, 2, 12345, 0x3FFFFF, 0x3FFFFFFF, 0x155555, 0xDEADBEEF):
        gold["hash"][str(x)] = O.xxh64_8(x)
    for k, m, s in ((31, 11, 1000), (31, 11, 100), (31, 13, 200), (21, 9, 5), (63, 15, 10), (31, 11, 1), (31, 11, 1.5)):
        gold["threshold"][f"{k},{m},{s}"] = O.threshold(k, m, s)
    with open(os.path.join(OUT, "golden.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    print("wrote", os.path.join(OUT, "golden.json"))


if __name__ == "__main__":
    main()
