#!/usr/bin/env python
"""Pin print_stat's totals (SubSampler.cpp:633-665) to the UNMODIFIED reference binary:
runs oracle/_ref/sub_sampler -v 1 on the seeded inputs of tests/golden_inputs.py and stores
the numbers it prints in tests/golden/stats.json.  Build container only (needs oracle/_ref).

    python tools/make_golden_stats.py
"""
import json
import os
import re
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O            # noqa: E402
from tests.golden_inputs import SKETCH_CASES, build_input   # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "stats.json")


def num(s):
    return int(re.sub(r"[^0-9]", "", s))


def main():
    O.build()
    assert O.have_ref(), "oracle/_ref missing: run make -C oracle"
    wd = tempfile.mkdtemp(prefix="golden_stats_")
    gold = {}
    for name, (inp, k, m, s, a) in sorted(SKETCH_CASES.items()):
        p = os.path.join(wd, inp + ".fa")
        if not os.path.exists(p):
            with open(p, "wb") as f:
                f.write(build_input(inp))
        r = subprocess.run([os.path.join(O.REF_DIR, "sub_sampler"), "-i", p, "-k", str(k), "-m", str(m), "-s", repr(float(s)),
                            "-v", "1", "-a", str(a)], cwd=wd, stdin=subprocess.DEVNULL, capture_output=True, text=True, check=True)
        st = {}
        for ln in r.stdout.splitlines():
            mm = re.match(r"I have seen ([0-9,. ]+) kmers and I selected ([0-9,. ]+) kmers", ln)
            if mm:
                st["total_kmers"], st["selected_kmers"] = num(mm.group(1)), num(mm.group(2))
            mm = re.match(r"I have seen ([0-9,. ]+) superkmers and I selected ([0-9,. ]+) superkmers", ln)
            if mm:
                st["total_superkmers"], st["selected_superkmers"] = num(mm.group(1)), num(mm.group(2))
            mm = re.match(r"After removing duplicate kmers, I selected ([0-9,. ]+) kmers", ln)
            if mm:
                st["distinct_kmers"] = num(mm.group(1))
            if "Crickets" in ln:
                st["none_selected"] = True
        gold[name] = st
        print(name, st)
    with open(OUT, "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
