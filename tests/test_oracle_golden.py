"""Pins the CPU restatement (oracle/spsp_oracle.c) to the unmodified reference:
tests/golden/golden.json holds SHA-256 of the gunzipped sketches / CSVs that
oracle/_ref (the reference built from /root/reference) wrote for the seeded
inputs of tests/golden_inputs.py (generator: tools/make_golden.py)."""
import hashlib
import os

import pytest

from tests.golden_inputs import COMPARE_CASES, SKETCH_CASES, build_input
from tests.conftest import GOLDEN_DIR


def sha(b):
    return hashlib.sha256(b).hexdigest()


@pytest.mark.parametrize("name", sorted(SKETCH_CASES))
def test_sketch_matches_reference(name, golden, oracle):
    inp, k, m, s, a = SKETCH_CASES[name]
    sk, st = oracle.sketch(build_input(inp), k, m, s, a)
    g = golden["sketch"][name]
    assert len(sk) == g["len"]
    assert sk.split(b"\n", 1)[0].decode() == g["header"]
    assert sha(sk) == g["sha256"]
    assert st["pb_events"] == 0
    raw = os.path.join(GOLDEN_DIR, name + ".sketch")
    if os.path.exists(raw):
        assert open(raw, "rb").read() == sk


@pytest.mark.parametrize("name", sorted(COMPARE_CASES))
def test_compare_matches_reference(name, golden, oracle):
    inputs, k, m, s, nq, prec, thr = COMPARE_CASES[name]
    sks = [oracle.sketch(build_input(i), k, m, s)[0] for i in inputs]
    g = golden["compare"][name]
    assert [sha(x) for x in sks] == g["sketch_sha256"]
    q = nq if nq else len(inputs)
    inter, sizes, kk, mm = oracle.compare(sks, q)
    assert (kk, mm) == (k, m)
    names = [i + ".gz" for i in inputs]
    cont = oracle.csv(names, q, inter, sizes, False, prec, thr)
    jac = oracle.csv(names, q, inter, sizes, True, prec, thr)
    assert sha(cont) == g["containment_sha256"]
    assert sha(jac) == g["jaccard_sha256"]
    raw = os.path.join(GOLDEN_DIR, name + ".jaccard.csv")
    if os.path.exists(raw):
        assert open(raw, "rb").read() == jac


def test_scalars(golden, oracle):
    for x, h in golden["hash"].items():
        assert oracle.xxh64_8(int(x)) == h
    for key, t in golden["threshold"].items():
        k, m, s = key.split(",")
        assert oracle.threshold(int(k), int(m), float(s)) == t
    # thresholds printed by the verified restatement
    assert oracle.threshold(31, 11, 1000) == 878834950402620
    assert oracle.threshold(31, 11, 100) == 8826267444307540
    assert oracle.threshold(31, 13, 200) == 4865941067100300


def test_closed_form_selection(oracle):
    """a k-mer occurrence is selected iff one of its k-m+1
    canonical m-mers has hash <= T (dense state machine vs closed form)."""
    import numpy as np
    for inp, k, m, s in (("multi", 31, 11, 20), ("nasty", 31, 11, 10), ("reads", 21, 9, 30)):
        fa = build_input(inp)
        _, st, sel = oracle.sketch(fa, k, m, s, trace=True)
        bases, offs = oracle.clean(fa)
        thr = oracle.threshold(k, m, s)
        want = []
        rid = 0
        for r in range(len(offs) - 1):
            seq = bases[int(offs[r]):int(offs[r + 1])]
            if seq.size < k:
                rid += 1
                continue
            pos, _, _, _ = oracle.hits(seq, m, thr)
            cover = np.zeros(seq.size - k + 2, np.int64)
            lo = np.maximum(pos.astype(np.int64) - (k - m), 0)
            hi = np.minimum(pos.astype(np.int64), seq.size - k) + 1
            np.add.at(cover, lo, 1)
            np.add.at(cover, hi, -1)
            starts = np.flatnonzero(np.cumsum(cover)[:-1] > 0)
            want += [(rid, int(x)) for x in starts]
            rid += 1
        got = sorted((int(a), int(b)) for a, b in sel)
        assert got == sorted(want)
        assert st["selected_kmers"] == len(want)


def test_oracle_stats_match_reference_print_stat(oracle):
    """The restatement's dense state machine reproduces the totals the reference binary prints
    (print_stat, SubSampler.cpp:633-665; tests/golden/stats.json from tools/make_golden_stats.py)."""
    import json
    with open(os.path.join(GOLDEN_DIR, "stats.json")) as f:
        stats = json.load(f)
    checked = 0
    for name, want in sorted(stats.items()):
        inp, k, m, s, a = SKETCH_CASES[name]
        _, st = oracle.sketch(build_input(inp), k, m, s, a)
        if want.get("none_selected"):
            assert st["selected_kmers"] == 0
            continue
        for key in ("total_kmers", "total_superkmers", "selected_kmers", "selected_superkmers"):
            assert st[key] == want[key], (name, key, st[key], want[key])
        checked += 1
    assert checked >= 25
