"""CPU tests of the product's host layer (libspsp_host.so): FASTA cleaner/packer,
the exact post-pass (sparse replay -> sketch bytes), the sketch decoder and the
CSV writer.  The GPU's job -- the hit list -- is supplied here by the oracle's
closed form (checker input only), so the bytes can be compared with the
reference goldens without a GPU."""
import hashlib
import os

import numpy as np
import pytest

import supersampler_b200 as S
from tests.conftest import GOLDEN_DIR
from tests.golden_inputs import COMPARE_CASES, SKETCH_CASES, build_input


@pytest.fixture(scope="module", autouse=True)
def _built():
    S.build()


def sha(b):
    return hashlib.sha256(b).hexdigest()


def unpack(words, n):
    w = np.asarray(words, np.uint32)
    codes = ((w[:, None] >> (30 - 2 * np.arange(16, dtype=np.uint32))[None, :]) & 3).reshape(-1)[:n]
    return np.frombuffer(b"ACTG", np.uint8)[codes]


def oracle_hits(oracle, words, n_bases, m, thr):
    """Closed-form hits over the whole packed buffer (records ignored, like the kernel)."""
    seq = unpack(words, n_bases)
    pos, cn, rv, _ = oracle.hits(seq, m, thr)
    h = np.zeros(pos.size, S.HIT_DTYPE)
    h["pos"], h["canon"], h["rev"] = pos, cn, rv
    return h


@pytest.mark.parametrize("inp", ["nasty", "multi", "noheader", "empty", "tiny", "reads", "wrap257"])
def test_packer_matches_clean_dna(inp, oracle):
    fa = build_input(inp)
    for k in (1, 21, 31):
        words, nb, offs = S.pack_fasta(fa, k)
        bases, o = oracle.clean(fa)
        keep = [(int(o[i]), int(o[i + 1])) for i in range(len(o) - 1) if int(o[i + 1]) - int(o[i]) >= k]
        want = np.concatenate([bases[a:b] for a, b in keep]) if keep else np.zeros(0, np.uint8)
        assert nb == want.size
        assert np.array_equal(unpack(words, nb), want)
        assert list(np.diff(offs.astype(np.int64))) == [b - a for a, b in keep]
        assert words.size == S.packed_words(nb)
        # zero padding behind the last base
        tail = unpack(words, words.size * 16)[nb:]
        assert (tail == ord("A")).all()


def test_packer_streaming_chunks(oracle):
    """Same result whatever the chunking (state machine across buffer boundaries)."""
    fa = build_input("nasty")
    ref = S.pack_fasta(fa, 31)
    crlf = fa.replace(b"\n", b"\r\n")
    w2, nb2, o2 = S.pack_fasta(crlf, 31)
    assert nb2 == ref[1] and np.array_equal(w2, ref[0]) and np.array_equal(o2, ref[2])


@pytest.mark.parametrize("name", sorted(SKETCH_CASES))
def test_postpass_reproduces_reference_sketch(name, golden, oracle):
    inp, k, m, s, a = SKETCH_CASES[name]
    fa = build_input(inp)
    words, nb, offs = S.pack_fasta(fa, k)
    thr = S.threshold(k, m, s)
    assert thr == oracle.threshold(k, m, s)
    hits = oracle_hits(oracle, words, nb, m, thr)
    rng = np.random.default_rng(0)
    rng.shuffle(hits)                       # the kernel's output is unordered
    sk, nsel = S.postpass(words, offs, hits, k, m, s, a)
    g = golden["sketch"][name]
    assert len(sk) == g["len"]
    assert sk.split(b"\n", 1)[0].decode() == g["header"]
    assert sha(sk) == g["sha256"]


def _decode_all(sks):
    el = [S.decode_sketch(x) for x in sks]
    return el


def _intersections(el):
    n = len(el)
    inter = np.zeros((n, n), np.uint32)
    keys = []
    for k, m, mn, lo, hi in el:
        rec = np.zeros(mn.size, dtype=[("m", "<u4"), ("h", "<u8"), ("l", "<u8")])
        rec["m"], rec["l"] = mn, lo
        if hi is not None:
            rec["h"] = hi
        assert np.unique(rec).size == rec.size          # elements are distinct
        assert (np.diff(mn.astype(np.int64)) >= 0).all()  # buckets ascending
        keys.append(rec)
    for i in range(n):
        for j in range(i + 1, n):
            inter[i, j] = np.intersect1d(keys[i], keys[j]).size
    return inter, np.array([x.size for x in keys], np.uint64)


@pytest.mark.parametrize("name", sorted(COMPARE_CASES))
def test_decode_and_csv_match_reference(name, golden, oracle):
    inputs, k, m, s, nq, prec, thr = COMPARE_CASES[name]
    sks = [oracle.sketch(build_input(i), k, m, s)[0] for i in inputs]
    el = _decode_all(sks)
    assert all(e[0] == k and e[1] == m for e in el)
    inter, sizes = _intersections(el)
    o_inter, o_sizes, _, _ = oracle.compare(sks, None)
    assert np.array_equal(sizes, o_sizes)
    assert np.array_equal(np.triu(inter, 1), np.triu(o_inter, 1))
    q = nq if nq else len(inputs)
    names = [i + ".gz" for i in inputs]
    g = golden["compare"][name]
    if q == len(inputs):
        cont = S.format_csv(names, q, inter, False, sizes, False, prec, thr)
        jac = S.format_csv(names, q, inter, False, sizes, True, prec, thr)
    else:
        full = (inter + inter.T)[:q]
        cont = S.format_csv(names, q, full, True, sizes, False, prec, thr)
        jac = S.format_csv(names, q, full, True, sizes, True, prec, thr)
    assert sha(cont) == g["containment_sha256"]
    assert sha(jac) == g["jaccard_sha256"]


def test_k63_elements_have_high_words(oracle):
    sk = oracle.sketch(build_input("nasty"), 63, 15, 10)[0]
    k, m, mn, lo, hi = S.decode_sketch(sk)
    assert (k, m) == (63, 15) and hi is not None and hi.size == lo.size and hi.max() > 0


def test_out_name_and_threshold():
    assert S.threshold(31, 11, 1) == 2 ** 64 - 1
    assert S.threshold(31, 11, 0.5) == 2 ** 64 - 1


def test_postpass_batch_one_scan_many_inputs(oracle):
    """Several inputs packed back to back, one hit list over everything (what a
    single kernel launch over a batch produces): per-input sketches are unchanged
    and hits in the padding are ignored."""
    k, m, s = 31, 11, 20
    inputs = ["multi", "nasty", "tiny", "reads", "empty"]
    ws, bo, nbs, ro, off = [], [], [], [], 0
    for i in inputs:
        w, nb, offs = S.pack_fasta(build_input(i), k)
        ws.append(w); bo.append(off); nbs.append(nb); ro.append(offs)
        off += w.size * 16
    packed = np.concatenate(ws)
    rec_first = np.array([0] + list(np.cumsum([r.size for r in ro])), np.uint64)
    thr = S.threshold(k, m, s)
    hits = oracle_hits(oracle, packed, packed.size * 16, m, thr)      # includes poly-A padding hits, if any
    out = S.postpass_batch(packed, np.array(bo), np.array(nbs), np.concatenate(ro), rec_first, hits, k, m, s, threads=3)
    for i, sk in zip(inputs, out):
        assert sk == oracle.sketch(build_input(i), k, m, s)[0], i


def test_sort_csv_matches_reference(tmp_path):
    """sortCSV mirror (sort_csv.cpp:26-122) against the reference binary's output
    (tests/golden/fam12_s100.jaccard.sorted.csv, generated with oracle/_ref/sortCSV)."""
    import gzip
    jac = open(os.path.join(GOLDEN_DIR, "fam12_s100.jaccard.csv"), "rb").read()
    want = open(os.path.join(GOLDEN_DIR, "fam12_s100.jaccard.sorted.csv"), "rb").read()
    order = open(os.path.join(GOLDEN_DIR, "fam12_s100.sort_order.txt"), "rb").read()
    (tmp_path / "names.txt").write_bytes(order)
    with gzip.open(tmp_path / "j.csv.gz", "wb") as f:
        f.write(jac)
    (tmp_path / "j.csv").write_bytes(jac)
    for src in ("j.csv.gz", "j.csv"):
        out = tmp_path / ("sorted_" + src + ".csv")
        assert S.run_sort_csv([str(tmp_path / src), str(out), str(tmp_path / "names.txt")]) == 0
        assert out.read_bytes() == want


def test_packer_random_messy_inputs(oracle):
    """Random line widths, CRLF, junk bytes, '>' inside lines, empty records: packer == clean_dna restatement,
    for the AVX-512 and the AVX2 block loops and for sliced feeding."""
    import subprocess, sys
    from tests.test_gpu_parity import _random_fasta
    for seed in range(25):
        rng = np.random.default_rng(seed)
        fa = _random_fasta(rng, int(rng.integers(1, 40)))
        for k in (1, 31):
            words, nb, offs = S.pack_fasta(fa, k)
            bases, o = oracle.clean(fa)
            keep = [(int(o[i]), int(o[i + 1])) for i in range(len(o) - 1) if int(o[i + 1]) - int(o[i]) >= k]
            want = np.concatenate([bases[a:b] for a, b in keep]) if keep else np.zeros(0, np.uint8)
            assert nb == want.size, (seed, k)
            assert np.array_equal(unpack(words, nb), want), (seed, k)
            assert list(np.diff(offs.astype(np.int64))) == [b - a for a, b in keep]
    # the same check with the AVX-512 loop disabled (run-time dispatch is decided once per process)
    code = ("import numpy as np, supersampler_b200 as S\n"
            "from tests.test_gpu_parity import _random_fasta\n"
            "import hashlib\n"
            "h = hashlib.sha256()\n"
            "for seed in range(25):\n"
            "    rng = np.random.default_rng(seed)\n"
            "    fa = _random_fasta(rng, int(rng.integers(1, 40)))\n"
            "    w, nb, o = S.pack_fasta(fa, 31)\n"
            "    h.update(w.tobytes()); h.update(o.tobytes())\n"
            "print(h.hexdigest())\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for env_extra in ({}, {"SPSP_NO_AVX512": "1"}):
        env = dict(os.environ, **env_extra)
        r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs.append(r.stdout.strip())
    assert outs[0] == outs[1]


def test_packer_long_messy_input_crosses_sweep_pieces(oracle):
    """Inputs longer than the packer's internal piece (512 KB of text between two sweeps into the output
    buffer): thousands of records of every length around min_len, so that short records are taken back and
    pending words are carried over right at the piece boundaries; plus one long record crossing several pieces."""
    from tests.test_gpu_parity import _random_fasta
    rng = np.random.default_rng(2024)
    parts = []
    size = 0
    while size < 2_300_000:
        if rng.integers(0, 40) == 0:
            L = int(rng.integers(200_000, 700_000))              # a record that spans pieces
            seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=L)].tobytes()
            w = int(rng.integers(40, 100))
            piece = b">long\n" + b"\n".join(seq[i:i + w] for i in range(0, L, w)) + b"\n"
        else:
            piece = _random_fasta(rng, int(rng.integers(1, 30)))
            if not piece.endswith(b"\n"):
                piece += b"\n"
        parts.append(piece)
        size += len(piece)
    fa = b"".join(parts)
    bases, o = oracle.clean(fa)
    for k in (1, 31, 63):
        words, nb, offs = S.pack_fasta(fa, k)
        keep = [(int(o[i]), int(o[i + 1])) for i in range(len(o) - 1) if int(o[i + 1]) - int(o[i]) >= k]
        want = np.concatenate([bases[a:b] for a, b in keep]) if keep else np.zeros(0, np.uint8)
        assert nb == want.size, k
        assert np.array_equal(unpack(words, nb), want), k
        assert list(np.diff(offs.astype(np.int64))) == [b - a for a, b in keep]
        assert words.size == S.packed_words(nb)
        assert (unpack(words, words.size * 16)[nb:] == ord("A")).all()



def test_streamed_csv_equals_in_memory_csv(tmp_path):
    """write_csv_gz (parallel row blocks, consecutive gzip members) gives the bytes of format_csv after gunzip:
    all-vs-all and query mode, containment and Jaccard, several block waves, an empty matrix."""
    import gzip
    rng = np.random.default_rng(5)
    for n, q, full in ((1, 1, False), (7, 7, False), (300, 300, False), (1200, 40, True), (2500, 2500, False), (0, 0, False)):
        names = [f"sketch_{i}.gz" for i in range(n)]
        sizes = rng.integers(1, 5000, size=max(n, 1)).astype(np.uint64)[:n]
        rows = q if full else n
        inter = rng.integers(0, 3, size=(max(rows, 1), max(n, 1))).astype(np.uint32)
        inter *= rng.integers(0, 900, size=inter.shape).astype(np.uint32)
        if n:
            inter = np.minimum(inter, np.minimum(sizes[:rows, None], sizes[None, :]).astype(np.uint32))
        for jac in (False, True):
            want = S.format_csv(names, q, inter, full, sizes, jac, 6, 0.01)
            p = tmp_path / f"m_{n}_{int(jac)}.csv.gz"
            nbytes = S.write_csv_gz(str(p), names, q, inter, full, sizes, jac, 6, 0.01, threads=3)
            with gzip.open(p, "rb") as f:
                got = f.read()
            assert got == want and nbytes == len(want), (n, q, jac)


def test_batch_sketches_sequence():
    """The lazy per-batch sketch list assembles header line + body on access and compares like a list."""
    from supersampler_b200.capi import BatchSketches
    body = b"AAAABBBCC"
    bs = BatchSketches(body, np.array([0, 4, 4, 7, 9], np.uint64), np.array([5, 0, 7, 9], np.uint64), "51 11 ", " 1000.000000\n")
    want = [b"51 11 5 1000.000000\nAAAA", b"51 11 0 1000.000000\n", b"51 11 7 1000.000000\nBBB", b"51 11 9 1000.000000\nCC"]
    assert len(bs) == 4 and list(bs) == want and bs == want and want == list(bs)
    assert bs[-1] == want[-1] and bs[1:3] == want[1:3] and [] + bs == want
    assert not (bs == want[:3]) and bs.nbytes == len(body) + 5 * 8 + 4 * 8
    with pytest.raises(IndexError):
        bs[4]


def test_cpulist_and_numa_binding_is_safe():
    from supersampler_b200 import distributed as D
    assert D._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert D._parse_cpulist("") == []
    before = os.sched_getaffinity(0)
    info = D.bind_rank_to_gpu_cores(0, 2, ["00000000:FF:1F.7", "00000000:FE:1F.7"])      # no such devices: must stay unbound
    assert info["bound"] is False and os.sched_getaffinity(0) == before


def test_truncated_gzip_is_not_a_partial_success(tmp_path, capfd):
    """A gzip stream that ends early or is corrupt must read as "cannot open", never as a shorter file (the
    reference's zstr throws there): checked through sortCSV, the CPU-only consumer of the shared reader."""
    import gzip
    jac = open(os.path.join(GOLDEN_DIR, "fam12_s100.jaccard.csv"), "rb").read()
    order = open(os.path.join(GOLDEN_DIR, "fam12_s100.sort_order.txt"), "rb").read()
    (tmp_path / "names.txt").write_bytes(order)
    z = gzip.compress(jac * 40)
    cases = {"cut.csv.gz": z[: len(z) // 2], "flip.csv.gz": z[:200] + bytes(b ^ 0x5A for b in z[200:260]) + z[260:]}
    for name, data in cases.items():
        (tmp_path / name).write_bytes(data)
        out = tmp_path / ("sorted_" + name)
        S.run_sort_csv([str(tmp_path / name), str(out), str(tmp_path / "names.txt")])
        assert not out.exists() or out.stat().st_size == 0, name
    assert "cant open file" in capfd.readouterr().out
