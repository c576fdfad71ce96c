"""GPU parity tests (run on the B200 box with -m gpu): every call goes through
the C ABI (include/spsp.h via ctypes, or the host layer that calls it) and is
compared bit-exactly with the oracle / the reference goldens."""
import gzip
import hashlib
import os
import subprocess

import numpy as np
import pytest

import supersampler_b200 as S
from supersampler_b200 import capi, synth
from tests.golden_inputs import COMPARE_CASES, SKETCH_CASES, build_input

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _built():
    S.build()


def sha(b):
    return hashlib.sha256(b).hexdigest()


def unpack(words, n):
    w = np.asarray(words, np.uint32)
    codes = ((w[:, None] >> (30 - 2 * np.arange(16, dtype=np.uint32))[None, :]) & 3).reshape(-1)[:n]
    return np.frombuffer(b"ACTG", np.uint8)[codes]


def sorted_hits(h):
    return np.sort(h, order="pos")


@pytest.mark.parametrize("inp,k,m,s", [
    ("c1", 31, 11, 1000), ("c1", 31, 11, 100), ("c1", 31, 13, 200), ("nasty", 31, 11, 10), ("nasty", 21, 9, 5),
    ("nasty", 15, 5, 3), ("nasty", 63, 15, 10), ("reads", 31, 11, 50), ("multi", 31, 15, 20), ("tiny", 31, 11, 1000),
    ("nasty", 31, 11, 1), ("multi", 17, 7, 2),
])
def test_scan_kernels_match_closed_form(inp, k, m, s, oracle):
    """Dense and q-gram-filter kernels both emit exactly the closed-form hit set."""
    words, nb, _ = S.pack_fasta(build_input(inp), k)
    thr = S.threshold(k, m, s)
    seq = unpack(words, nb)
    pos, cn, rv, _ = oracle.hits(seq, m, thr)
    ctx = S.DeviceContext(k, m, thr)
    for mode in (capi.SCAN_DENSE, capi.SCAN_FILTER, capi.SCAN_AUTO):
        ctx.config(mode)
        h = sorted_hits(ctx.scan(words, nb))
        assert h.size == pos.size, (mode, h.size, pos.size)
        assert np.array_equal(h["pos"], pos)
        assert np.array_equal(h["canon"], cn)
        assert np.array_equal(h["rev"], rv.astype(np.uint32))
    ctx.close()


def test_scan_random_lengths(oracle):
    """Ragged sizes around the 64-base thread chunk / tile boundaries."""
    rng = np.random.default_rng(3)
    for k, m, s in ((31, 11, 20), (21, 9, 3), (31, 15, 50)):
        thr = S.threshold(k, m, s)
        ctx = S.DeviceContext(k, m, thr)
        for n in [0, 1, m - 1, m, m + 1, 63, 64, 65, 64 + m - 1, 127, 128, 4095, 4096, 4097, 16384 + 7, 65536 + 63,
                  int(rng.integers(100000, 300000))]:
            seq = synth.random_genome(n, int(rng.integers(1 << 30)))
            fa = b">r\n" + seq.tobytes() + b"\n"
            words, nb, _ = S.pack_fasta(fa, 1)
            assert nb == n
            pos, cn, rv, _ = oracle.hits(seq, m, thr)
            for mode in (capi.SCAN_DENSE, capi.SCAN_FILTER):
                ctx.config(mode)
                h = sorted_hits(ctx.scan(words, nb))
                assert np.array_equal(h["pos"], pos), (n, mode)
                assert np.array_equal(h["canon"], cn)
        ctx.close()


def test_scan_kernel_choice():
    """The cost model's choice per (m, T): (kind, probe stride).  Row-bit kernel for the reference's defaults, byte
    table of phase masks just below its crossover, bit tables (hashed where 2q > 20 bits) for dense thresholds
    and for m = 13 (C5).  Every choice returns the same hits (the tests above); this pins the selection."""
    want = {(31, 11, 1000.0): (2, 4), (31, 11, 600.0): (2, 4), (31, 11, 400.0): (1, 4), (31, 11, 100.0): (0, 2),
            (31, 13, 200.0): (0, 2)}
    for (k, m, s), (kind, g) in want.items():
        ctx = S.DeviceContext(k, m, S.threshold(k, m, s))
        info = ctx.filter_info()
        assert (info["kind"], info["g"]) == (kind, g), ((k, m, s), info)
        assert info["n_selected"] > 0
        ctx.close()


def test_scan_rowbit_filter(oracle, monkeypatch):
    """kind 2 (bank-private bit table + exact hash set, m == 11): same hit set as the closed form over ragged
    lengths, dense and sparse thresholds, all-A / repeated sequence (every probe flagged) and position 0."""
    monkeypatch.setenv("SPSP_FILTER_KIND", "2")
    rng = np.random.default_rng(17)
    for s in (250.0, 1000.0, 6000.0):
        thr = S.threshold(31, 11, s)
        ctx = S.DeviceContext(31, 11, thr)
        assert ctx.filter_info()["kind"] == 2, ctx.filter_info()
        ctx.config(capi.SCAN_FILTER)
        # a selected 11-mer planted at position 0 and at every alignment near chunk / pass boundaries
        probe = synth.random_genome(400000, 5)
        pos, cn, rv, _ = oracle.hits(probe, 11, thr)
        assert pos.size > 0
        sel = probe[int(pos[0]):int(pos[0]) + 11].copy()
        lens = [0, 1, 10, 11, 12, 15, 63, 64, 65, 74, 75, 127, 2047, 2048, 2049, 2048 + 11, 4096 + 3, 65536 + 63,
                int(rng.integers(100000, 300000)), 1_500_000]
        for n in lens:
            seq = synth.random_genome(n, int(rng.integers(1 << 30)))
            if n >= 11:
                for at in (0, 1, 2, 3, 4, 53, 54, 60, 61, 62, 63, 64, 65, 2037, 2040, 2048, n - 11, n - 12):
                    if 0 <= at and at + 11 <= n:
                        seq[at:at + 11] = sel
            fa = b">r\n" + seq.tobytes() + b"\n"
            words, nb, _ = S.pack_fasta(fa, 1)
            assert nb == n
            pos, cn, rv, _ = oracle.hits(seq, 11, thr)
            h = sorted_hits(ctx.scan(words, nb))
            assert np.array_equal(h["pos"], pos), (s, n, h.size, pos.size)
            assert np.array_equal(h["canon"], cn)
            assert np.array_equal(h["rev"], rv.astype(np.uint32))
        # the selected 11-mer repeated back to back: every probe of every lane is flagged, rings wrap
        seq = np.tile(sel, 30000)
        words, nb, _ = S.pack_fasta(b">r\n" + seq.tobytes() + b"\n", 1)
        pos, cn, rv, _ = oracle.hits(seq, 11, thr)
        h = sorted_hits(ctx.scan(words, nb))
        assert np.array_equal(h["pos"], pos) and np.array_equal(h["canon"], cn)
        ctx.close()


@pytest.mark.parametrize("name", sorted(SKETCH_CASES))
def test_sketch_bytes_match_reference(name, golden):
    inp, k, m, s, a = SKETCH_CASES[name]
    sk = S.sketch_buffers([build_input(inp)], k, m, s, a, threads=1)[0]
    g = golden["sketch"][name]
    assert len(sk) == g["len"]
    assert sha(sk) == g["sha256"]


def test_sketch_many_inputs_threads(golden):
    """File-parallel workers on separate streams give the same bytes."""
    names = [n for n in sorted(SKETCH_CASES) if SKETCH_CASES[n][1:] == (31, 11, 100, 1) or n.endswith("k31_m11_s100")]
    names = [n for n in names if SKETCH_CASES[n][1:4] == (31, 11, 100)]
    fas = [build_input(SKETCH_CASES[n][0]) for n in names] * 4
    out = S.sketch_buffers(fas, 31, 11, 100, 1, threads=6)
    for i, sk in enumerate(out):
        assert sha(sk) == golden["sketch"][names[i % len(names)]]["sha256"]


@pytest.mark.parametrize("name", sorted(COMPARE_CASES))
def test_compare_matches_reference(name, golden, oracle):
    inputs, k, m, s, nq, prec, thr = COMPARE_CASES[name]
    sks = S.sketch_buffers([build_input(i) for i in inputs], k, m, s, threads=4)
    g = golden["compare"][name]
    assert [sha(x) for x in sks] == g["sketch_sha256"]
    q = nq if nq else len(inputs)
    inter, sizes, full = S.compare_buffers(sks, q)
    o_inter, o_sizes, _, _ = oracle.compare(sks, q)
    assert np.array_equal(sizes, o_sizes)
    if full:
        want = (o_inter + o_inter.T)[:q]
        got = inter.copy()
        np.fill_diagonal(got[:, :q], 0)
        assert np.array_equal(got, want)
    else:
        assert np.array_equal(np.triu(inter, 1), np.triu(o_inter, 1))
    names = [i + ".gz" for i in inputs]
    cont = S.format_csv(names, q, inter, full, sizes, False, prec, thr)
    jac = S.format_csv(names, q, inter, full, sizes, True, prec, thr)
    assert sha(cont) == g["containment_sha256"]
    assert sha(jac) == g["jaccard_sha256"]


def test_compare_k63_and_multi_tile(oracle):
    """k > 32 (128-bit keys) and more than one 32x32 tile, ragged last tile."""
    fas = [synth.fasta_bytes([(nm, g)]) for nm, g in synth.genome_family(70, 40_000, seed=9)]
    for k, m, s in ((63, 15, 6), (31, 11, 8)):
        sks = S.sketch_buffers(fas, k, m, s, threads=8)
        for i in (0, 17, 69):
            assert sks[i] == oracle.sketch(fas[i], k, m, s)[0]
        inter, sizes, full = S.compare_buffers(sks)
        o_inter, o_sizes, _, _ = oracle.compare(sks)
        assert not full
        assert np.array_equal(sizes, o_sizes)
        assert np.array_equal(np.triu(inter, 1), np.triu(o_inter, 1))
        # query mode: 5 queries against all
        inter, sizes, full = S.compare_buffers(sks, 5)
        assert full
        want = (o_inter + o_inter.T)[:5]
        got = inter.copy()
        np.fill_diagonal(got[:, :5], 0)
        assert np.array_equal(got, want)


def test_compare_tile_ranks_partition(oracle):
    """Tiles dealt to R ranks sum to the single-rank result (multi-GPU sharding logic)."""
    fas = [synth.fasta_bytes([(nm, g)]) for nm, g in synth.genome_family(40, 30_000, seed=4)]
    k, m, s = 31, 11, 10
    sks = S.sketch_buffers(fas, k, m, s, threads=8)
    el = [S.decode_sketch(x) for x in sks]
    off = np.concatenate([[0], np.cumsum([e[2].size for e in el])]).astype(np.uint64)
    mn = np.concatenate([e[2] for e in el]); lo = np.concatenate([e[3] for e in el])
    ctx = S.DeviceContext(k, m, 0)
    ctx.cmp_load(off, mn, lo)
    n = len(sks)
    whole = ctx.cmp_run((0, n), (0, n), True)
    acc = np.zeros_like(whole)
    for r in range(3):
        acc += ctx.cmp_run((0, n), (0, n), True, r, 3)
    assert np.array_equal(acc, whole)
    o_inter, _, _, _ = oracle.compare(sks)
    assert np.array_equal(np.triu(whole, 1), np.triu(o_inter, 1))
    # a rectangular block
    blk = ctx.cmp_run((3, 9), (10, 40), False)
    assert np.array_equal(blk, (o_inter + o_inter.T)[3:9, 10:40])
    ctx.close()


def test_cli_end_to_end(tmp_path, golden):
    """The drop-in executables: same flags, files in the CWD, gz outputs."""
    exe_s = os.path.join(capi.BIN_DIR, "sub_sampler")
    exe_c = os.path.join(capi.BIN_DIR, "comparator")
    name = "fam12_s100"
    inputs, k, m, s, nq, prec, thr = COMPARE_CASES[name]
    paths = []
    for i, inp in enumerate(inputs):
        p = tmp_path / (inp + (".fa.gz" if i % 2 else ".fa"))
        data = build_input(inp)
        if i % 2:
            with gzip.open(p, "wb", compresslevel=1) as f:
                f.write(data)
        else:
            p.write_bytes(data)
        paths.append(str(p))
    fof = tmp_path / "genomes.txt"
    fof.write_text("\n".join(paths) + "\n")
    r = subprocess.run([exe_s, "-f", str(fof), "-k", str(k), "-m", str(m), "-s", str(s), "-t", "4", "-v", "0"],
                       cwd=tmp_path, stdin=subprocess.DEVNULL, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out_fof = tmp_path / "subsampled_genomes.txt"
    listed = out_fof.read_text().split()
    assert listed == ["subsampled_" + inp + ".gz" for inp in inputs]
    g = golden["compare"][name]
    for inp, want in zip(inputs, g["sketch_sha256"]):
        with gzip.open(tmp_path / ("subsampled_" + inp + ".gz"), "rb") as f:
            assert sha(f.read()) == want
    # comparator reads names relative to the CWD; rename to the golden's names
    for inp in inputs:
        os.rename(tmp_path / ("subsampled_" + inp + ".gz"), tmp_path / (inp + ".gz"))
    sk_fof = tmp_path / "sk.txt"
    sk_fof.write_text("\n".join(inp + ".gz" for inp in inputs) + "\n")
    r = subprocess.run([exe_c, "-f", "sk.txt", "-o", "res"], cwd=tmp_path, stdin=subprocess.DEVNULL,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    with gzip.open(tmp_path / "res_containment.csv.gz", "rb") as f:
        assert sha(f.read()) == g["containment_sha256"]
    with gzip.open(tmp_path / "res_jaccard.csv.gz", "rb") as f:
        assert sha(f.read()) == g["jaccard_sha256"]
    # single-input mode + in-process entry point
    os.chdir(tmp_path)
    assert S.run_sub_sampler(["-i", paths[0], "-k", str(k), "-m", str(m), "-s", str(s), "-p", "one_", "-v", "0"]) == 0
    with gzip.open(tmp_path / ("one_" + inputs[0] + ".gz"), "rb") as f:
        assert sha(f.read()) == g["sketch_sha256"][0]


def test_full_size_properties(oracle):
    """BASELINE config 2 shape (64 x 5 Mbp is bench territory; here 16 x 5 Mbp):
    size-independent properties + spot checks against the oracle."""
    fas = [synth.fasta_bytes([(nm, g)]) for nm, g in synth.genome_family(16, 5_000_000, seed=42)]
    k, m, s = 31, 11, 1000
    sks = S.sketch_buffers(fas, k, m, s, threads=8)
    again = S.sketch_buffers(fas[::-1], k, m, s, threads=3)[::-1]
    assert sks == again                                   # deterministic, order independent
    for i in (0, 7, 15):
        assert sks[i] == oracle.sketch(fas[i], k, m, s)[0]
    inter, sizes, _ = S.compare_buffers(sks)
    o_inter, o_sizes, _, _ = oracle.compare(sks)
    assert np.array_equal(sizes, o_sizes)
    assert np.array_equal(np.triu(inter, 1), np.triu(o_inter, 1))
    # self-comparison: a sketch against itself shares everything
    inter2, sizes2, _ = S.compare_buffers([sks[0], sks[0], sks[1]])
    assert inter2[0, 1] == sizes2[0] == sizes2[1]
    assert inter2[0, 2] == inter2[1, 2] == inter[0, 1]
    # intersections never exceed the smaller set
    iu = np.triu_indices(len(sks), 1)
    assert (inter[iu] <= np.minimum(sizes[iu[0]], sizes[iu[1]])).all()


# ---------------------------------------------------------------- whole-batch device path

def _batch(inputs, k):
    ws, ros = [], []
    for inp in inputs:
        w, nb, ro = S.pack_fasta(build_input(inp), k)
        ws.append(w); ros.append(ro)
    return S.batch_layout(ws, ros)


@pytest.mark.parametrize("k,m,s,a,inputs", [
    (31, 11, 1000, 1, ["c1", "c1mut", "nasty", "tiny", "empty", "reads"]),
    (31, 11, 100, 1, ["fam12_0", "fam12_1", "nasty", "multi", "reads", "noheader"]),
    (31, 11, 10, 1, ["nasty", "multi", "reads"]),
    (31, 11, 2, 1, ["nasty", "multi"]),
    (31, 11, 1, 1, ["nasty", "wrap256", "wrap257"]),
    (31, 11, 1, 2, ["wrap257", "nasty"]),
    (21, 9, 5, 1, ["nasty", "nasty21", "multi", "fam12_3"]),
    (15, 5, 3, 1, ["nasty", "multi"]),
    (63, 15, 10, 1, ["nasty", "multi", "fam12_2"]),
    (41, 13, 7, 1, ["nasty", "reads"]),
    (31, 13, 200, 1, ["c1", "nasty"]),
    (33, 13, 4, 1, ["nasty"]),
    (61, 15, 3, 1, ["nasty", "multi"]),
    (17, 15, 6, 1, ["nasty"]),
])
def test_device_postpass_matches_oracle(k, m, s, a, inputs, oracle):
    """Scan + device post-pass of a whole batch: sketch bytes of every input are
    the oracle's, and the device-resident compare elements give the oracle's counts."""
    words, nb, rb, re_, ri = _batch(inputs, k)
    ctx = S.DeviceContext(k, m, S.threshold(k, m, s))
    info = {}
    sks = ctx.sketch_batch(words, nb, rb, re_, ri, len(inputs), s, a, info=info)
    want = [oracle.sketch(build_input(i), k, m, s, a)[0] for i in inputs]
    for i, (g, w) in enumerate(zip(sks, want)):
        assert g == w, (inputs[i], len(g), len(w))
    # sketch -> compare hand-off on the device
    ctx.cmp_load_batch()
    n = len(inputs)
    inter = ctx.cmp_run((0, n), (0, n), True)
    o_inter, o_sizes, _, _ = oracle.compare(want)
    assert np.array_equal(np.diff(np.array(info["elem_off"], np.uint64)), o_sizes)
    assert np.array_equal(np.triu(inter, 1), np.triu(o_inter, 1))
    ctx.close()


def test_device_postpass_golden_and_host_agree(golden):
    """Every golden sketch case through the batch path, one input per batch."""
    for name in sorted(SKETCH_CASES):
        inp, k, m, s, a = SKETCH_CASES[name]
        words, nb, rb, re_, ri = _batch([inp], k)
        ctx = S.DeviceContext(k, m, S.threshold(k, m, s))
        sk = ctx.sketch_batch(words, nb, rb, re_, ri, 1, s, a)[0]
        ctx.close()
        assert sha(sk) == golden["sketch"][name]["sha256"], name


# ---------------------------------------------------------------- batch pipeline (host threads + one device batch)

def _golden_groups():
    groups = {}
    for name in sorted(SKETCH_CASES):
        inp, k, m, s, a = SKETCH_CASES[name]
        groups.setdefault((k, m, s, a), []).append((name, inp))
    return groups


def test_pipeline_matches_reference_goldens(golden):
    """Every golden sketch case through the batch pipeline, cases of one (k, m, s, a) in one batch."""
    for (k, m, s, a), cases in _golden_groups().items():
        pl = S.Pipeline(k, m, s, a, threads=4)
        info = {}
        sks = pl.sketch([build_input(inp) for _, inp in cases], info=info)
        pl.close()
        assert info["batches"] == 1
        for (name, _), sk in zip(cases, sks):
            assert sha(sk) == golden["sketch"][name]["sha256"], name


def test_pipeline_files_gz_missing_and_compare(tmp_path, golden, oracle):
    """File inputs (plain, gzip, unopenable) + device-resident hand-off to the compare stage."""
    name = "fam12_s100"
    inputs, k, m, s, nq, prec, thr = COMPARE_CASES[name]
    paths = []
    for i, inp in enumerate(inputs):
        p = tmp_path / (inp + (".fa.gz" if i % 3 == 1 else ".fa"))
        data = build_input(inp)
        if i % 3 == 1:
            with gzip.open(p, "wb", compresslevel=1) as f:
                f.write(data)
        else:
            p.write_bytes(data)
        paths.append(str(p))
    g = golden["compare"][name]
    pl = S.Pipeline(k, m, s, threads=5)
    sks = pl.sketch(paths)
    assert [sha(x) for x in sks] == g["sketch_sha256"]
    inter, sizes, full = pl.compare()
    o_inter, o_sizes, _, _ = oracle.compare(sks)
    assert not full
    assert np.array_equal(sizes, o_sizes)
    assert np.array_equal(np.triu(inter, 1), np.triu(o_inter, 1))
    names = [i + ".gz" for i in inputs]
    assert sha(S.format_csv(names, len(names), inter, full, sizes, False, prec, thr)) == g["containment_sha256"]
    assert sha(S.format_csv(names, len(names), inter, full, sizes, True, prec, thr)) == g["jaccard_sha256"]
    # query mode on the same device-resident elements
    inter_q, _, full_q = pl.compare(3)
    assert full_q
    want = (o_inter + o_inter.T)[:3]
    got = inter_q.copy()
    np.fill_diagonal(got[:, :3], 0)
    assert np.array_equal(got, want)
    # an unopenable file is skipped, the others are unaffected (SubSampler.cpp:313-322)
    mixed = pl.sketch([paths[0], str(tmp_path / "nope.fa"), build_input(inputs[2])])
    assert mixed[1] is None and sha(mixed[0]) == g["sketch_sha256"][0] and sha(mixed[2]) == g["sketch_sha256"][2]
    pl.close()


def test_pipeline_several_batches(oracle):
    """A job cut into several device batches gives the same bytes and counts as one batch."""
    fas = [synth.fasta_bytes([(nm, g)]) for nm, g in synth.genome_family(20, 60_000, seed=11)]
    fas.insert(7, build_input("reads"))
    fas.insert(3, build_input("empty"))
    k, m, s = 31, 11, 20
    one = S.Pipeline(k, m, s, threads=6)
    many = S.Pipeline(k, m, s, threads=3, max_batch_bases=200_000)
    i1, i2 = {}, {}
    a = one.sketch(fas, info=i1)
    b = many.sketch(fas, info=i2)
    assert i1["batches"] == 1 and i2["batches"] > 3
    assert a == b
    assert a == S.sketch_buffers(fas, k, m, s, threads=4)          # per-file path with the host post-pass
    ra, rb = one.compare(), many.compare()
    assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1])
    o_inter, o_sizes, _, _ = oracle.compare(a)
    assert np.array_equal(ra[1], o_sizes)
    assert np.array_equal(np.triu(ra[0], 1), np.triu(o_inter, 1))
    one.close(); many.close()


def test_pipeline_k63(oracle):
    fas = [build_input("nasty"), build_input("multi"), build_input("fam12_2")]
    k, m, s = 63, 15, 10
    for mb in (None, 100_000):
        pl = S.Pipeline(k, m, s, threads=2, max_batch_bases=mb)
        sks = pl.sketch(fas)
        assert sks == [oracle.sketch(f, k, m, s)[0] for f in fas]
        inter, sizes, _ = pl.compare()
        o_inter, o_sizes, _, _ = oracle.compare(sks)
        assert np.array_equal(sizes, o_sizes)
        assert np.array_equal(np.triu(inter, 1), np.triu(o_inter, 1))
        pl.close()


# ---------------------------------------------------------------- dense minimizer machine

@pytest.mark.parametrize("k,m,s,inputs", [
    (31, 11, 1000, ["c1", "nasty", "tiny", "empty", "reads", "multi"]),
    (31, 11, 10, ["nasty", "multi", "noheader", "fam12_0"]),
    (21, 9, 5, ["nasty21", "multi", "wrap257"]),
    (15, 5, 3, ["nasty", "multi"]),
    (63, 15, 10, ["nasty", "multi", "fam12_2"]),
    (63, 3, 4, ["nasty", "multi"]),
    (31, 13, 200, ["nasty", "reads"]),
    (17, 15, 6, ["nasty", "multi"]),
    (31, 11, 1, ["nasty", "multi"]),
])
def test_dense_machine_matches_oracle(k, m, s, inputs, oracle):
    """Warp-shuffle sliding-window minimum + parallel replay of the minimizer state machine:
    total_superkmer_number of every input equals the reference loop's (oracle: literal dense
    restatement of SubSampler.cpp:352-454), and the number of k-mers whose window minimum is
    <= T equals the selected k-mer count of the sketch header."""
    words, nb, rb, re_, ri = _batch(inputs, k)
    ctx = S.DeviceContext(k, m, S.threshold(k, m, s))
    info = {}
    tot, sel = ctx.dense_stats(words, nb, rb, re_, ri, len(inputs), info=info)
    for i, inp in enumerate(inputs):
        sk, st = oracle.sketch(build_input(inp), k, m, s)
        assert int(tot[i]) == st["total_superkmers"], (inp, int(tot[i]), st["total_superkmers"])
        assert int(sel[i]) == st["selected_kmers"], (inp, int(sel[i]), st["selected_kmers"])
        assert int(sel[i]) == int(sk.split(b"\n", 1)[0].split()[2])
    ctx.close()


def test_dense_machine_full_size_property():
    """Size-independent cross-check at full genome size: the dense kernel's selected k-mer count
    (window minimum <= T at every k-mer) equals the header field of the sketch built by the sparse
    hit path + device post-pass, for every input of a 16 x 5 Mbp batch."""
    fam = synth.Family(5_000_000, 42)
    k, m, s = 31, 11, 1000
    ws, ros = [], []
    for i in range(16):
        w, nb, ro = S.pack_fasta(fam.fasta(i), k)
        ws.append(w); ros.append(ro)
    words, n_total, rb, re_, ri = S.batch_layout(ws, ros)
    ctx = S.DeviceContext(k, m, S.threshold(k, m, s))
    sks = ctx.sketch_batch(words, n_total, rb, re_, ri, 16, s)
    tot, sel = ctx.dense_stats(words, n_total, rb, re_, ri, 16)
    for i in range(16):
        assert int(sel[i]) == int(sks[i].split(b"\n", 1)[0].split()[2])
        # every record has at least one super-k-mer per w+1 k-mers and at most one per k-mer
        kmers = 5_000_000 - k + 1
        assert kmers // (k - m + 2) <= int(tot[i]) <= kmers
    ctx.close()


def test_dense_machine_matches_reference_print_stat():
    """total_superkmer_number / selected k-mers of the dense kernels against the numbers the reference
    binary printed (tests/golden/stats.json), one single-input batch per golden sketch case."""
    import json
    from tests.conftest import GOLDEN_DIR
    with open(os.path.join(GOLDEN_DIR, "stats.json")) as f:
        stats = json.load(f)
    checked = 0
    for name, want in sorted(stats.items()):
        inp, k, m, s, a = SKETCH_CASES[name]
        if want.get("none_selected"):
            continue
        words, nb, rb, re_, ri = _batch([inp], k)
        ctx = S.DeviceContext(k, m, S.threshold(k, m, s))
        tot, sel = ctx.dense_stats(words, nb, rb, re_, ri, 1)
        ctx.close()
        assert int(tot[0]) == want["total_superkmers"], (name, int(tot[0]), want["total_superkmers"])
        assert int(sel[0]) == want["selected_kmers"], (name, int(sel[0]), want["selected_kmers"])
        checked += 1
    assert checked >= 25


def test_cli_print_stat_totals(tmp_path):
    """sub_sampler -v 1 prints the reference's totals (kmers / superkmers seen) for -i and -f."""
    import json
    from tests.conftest import GOLDEN_DIR
    with open(os.path.join(GOLDEN_DIR, "stats.json")) as f:
        stats = json.load(f)
    exe_s = os.path.join(capi.BIN_DIR, "sub_sampler")
    name = "multi_k31_m11_s20"
    inp, k, m, s, a = SKETCH_CASES[name]
    p = tmp_path / (inp + ".fa")
    p.write_bytes(build_input(inp))
    want = stats[name]

    def commas(n):
        return f"{n:,}"
    line_k = f"I have seen {commas(want['total_kmers'])} kmers and I selected {commas(want['selected_kmers'])} kmers"
    r = subprocess.run([exe_s, "-i", str(p), "-k", str(k), "-m", str(m), "-s", str(s)], cwd=tmp_path,
                       stdin=subprocess.DEVNULL, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert line_k in r.stdout
    assert (f"I have seen {commas(want['total_superkmers'])} superkmers and I selected "
            f"{commas(want['selected_superkmers'])} superkmers") in r.stdout
    fof = tmp_path / "in.txt"
    fof.write_text(str(p) + "\n")
    r = subprocess.run([exe_s, "-f", str(fof), "-k", str(k), "-m", str(m), "-s", str(s), "-t", "2"], cwd=tmp_path,
                       stdin=subprocess.DEVNULL, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert line_k in r.stdout
    assert f"I have seen {commas(want['total_superkmers'])} superkmers" in r.stdout


# ---------------------------------------------------------------- BASELINE config shapes (scaled down)

def test_config5_query_mode_m13_s200(oracle):
    """C5 shape: N-vs-All query mode at k31 m13 s200 -- 6 queries against 70 references (3 column tiles),
    sketches from the batch pipeline, compare from the device-resident elements; CSV text at -p 6."""
    fas = [synth.fasta_bytes([(nm, g)]) for nm, g in synth.genome_family(76, 120_000, seed=21)]
    k, m, s = 31, 13, 200
    pl = S.Pipeline(k, m, s, threads=8)
    sks = pl.sketch(fas)
    for i in (0, 5, 40, 75):
        assert sks[i] == oracle.sketch(fas[i], k, m, s)[0]
    q = 6
    inter, sizes, full = pl.compare(q)
    assert full and inter.shape == (q, len(fas))
    o_inter, o_sizes, _, _ = oracle.compare(sks, q)
    assert np.array_equal(sizes, o_sizes)
    # the reference fills pairs that involve a query; row i of the query block
    sym = o_inter + o_inter.T
    got = inter.copy()
    np.fill_diagonal(got[:, :q], 0)
    assert np.array_equal(got, sym[:q])
    names = [f"g{i}.gz" for i in range(len(fas))]
    for jac in (False, True):
        assert S.format_csv(names, q, inter, full, sizes, jac, 6, 0.0) == oracle.csv(names, q, o_inter, o_sizes, jac, 6, 0.0)
    pl.close()


def test_config4_read_sets(oracle):
    """C4 shape: 150 bp reads (one record each, both strands), several read sets in one batch: short-record
    boundaries everywhere, uint8 counts well above 1, dense totals per read set."""
    g = synth.random_genome(120_000, 77)
    sets = [synth.reads_fasta_bytes(synth.read_set(8_000, 150, g, 100 + i)) for i in range(4)]
    sets.append(synth.reads_fasta_bytes(synth.read_set(300, 40, g, 9)))          # reads of k+9 bases
    k, m, s = 31, 11, 100
    pl = S.Pipeline(k, m, s, threads=4)
    sks = pl.sketch(sets)
    want = [oracle.sketch(x, k, m, s) for x in sets]
    assert sks == [w[0] for w in want]
    inter, sizes, _ = pl.compare()
    o_inter, o_sizes, _, _ = oracle.compare(sks)
    assert np.array_equal(sizes, o_sizes) and np.array_equal(np.triu(inter, 1), np.triu(o_inter, 1))
    pl.close()
    ws, ros = zip(*[(lambda r: (r[0], r[2]))(S.pack_fasta(x, k)) for x in sets])
    words, nb, rb, re_, ri = S.batch_layout(list(ws), list(ros))
    ctx = S.DeviceContext(k, m, S.threshold(k, m, s))
    tot, sel = ctx.dense_stats(words, nb, rb, re_, ri, len(sets))
    for i, (_, st) in enumerate(want):
        assert int(tot[i]) == st["total_superkmers"] and int(sel[i]) == st["selected_kmers"]
    ctx.close()


# ---------------------------------------------------------------- randomized inputs

def _random_fasta(rng, n_records):
    """Adversarial-ish FASTA: random line widths, CRLF, N runs, lower case, tandem repeats, palindromes,
    short records, empty lines, records without a trailing newline."""
    out = bytearray()
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    for r in range(n_records):
        kind = int(rng.integers(0, 6))
        L = int(rng.integers(0, 1500))
        seq = bytes(rng.choice(list(b"ACGT"), size=L).astype(np.uint8))
        if kind == 1 and L > 20:                         # tandem repeat
            u = seq[: int(rng.integers(1, 12))]
            seq = (u * (L // len(u) + 1))[:L]
        elif kind == 2 and L > 20:                       # u rc(u) u rc(u)
            u = seq[: L // 4]
            seq = (u + u.translate(comp)[::-1]) * 2
        elif kind == 3:                                  # junk characters that clean_dna deletes
            seq = bytearray(seq)
            for _ in range(int(rng.integers(0, 8))):
                p = int(rng.integers(0, max(1, len(seq))))
                seq[p:p] = bytes(rng.choice(list(b"NnRYKM-*>x "), size=int(rng.integers(1, 15))).astype(np.uint8))
            seq = bytes(seq)
        elif kind == 4:
            seq = seq.lower()
        width = int(rng.integers(1, 120))
        nl = b"\r\n" if rng.integers(0, 4) == 0 else b"\n"
        out += b">rec%d some text ACGT\n" % r if rng.integers(0, 10) else b">\n"
        for i in range(0, len(seq), width):
            line = seq[i:i + width]
            if line.startswith(b">"):                    # a sequence line must not look like a header
                line = b"N" + line
            out += line + nl
        if rng.integers(0, 6) == 0:
            out += nl
    if rng.integers(0, 2):
        out = out.rstrip(b"\r\n")
    return bytes(out)


@pytest.mark.parametrize("seed", range(8))
def test_random_inputs_all_routes_agree(seed, oracle):
    """Random (k, m, s, a) and random messy inputs: batch pipeline == per-file route == oracle (bytes),
    compare counts == oracle, dense totals == oracle."""
    rng = np.random.default_rng(1000 + seed)
    m = int(rng.choice([3, 5, 7, 9, 11, 13, 15]))
    k = int(rng.choice([x for x in (15, 17, 21, 27, 31, 33, 41, 63) if x > m + 1]))
    s = float(rng.choice([1, 1.5, 2, 3, 5, 10, 30, 100]))
    a = int(rng.choice([1, 1, 1, 2]))
    fas = [_random_fasta(rng, int(rng.integers(1, 40))) for _ in range(int(rng.integers(1, 7)))]
    want = [oracle.sketch(f, k, m, s, a) for f in fas]
    pl = S.Pipeline(k, m, s, a, threads=3)
    got = pl.sketch(fas)
    assert got == [w[0] for w in want], (k, m, s, a)
    assert S.sketch_buffers(fas, k, m, s, a, threads=2) == got
    inter, sizes, _ = pl.compare()
    o_inter, o_sizes, _, _ = oracle.compare(got)
    assert np.array_equal(sizes, o_sizes)
    assert np.array_equal(np.triu(inter, 1), np.triu(o_inter, 1))
    ws, ros = zip(*[(lambda r: (r[0], r[2]))(S.pack_fasta(f, k)) for f in fas])
    words, nb, rb, re_, ri = S.batch_layout(list(ws), list(ros))
    tot, sel = pl.device_context().dense_stats(words, nb, rb, re_, ri, len(fas))
    for i, (_, st) in enumerate(want):
        assert int(tot[i]) == st["total_superkmers"], (k, m, s, i)
        assert int(sel[i]) == st["selected_kmers"]
    pl.close()


def test_batch_stream_overlapped_jobs(oracle):
    """BatchStream: pack of job i+1 overlaps the device phase + compare of job i on a second pipeline; every
    job's sketches and counts equal the plain Pipeline's (and the oracle's)."""
    k, m, s = 31, 11, 50
    jobs = []
    for j in range(5):
        fam = [synth.fasta_bytes([(nm, g)]) for nm, g in synth.genome_family(5 + j, 40_000 + 7_000 * j, seed=30 + j)]
        if j == 2:
            fam.append(build_input("nasty"))
        jobs.append(fam)
    stream = S.BatchStream(k, m, s, threads=4)
    results = []
    for job in jobs:
        done = stream.submit(job)
        if done is not None:
            results.append(done)
    results.append(stream.drain())
    assert stream.drain() is None
    stream.close()
    assert len(results) == len(jobs)
    pl = S.Pipeline(k, m, s, threads=2)
    for job, (sks, (inter, sizes, full), info, cinfo) in zip(jobs, results):
        assert sks == pl.sketch(job)
        assert sks[0] == oracle.sketch(job[0], k, m, s)[0]
        i2, s2, f2 = pl.compare()
        assert np.array_equal(inter, i2) and np.array_equal(sizes, s2) and full == f2
        assert info["launches"] > 0 and cinfo["launches"] > 0
    pl.close()
