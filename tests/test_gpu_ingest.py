"""GPU parity of the device-side FASTA ingest (csrc/device/ingest.cu, spsp_batch_text_*): raw text on the
device -> cleaned 2-bit regions + record table, against the oracle's restatement of getLineFasta + clean_dna
(utils.cpp:706-718, 675-702), and the three ingest modes of the batch pipeline against each other, the oracle
and the reference goldens."""
import gzip

import numpy as np
import pytest

import supersampler_b200 as S
from supersampler_b200 import synth
from tests.golden_inputs import SKETCH_CASES, build_input
from tests.test_gpu_parity import _golden_groups, _random_fasta, sha, unpack

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _built():
    S.build()


def _check_ingest(ctx, texts, oracle):
    nb, words, region, (rb, re_, ri), n_total = ctx.ingest_texts(texts)
    assert np.all(np.diff(rb.astype(np.int64)) >= 0) and np.all(np.diff(ri.astype(np.int64)) >= 0)
    for i, t in enumerate(texts):
        bases, offs = oracle.clean(t)
        assert int(nb[i]) == bases.size, (i, int(nb[i]), bases.size)
        assert np.array_equal(unpack(words[i], bases.size), bases), i
        # zero padding after the last base (the scan reads it)
        tail = unpack(words[i], words[i].size * 16)[bases.size:]
        assert np.all(tail == ord("A"))
        # records: every non-empty record of the oracle, in order, in batch coordinates
        mine = [(int(b), int(e)) for b, e, x in zip(rb, re_, ri) if x == i and e > b]
        want = [(int(region[i]) + int(offs[r]), int(region[i]) + int(offs[r + 1])) for r in range(offs.size - 1)
                if offs[r + 1] > offs[r]]
        assert mine == want, i
        sel = ri == i
        assert np.all(re_[sel] >= rb[sel]) and (not sel.any() or int(re_[sel][-1]) == int(region[i]) + bases.size)


def test_ingest_edge_cases(oracle):
    ctx = S.DeviceContext(31, 11, S.threshold(31, 11, 10.0))
    texts = [
        b"", b"\n", b">", b">x", b">x\n", b">x\nACGT", b">x\nACGT\n", b"ACGT\nACGT\n", b"\n\nACGT\n>\n>\nAC\n\nGT\n>y",
        b">a\r\nACGT\r\nNNNN\r\nacgt\r\n>b\r\n", b">x\n>ACGT\nTTTT\n", b">x\nAC>GT\n", b">x\nAC\n >not a header\nGT\n",
        b"\xff\xfe\n\xffACGT\nAC\x00GT\n", b"A" * 100, b">x\n" + b"A" * 16381, b">x\n" + b"A" * 16382 + b"\n>y\nCC",
        (b">r\n" + b"ACGTTGCA" * 20 + b"\n") * 300,
    ]
    _check_ingest(ctx, texts, oracle)
    # tile boundaries (16 KB): a line end, a header start and a header body exactly at / across the boundary
    for pad in (16379, 16380, 16381, 16382, 16383, 16384, 16385):
        t = b">h\n" + b"C" * (pad - 3) + b"\n>second header that crosses " + b"x" * 40 + b"\nGATTACA\n"
        t2 = b">h\n" + b"G" * (pad - 4) + b"\n" + b">" + b"\nTT\n"
        _check_ingest(ctx, [t, t2, t[:pad], t[:pad + 1]], oracle)
    ctx.close()


@pytest.mark.parametrize("seed", range(4))
def test_ingest_random_texts(seed, oracle):
    rng = np.random.default_rng(500 + seed)
    ctx = S.DeviceContext(31, 11, S.threshold(31, 11, 10.0))
    texts = [_random_fasta(rng, int(rng.integers(1, 120))) for _ in range(int(rng.integers(1, 9)))]
    # a long header line and a long single-line sequence (several tiles without a line start)
    texts.append(b">" + b"ACGT" * 12000 + b"\n" + bytes(rng.choice(list(b"ACGTNacgt\r"), size=70_000).astype(np.uint8)) + b"\n>z\nACGT")
    _check_ingest(ctx, texts, oracle)
    ctx.close()


def test_ingest_then_sketch_staged(oracle):
    """The C ABI route by hand: ingest on the device, then spsp_sketch_batch_staged with no host records."""
    k, m, s = 31, 11, 20.0
    texts = [build_input("multi"), build_input("nasty"), build_input("reads"), b"", build_input("noheader")]
    ctx = S.DeviceContext(k, m, S.threshold(k, m, s))
    _, _, _, _, n_total = ctx.ingest_texts(texts)
    got = ctx.sketch_staged(n_total, [], [], [], len(texts), s)
    assert got == [oracle.sketch(t, k, m, s)[0] for t in texts]
    ctx.close()


@pytest.mark.parametrize("mode", ["device", "auto"])
def test_pipeline_ingest_modes_match_goldens(mode, golden):
    """Every golden sketch case through the batch pipeline with the text cleaned + packed on the device."""
    for (k, m, s, a), cases in _golden_groups().items():
        pl = S.Pipeline(k, m, s, a, threads=4, ingest=mode)
        info = {}
        sks = pl.sketch([build_input(inp) for _, inp in cases], info=info)
        pl.close()
        if mode == "device":
            assert info["text_inputs"] == len(cases)
        for (name, _), sk in zip(cases, sks):
            assert sha(sk) == golden["sketch"][name]["sha256"], (mode, name)


@pytest.mark.parametrize("seed", range(4))
def test_pipeline_ingest_random_all_modes(seed, oracle, tmp_path):
    """Random parameters and messy inputs (memory, pinned memory, plain files, gzip files): host, device and mixed
    ingest give the oracle's bytes and the same compare counts."""
    rng = np.random.default_rng(2000 + seed)
    m = int(rng.choice([5, 7, 9, 11, 13, 15]))
    k = int(rng.choice([x for x in (17, 21, 31, 33, 63) if x > m + 1]))
    s = float(rng.choice([1, 2, 5, 10, 30, 100]))
    fas = [_random_fasta(rng, int(rng.integers(1, 60))) for _ in range(int(rng.integers(3, 12)))]
    want = [oracle.sketch(f, k, m, s)[0] for f in fas]
    inputs = []
    for i, f in enumerate(fas):
        if i % 4 == 1:
            p = tmp_path / f"in{i}.fa"
            p.write_bytes(f)
            inputs.append(str(p))
        elif i % 4 == 2:
            p = tmp_path / f"in{i}.fa.gz"
            with gzip.open(p, "wb", compresslevel=1) as g:
                g.write(f)
            inputs.append(str(p))
        elif i % 4 == 3:
            inputs.append(S.PinnedBuffer(f))
        else:
            inputs.append(f)
    ref = None
    for mode in ("host", "device", "auto"):
        pl = S.Pipeline(k, m, s, threads=3, ingest=mode)
        got = pl.sketch(inputs)
        assert got == want, (mode, k, m, s)
        inter, sizes, _ = pl.compare()
        if ref is None:
            ref = (inter, sizes)
        assert np.array_equal(inter, ref[0]) and np.array_equal(sizes, ref[1])
        # several batches (records tables are merged per batch)
        pl2 = S.Pipeline(k, m, s, threads=2, ingest=mode, max_batch_bases=1 << 16)
        assert pl2.sketch(inputs) == want
        pl.close(); pl2.close()


def test_pipeline_ingest_genomes_and_reads(oracle):
    """Whole genomes (80-column lines) and a read set of short records (C4 shape, one header per 150 bases)."""
    k, m, s = 31, 11, 1000.0
    fam = [synth.fasta_bytes([(nm, g)]) for nm, g in synth.genome_family(6, 1_500_000, seed=77)]
    g = synth.random_genome(400_000, 5)
    reads = synth.reads_fasta_bytes(synth.read_set(60_000, 150, g, 6))
    inputs = fam[:3] + [reads] + fam[3:]
    host = S.Pipeline(k, m, s, threads=4, ingest="host")
    want = host.sketch(inputs)
    assert want[3] == oracle.sketch(reads, k, m, s)[0]
    i0, s0, _ = host.compare()
    for mode in ("device", "auto"):
        pl = S.Pipeline(k, m, s, threads=4, ingest=mode)
        info = {}
        assert pl.sketch(inputs, info=info) == want, mode
        i1, s1, _ = pl.compare()
        assert np.array_equal(i0, i1) and np.array_equal(s0, s1)
        assert info["text_inputs"] >= 1 and info["ingest_ms"] > 0
        # overlapped halves (what BatchStream does)
        pins = [S.PinnedBuffer(x) for x in inputs]      # must outlive the job: the copies are asynchronous
        pl.pack(pins)
        assert pl.finish() == want
        del pins
        pl.close()
    host.close()


@pytest.mark.parametrize("mode", ["host", "device", "auto"])
def test_cli_batch_stats_and_sketches_every_ingest_mode(mode, tmp_path, golden):
    """`sub_sampler -f -v 1` on genomes, read sets and messy records in one file of files: the printed totals
    (k-mers / super-k-mers seen: the dense machine on the record table the device used, host-packed, ingested or
    merged) equal the reference's printed numbers and every sketch file holds the golden bytes."""
    import json
    import os
    import subprocess
    from supersampler_b200 import capi
    from tests.conftest import GOLDEN_DIR
    with open(os.path.join(GOLDEN_DIR, "stats.json")) as f:
        stats = json.load(f)
    cases = ["c1_k31_m11_s1000", "reads_k31_m11_s1000", "nasty_k31_m11_s1000", "tiny_k31_m11_s1000", "empty_k31_m11_s1000"]
    paths = []
    for c in cases:
        inp = SKETCH_CASES[c][0]
        p = tmp_path / (inp + ".fa")
        p.write_bytes(build_input(inp))
        paths.append(p)
    fof = tmp_path / "in.txt"
    fof.write_text("\n".join(str(p) for p in paths) + "\n")
    env = dict(os.environ, SPSP_INGEST=mode)
    r = subprocess.run([os.path.join(capi.BIN_DIR, "sub_sampler"), "-f", str(fof), "-k", "31", "-m", "11", "-s", "1000", "-t", "3"],
                       cwd=tmp_path, env=env, stdin=subprocess.DEVNULL, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for c in cases:
        inp = SKETCH_CASES[c][0]
        with gzip.open(tmp_path / f"subsampled_{inp}.gz", "rb") as f:
            assert sha(f.read()) == golden["sketch"][c]["sha256"], (mode, c)
        want = stats[c]
        if want.get("selected_kmers"):
            assert f"I have seen {want['total_kmers']:,} kmers and I selected {want['selected_kmers']:,} kmers" in r.stdout, (mode, c)
            assert f"I have seen {want['total_superkmers']:,} superkmers" in r.stdout, (mode, c)
