"""Seeded synthetic FASTA generators for tests and bench.

Everything is numpy + bytes; no reference code is involved.  The same seeds
always give the same bytes, so golden hashes made from these inputs in the
build container stay valid on the GPU box.
"""
from __future__ import annotations

import gzip
import os
from typing import Iterable, List, Sequence, Tuple

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTacgt", b"TGCAtgca"):
    _COMP[_a] = _b


def random_genome(n: int, seed: int) -> np.ndarray:
    """i.i.d. uniform ACGT genome as an uint8 array of ASCII codes."""
    rng = np.random.default_rng(seed)
    return _ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]


def mutate(genome: np.ndarray, rate: float, seed: int) -> np.ndarray:
    """Bernoulli(rate) positions get a *different* base (substitutions only)."""
    rng = np.random.default_rng(seed)
    out = genome.copy()
    pos = np.flatnonzero(rng.random(genome.size) < rate)
    if pos.size:
        lut = np.zeros(256, dtype=np.uint8)
        lut[_ACGT] = np.arange(4, dtype=np.uint8)
        cur = lut[out[pos]]
        out[pos] = _ACGT[(cur + rng.integers(1, 4, size=pos.size, dtype=np.uint8)) & 3]
    return out


def revcomp(seq: np.ndarray) -> np.ndarray:
    return _COMP[seq[::-1]]


def fasta_bytes(records: Sequence[Tuple[str, np.ndarray]], width: int = 80) -> bytes:
    """FASTA text: '>name' then the sequence folded at `width` (0 = one line)."""
    parts: List[bytes] = []
    for name, seq in records:
        parts.append(b">" + name.encode() + b"\n")
        n = int(seq.size)
        if n == 0:
            continue
        if width <= 0 or n <= width:
            parts.append(seq.tobytes() + b"\n")
            continue
        full = (n // width) * width
        body = np.empty((n // width, width + 1), dtype=np.uint8)
        body[:, :width] = seq[:full].reshape(-1, width)
        body[:, width] = 10
        parts.append(body.tobytes())
        if full < n:
            parts.append(seq[full:].tobytes() + b"\n")
    return b"".join(parts)


def write_fasta(path: str, records: Sequence[Tuple[str, np.ndarray]], width: int = 80,
                gz: bool = False) -> str:
    data = fasta_bytes(records, width)
    if gz:
        with gzip.GzipFile(path, "wb", compresslevel=1, mtime=0) as f:
            f.write(data)
    else:
        with open(path, "wb") as f:
            f.write(data)
    return path


def genome_family(n_genomes: int, n_bases: int, seed: int = 42,
                  lo_exp: float = -3.0, hi_exp: float = -1.0) -> Iterable[Tuple[str, np.ndarray]]:
    """C2/C3/C5 recipe: one ancestor, genome g is a copy with substitution
    rate 10^U(lo_exp, hi_exp) (rate and mutation seeds derive from `seed`)."""
    anc = random_genome(n_bases, seed)
    rates = 10.0 ** np.random.default_rng(seed + 1).uniform(lo_exp, hi_exp, size=n_genomes)
    for g in range(n_genomes):
        yield f"g{g:05d}", mutate(anc, float(rates[g]), seed * 1000003 + g)


def read_set(n_reads: int, read_len: int, genome: np.ndarray, seed: int) -> np.ndarray:
    """C4 recipe: reads sampled uniformly, 50 % reverse-complemented.
    Returns an (n_reads, read_len) uint8 array of ASCII codes."""
    rng = np.random.default_rng(seed)
    starts = rng.integers(0, genome.size - read_len + 1, size=n_reads)
    idx = starts[:, None] + np.arange(read_len)[None, :]
    reads = genome[idx]
    flip = rng.random(n_reads) < 0.5
    reads[flip] = _COMP[reads[flip][:, ::-1]]
    return reads


def reads_fasta_bytes(reads: np.ndarray) -> bytes:
    """2-line FASTA ('>r' header, one line per read)."""
    n, L = reads.shape
    out = np.empty((n, L + 4), dtype=np.uint8)
    out[:, 0] = ord(">")
    out[:, 1] = ord("r")
    out[:, 2] = 10
    out[:, 3:3 + L] = reads
    out[:, 3 + L] = 10
    return out.tobytes()


def nasty_records(k: int = 31, seed: int = 7) -> List[Tuple[str, bytes]]:
    """Adversarial records: repeats, palindromes, N runs, lower case, IUPAC,
    records of length k-1 / k / k+1 / 0.  Returned as raw byte strings so they
    may contain characters the cleaner must delete."""
    rng = np.random.default_rng(seed)

    def rnd(n: int) -> bytes:
        return _ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)].tobytes()

    def rc(b: bytes) -> bytes:
        return _COMP[np.frombuffer(b, dtype=np.uint8)[::-1]].tobytes()

    recs: List[Tuple[str, bytes]] = []
    recs.append(("homoA", b"A" * 200))
    recs.append(("homoT", b"T" * 137))
    recs.append(("AC", b"AC" * 150))
    recs.append(("ACG", b"ACG" * 100))
    recs.append(("AACTAACTA", b"AACTAACTA" * 30))
    u = rnd(40)
    recs.append(("pal4", u + rc(u) + u + rc(u)))
    for i in range(300):
        ul = int(rng.integers(5, 26))
        yl = int(rng.integers(0, 13))
        u = rnd(ul)
        recs.append((f"inv{i}", rnd(int(rng.integers(0, 60))) + u + rnd(yl) + rc(u)
                     + rnd(int(rng.integers(0, 60)))))
    for i in range(300):
        unit = rnd(int(rng.integers(1, 24)))
        reps = int(rng.integers(2, 40))
        recs.append((f"tr{i}", rnd(int(rng.integers(0, 40))) + unit * reps
                     + rnd(int(rng.integers(0, 40)))))
    recs.append(("rand20k", rnd(20000)))
    recs.append(("withN", rnd(300) + b"N" * 10 + rnd(300)))
    recs.append(("lower", rnd(100) + rnd(120).lower() + rnd(100)))
    recs.append(("iupac", rnd(90) + b"RYKMSWBDHVN" + rnd(90) + b"-*" + rnd(50)))
    recs.append(("km1", rnd(k - 1)))
    recs.append(("k", rnd(k)))
    recs.append(("kp1", rnd(k + 1)))
    recs.append(("empty", b""))
    recs.append(("allN", b"N" * 50))
    recs.append(("crlf", rnd(100) + b"\r"))
    dup = rnd(60)
    for i in range(5):
        recs.append((f"dup{i}", dup))
    return recs


def raw_fasta_bytes(records: Sequence[Tuple[str, bytes]], width: int = 60) -> bytes:
    parts: List[bytes] = []
    for name, seq in records:
        parts.append(b">" + name.encode() + b"\n")
        for i in range(0, len(seq), width):
            parts.append(seq[i:i + width] + b"\n")
    return b"".join(parts)


def ensure_dir(path: str) -> str:
    os.makedirs(path, exist_ok=True)
    return path


class Family:
    """Random access to the members of a genome family (same recipe as
    genome_family, but the mutation rates come from a fixed pool so that member
    idx is the same genome whatever the number of ranks)."""

    def __init__(self, n_bases: int, seed: int = 42, pool: int = 16384, lo_exp: float = -3.0, hi_exp: float = -1.0):
        self.anc = random_genome(n_bases, seed)
        self.rates = 10.0 ** np.random.default_rng(seed + 1).uniform(lo_exp, hi_exp, size=pool)
        self.seed = seed

    def member(self, idx: int) -> np.ndarray:
        return mutate(self.anc, float(self.rates[idx % self.rates.size]), self.seed * 1000003 + idx)

    def fasta(self, idx: int, width: int = 80) -> bytes:
        return fasta_bytes([(f"g{idx:05d}", self.member(idx))], width)
