"""Synthetic inputs of the BASELINE configs generated ON the GPU (torch ops: plumbing, not product).

The large configs (C3: 1 024 x 5 Mbp, C4: 1 Gbp read sets, C5: 10 100 x 5 Mbp) are too big to
synthesise with numpy inside a benchmark run, so the genomes are drawn on the device, packed
there into the 2-bit layout the scan kernels read (include/spsp.h), and only the subset the
reference has to see is turned into FASTA text and copied to the host.

Recipes are those of SURVEY.md App. D (one ancestor, member g = copy with substitution rate
10^U(-3,-1); reads sampled uniformly, half reverse-complemented); the random streams are
torch's, so the bytes differ from synth.py's numpy genomes of the same seed.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch


def words_per_input(n_bases: int) -> int:
    """spsp_packed_words (include/spsp.h): every input region starts on a 64-base boundary."""
    return 4 * ((n_bases + 4 + 63) // 64) + 8


_W16 = None


def pack_codes(codes: torch.Tensor) -> torch.Tensor:
    """uint8 codes (A0 C1 T2 G3), last dim a multiple of 16 -> int32 words, first base in the MSBs."""
    global _W16
    if _W16 is None or _W16.device != codes.device:
        _W16 = (4 ** torch.arange(15, -1, -1, dtype=torch.int64, device=codes.device))
    shp = codes.shape[:-1] + (codes.shape[-1] // 16, 16)
    w = (codes.reshape(shp).to(torch.int64) * _W16).sum(-1)
    return torch.where(w >= 2 ** 31, w - 2 ** 32, w).to(torch.int32)


_LUT = None


def codes_to_ascii(codes: torch.Tensor) -> torch.Tensor:
    global _LUT
    if _LUT is None or _LUT.device != codes.device:
        _LUT = torch.tensor(list(b"ACTG"), dtype=torch.uint8, device=codes.device)
    return _LUT[codes.long()]


class DeviceFamily:
    """Genome family on the device: member(i) = ancestor with Bernoulli(rate_i) substitutions."""

    def __init__(self, n_bases: int, seed: int = 42, pool: int = 16384, device="cuda"):
        self.n = n_bases
        self.dev = torch.device(device)
        g = torch.Generator(device=self.dev)
        g.manual_seed(seed)
        self.anc = torch.randint(0, 4, (n_bases,), dtype=torch.uint8, device=self.dev, generator=g)
        self.rates = torch.from_numpy(10.0 ** np.random.default_rng(seed + 1).uniform(-3.0, -1.0, size=pool)).to(
            self.dev, torch.float32)
        self.seed = seed

    def codes(self, first: int, count: int) -> torch.Tensor:
        """[count, n] uint8 codes of members first .. first+count-1 (each member has its own stream)."""
        out = torch.empty((count, self.n), dtype=torch.uint8, device=self.dev)
        g = torch.Generator(device=self.dev)
        for j in range(count):
            idx = first + j
            g.manual_seed(self.seed * 1000003 + idx)
            hit = torch.rand(self.n, device=self.dev, generator=g) < self.rates[idx % self.rates.numel()]
            sub = torch.randint(1, 4, (self.n,), dtype=torch.uint8, device=self.dev, generator=g)
            out[j] = (self.anc + hit.to(torch.uint8) * sub) & 3
        return out

    def packed_batch(self, first: int, count: int) -> Tuple[torch.Tensor, int, np.ndarray, np.ndarray, np.ndarray]:
        """Members packed back to back as one scan batch:
        (int32 words on the device, n_bases of the batch, rec_begin, rec_end, rec_input)."""
        wpi = words_per_input(self.n)
        buf = torch.zeros(count * wpi + 64, dtype=torch.int32, device=self.dev)
        n16 = (self.n + 15) // 16
        view = buf[: count * wpi].view(count, wpi)
        step = 32
        for a in range(0, count, step):
            c = self.codes(first + a, min(step, count - a))
            if self.n % 16:
                c = torch.nn.functional.pad(c, (0, 16 - self.n % 16))
            view[a:a + c.shape[0], :n16] = pack_codes(c)
        begins = np.arange(count, dtype=np.uint64) * np.uint64(wpi * 16)
        return buf, count * wpi * 16, begins, begins + np.uint64(self.n), np.arange(count, dtype=np.uint32)

    def fasta(self, first: int, count: int, width: int = 80) -> List[bytes]:
        """FASTA text (host bytes) of members first .. first+count-1, lines of `width` bases."""
        out = []
        c = codes_to_ascii(self.codes(first, count))
        full = (self.n // width) * width
        body = torch.empty((count, self.n // width, width + 1), dtype=torch.uint8, device=self.dev)
        body[:, :, :width] = c[:, :full].view(count, -1, width)
        body[:, :, width] = 10
        body = body.view(count, -1).cpu().numpy()
        tail = c[:, full:].cpu().numpy() if full < self.n else None
        for j in range(count):
            parts = [b">g%05d\n" % (first + j), body[j].tobytes()]
            if tail is not None:
                parts.append(tail[j].tobytes() + b"\n")
            out.append(b"".join(parts))
        return out


class DeviceReadSet:
    """C4: reads of `read_len` sampled uniformly from a random genome, half reverse-complemented."""

    def __init__(self, genome_bases: int = 50_000_000, read_len: int = 150, seed: int = 7, device="cuda"):
        self.dev = torch.device(device)
        g = torch.Generator(device=self.dev)
        g.manual_seed(seed)
        self.genome = torch.randint(0, 4, (genome_bases,), dtype=torch.uint8, device=self.dev, generator=g)
        self.L = read_len
        self.seed = seed

    def codes(self, set_idx: int, n_reads: int) -> torch.Tensor:
        """[n_reads, L] uint8 codes of read set `set_idx`."""
        g = torch.Generator(device=self.dev)
        g.manual_seed(self.seed * 7919 + 1 + set_idx)
        out = torch.empty((n_reads, self.L), dtype=torch.uint8, device=self.dev)
        ar = torch.arange(self.L, device=self.dev)
        step = 1 << 20
        for a in range(0, n_reads, step):
            n = min(step, n_reads - a)
            starts = torch.randint(0, self.genome.numel() - self.L + 1, (n,), device=self.dev, generator=g)
            flip = torch.rand(n, device=self.dev, generator=g) < 0.5
            r = self.genome[starts[:, None] + ar[None, :]]
            rc = torch.flip(r, dims=(1,)) ^ 2
            out[a:a + n] = torch.where(flip[:, None], rc, r)
        return out

    def packed(self, codes: torch.Tensor):
        """One input = all reads back to back (one record per read)."""
        n_reads, L = codes.shape
        nb = n_reads * L
        wpi = words_per_input(nb)
        buf = torch.zeros(wpi + 64, dtype=torch.int32, device=self.dev)
        flat = codes.reshape(-1)
        pad = (-nb) % 16
        step = 1 << 28
        for a in range(0, nb, step):
            piece = flat[a:a + step]
            if piece.numel() % 16:
                piece = torch.nn.functional.pad(piece, (0, pad))
            buf[a // 16: a // 16 + piece.numel() // 16] = pack_codes(piece)
        begins = torch.arange(n_reads, dtype=torch.int64, device=self.dev) * L
        return buf, wpi * 16, begins, begins + L, torch.zeros(n_reads, dtype=torch.int32, device=self.dev)

    def fasta(self, codes: torch.Tensor) -> bytes:
        """2-line FASTA ('>r' header, one line per read), host bytes."""
        n, L = codes.shape
        out = torch.empty((n, L + 4), dtype=torch.uint8, device=self.dev)
        out[:, 0] = ord(">"); out[:, 1] = ord("r"); out[:, 2] = 10; out[:, 3 + L] = 10
        step = 1 << 21
        for a in range(0, n, step):
            out[a:a + step, 3:3 + L] = codes_to_ascii(codes[a:a + step])
        return out.cpu().numpy().tobytes()
