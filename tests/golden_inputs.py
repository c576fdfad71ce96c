"""Named seeded inputs + the golden case tables shared by tools/make_golden.py
(build container, runs the reference) and the tests (any box)."""
from __future__ import annotations

import functools

import numpy as np

from supersampler_b200 import synth


@functools.lru_cache(maxsize=64)
def build_input(name: str) -> bytes:
    """FASTA bytes of a named input."""
    if name == "c1":                       # BASELINE config 1: one 5 Mbp genome
        return synth.fasta_bytes([("c1", synth.random_genome(5_000_000, 1))])
    if name == "c1mut":                    # ... and its 1 % mutated copy
        g = synth.random_genome(5_000_000, 1)
        return synth.fasta_bytes([("c1mut", synth.mutate(g, 0.01, 2))])
    if name == "nasty":
        return synth.raw_fasta_bytes(synth.nasty_records())
    if name == "nasty21":
        return synth.raw_fasta_bytes(synth.nasty_records(k=21, seed=11))
    if name == "reads":                    # C4 shape, scaled down: 150 bp reads
        g = synth.random_genome(200_000, 3)
        return synth.reads_fasta_bytes(synth.read_set(20_000, 150, g, 4))
    if name == "multi":                    # several records incl. short ones, 70-col lines
        rng = np.random.default_rng(9)
        recs = [(f"r{i}", synth.random_genome(int(rng.integers(10, 5000)), 100 + i)) for i in range(40)]
        return synth.fasta_bytes(recs, width=70)
    if name == "noheader":                 # first line is always dropped (utils.cpp:708)
        return synth.random_genome(500, 21).tobytes() + b"\n" + synth.random_genome(700, 22).tobytes() + b"\n>x\n" \
            + synth.random_genome(300, 23).tobytes()
    if name == "ffline":                   # lines that start with the byte 0xFF end a record like '>' lines do
        g = [synth.random_genome(n, 40 + i).tobytes() for i, n in enumerate((400, 400, 400, 90, 250))]
        return (b">r1\n" + g[0] + b"\n\xff" + g[1] + b"\n" + g[2] + b"\n\xff\n" + g[3][:45] + b"\xff" + g[3][45:] + b"\n>r2\n\xffNNN\n" + g[4])
    if name == "empty":
        return b""
    if name == "tiny":
        return b">x\nACGT\n"
    if name == "wrap256":                  # uint8 count wrap (SubSampler.h:24): 256 copies of one read
        r = synth.random_genome(60, 31).tobytes()
        return b"".join(b">r\n" + r + b"\n" for _ in range(256))
    if name == "wrap257":
        r = synth.random_genome(60, 31).tobytes()
        return b"".join(b">r\n" + r + b"\n" for _ in range(257))
    if name.startswith("fam"):             # famN_i: genome i of a related family (C2 recipe, 300 kbp)
        n, i = name[3:].split("_")
        for j, (nm, g) in enumerate(synth.genome_family(int(n), 300_000, seed=5)):
            if j == int(i):
                return synth.fasta_bytes([(nm, g)])
    raise KeyError(name)


# name -> (input, k, m, s, abundance)
SKETCH_CASES = {
    "c1_k31_m11_s1000": ("c1", 31, 11, 1000, 1),
    "c1_k31_m11_s100": ("c1", 31, 11, 100, 1),
    "c1_k31_m13_s200": ("c1", 31, 13, 200, 1),
    "c1mut_k31_m11_s1000": ("c1mut", 31, 11, 1000, 1),
    "reads_k31_m11_s1000": ("reads", 31, 11, 1000, 1),
    "reads_k31_m11_s50": ("reads", 31, 11, 50, 1),
    "multi_k31_m11_s20": ("multi", 31, 11, 20, 1),
    "multi_k21_m9_s5": ("multi", 21, 9, 5, 1),
    "noheader_k31_m11_s4": ("noheader", 31, 11, 4, 1),
    "ffline_k31_m11_s4": ("ffline", 31, 11, 4, 1),
    "empty_k31_m11_s1000": ("empty", 31, 11, 1000, 1),
    "tiny_k31_m11_s1000": ("tiny", 31, 11, 1000, 1),
    "wrap256_k31_m11_s1": ("wrap256", 31, 11, 1, 1),
    "wrap257_k31_m11_s1": ("wrap257", 31, 11, 1, 1),
    "wrap257_k31_m11_s1_a2": ("wrap257", 31, 11, 1, 2),
}
for _k, _m, _s in [(31, 11, 1000), (31, 11, 100), (31, 11, 10), (31, 11, 2), (31, 11, 1), (21, 9, 5), (21, 9, 50),
                   (15, 5, 3), (31, 15, 20), (63, 15, 10), (41, 13, 7), (31, 13, 200), (31, 11, 1.5), (33, 13, 4),
                   (61, 15, 3), (17, 15, 6)]:
    SKETCH_CASES[f"nasty_k{_k}_m{_m}_s{_s}"] = ("nasty", _k, _m, _s, 1)
SKETCH_CASES["nasty21_k21_m9_s5"] = ("nasty21", 21, 9, 5, 1)

# name -> (inputs (queries first), k, m, s, n_query (0 = all-vs-all), precision, min_threshold)
_FAM = [f"fam12_{i}" for i in range(12)]
COMPARE_CASES = {
    "c1_pair": (["c1", "c1mut"], 31, 11, 1000, 0, 6, 0.0),
    "fam12_s100": (_FAM + ["nasty", "tiny"], 31, 11, 100, 0, 6, 0.0),
    "fam12_m13_s20": (_FAM[:3] + ["tiny"] + _FAM[3:] + ["nasty"], 31, 13, 20, 0, 6, 0.0),
    "fam12_k21_query": (_FAM + ["nasty"], 21, 9, 5, 3, 3, 0.3),
    "fam12_query_p8": (_FAM, 31, 11, 50, 4, 8, 0.0),
}
