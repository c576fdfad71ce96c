"""ctypes front-end of oracle/spsp_oracle.c and of the compiled reference
binaries in oracle/_ref.

TEST INFRASTRUCTURE ONLY: import this from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / reference arm -- never from supersampler_b200.
"""
from __future__ import annotations

import ctypes as C
import gzip
import os
import subprocess
import tempfile
from typing import List, Optional, Sequence, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE_SRC = "/root/reference"


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "records", "bases", "total_kmers", "total_superkmers", "selected_kmers",
        "selected_superkmers", "maximal_superkmers", "buckets", "distinct_kmers",
        "out_superkmers", "out_maximal", "pb_events")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def build(with_ref: bool = True) -> None:
    """Compile the C restatement; compile oracle/_ref too when the reference
    sources are mounted (build container only)."""
    src = os.path.join(HERE, "spsp_oracle.c")
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "oracle"], check=True, stdout=subprocess.DEVNULL)
    if with_ref and os.path.isdir(REFERENCE_SRC) and not have_ref():
        subprocess.run(["make", "-C", HERE, "all", "-j4"], check=True, stdout=subprocess.DEVNULL)


def have_ref() -> bool:
    return all(os.access(os.path.join(REF_DIR, b), os.X_OK) for b in ("sub_sampler", "comparator"))


_lib = None


def lib():
    global _lib
    if _lib is None:
        build(with_ref=False)
        L = C.CDLL(SO)
        L.spo_hash.restype = C.c_uint64
        L.spo_hash.argtypes = [C.c_uint64]
        L.spo_threshold.restype = C.c_uint64
        L.spo_threshold.argtypes = [C.c_uint, C.c_uint, C.c_double]
        L.spo_sketch.restype = C.c_int
        L.spo_sketch.argtypes = [C.c_char_p, C.c_size_t, C.c_uint, C.c_uint, C.c_double, C.c_uint,
                                 C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(Stats),
                                 C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.spo_clean.restype = C.c_int
        L.spo_clean.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                C.POINTER(C.c_size_t)]
        L.spo_hits.restype = C.c_size_t
        L.spo_hits.argtypes = [C.c_void_p, C.c_size_t, C.c_uint, C.c_uint64, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_size_t]
        L.spo_compare.restype = C.c_int
        L.spo_compare.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_uint, C.c_uint,
                                  C.c_void_p, C.c_void_p, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
        L.spo_csv.restype = C.c_int
        L.spo_csv.argtypes = [C.POINTER(C.c_char_p), C.c_uint, C.c_uint, C.c_void_p, C.c_void_p, C.c_int,
                              C.c_uint, C.c_double, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.spo_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def xxh64_8(x: int) -> int:
    return int(lib().spo_hash(x))


def threshold(k: int, m: int, s: float) -> int:
    return int(lib().spo_threshold(k, m, float(np.float32(s))))


def _take(ptr: C.c_void_p, n: int) -> bytes:
    data = C.string_at(ptr, n) if n else b""
    lib().spo_free(ptr)
    return data


def sketch(fasta: bytes, k: int = 31, m: int = 11, s: float = 1000.0, abundance: int = 1,
           trace: bool = False):
    """FASTA text -> (sketch bytes before gzip, stats dict[, selected (rec,start) array])."""
    L = lib()
    out, n = C.c_void_p(), C.c_size_t()
    st = Stats()
    sel, nsel = C.c_void_p(), C.c_size_t()
    rc = L.spo_sketch(fasta, len(fasta), k, m, float(np.float32(s)), abundance, C.byref(out), C.byref(n),
                      C.byref(st), C.byref(sel) if trace else None, C.byref(nsel) if trace else None)
    if rc != 0:
        raise ValueError("oracle: bad parameters")
    data = _take(out, n.value)
    if not trace:
        return data, st.as_dict()
    arr = np.frombuffer(C.string_at(sel, nsel.value * 16), dtype=np.uint64).reshape(-1, 2).copy() \
        if nsel.value else np.zeros((0, 2), np.uint64)
    L.spo_free(sel)
    return data, st.as_dict(), arr


def clean(fasta: bytes) -> Tuple[np.ndarray, np.ndarray]:
    """FASTA text -> (cleaned bases of all records concatenated, offsets[n_rec+1])."""
    L = lib()
    b, o, n = C.c_void_p(), C.c_void_p(), C.c_size_t()
    L.spo_clean(fasta, len(fasta), C.byref(b), C.byref(o), C.byref(n))
    offs = np.frombuffer(C.string_at(o, (n.value + 1) * 8), dtype=np.uint64).copy()
    bases = np.frombuffer(C.string_at(b, int(offs[-1])), dtype=np.uint8).copy()
    L.spo_free(b)
    L.spo_free(o)
    return bases, offs


def hits(seq: np.ndarray, m: int, thr: int):
    """Closed-form hit list of one cleaned record: (pos u64, canon u32, rev u8, hash u64)."""
    L = lib()
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    cap = max(1024, seq.size // 64)
    while True:
        pos = np.empty(cap, np.uint64); cn = np.empty(cap, np.uint32)
        rv = np.empty(cap, np.uint8); hs = np.empty(cap, np.uint64)
        n = L.spo_hits(seq.ctypes.data, seq.size, m, thr, pos.ctypes.data, cn.ctypes.data,
                       rv.ctypes.data, hs.ctypes.data, cap)
        if n <= cap:
            return pos[:n], cn[:n], rv[:n], hs[:n]
        cap = int(n)


def compare(sketches: Sequence[bytes], query_size: Optional[int] = None):
    """Sketch bytes -> (inter[n,n] uint32 upper triangle, sizes[n] uint64, k, m)."""
    L = lib()
    n = len(sketches)
    q = n if query_size is None else query_size
    arr = (C.c_char_p * n)(*sketches)
    lens = (C.c_size_t * n)(*[len(s) for s in sketches])
    inter = np.zeros((n, n), np.uint32)
    sizes = np.zeros(n, np.uint64)
    k, m = C.c_uint(), C.c_uint()
    rc = L.spo_compare(arr, lens, n, q, inter.ctypes.data, sizes.ctypes.data, C.byref(k), C.byref(m))
    if rc != 0:
        raise ValueError("oracle: undecodable sketch")
    return inter, sizes, k.value, m.value


def csv(names: Sequence[str], query_size: int, inter: np.ndarray, sizes: np.ndarray, jaccard: bool,
        precision: int = 6, min_threshold: float = 0.0) -> bytes:
    L = lib()
    n = len(names)
    arr = (C.c_char_p * n)(*[s.encode() for s in names])
    inter = np.ascontiguousarray(inter, np.uint32)
    sizes = np.ascontiguousarray(sizes, np.uint64)
    out, ln = C.c_void_p(), C.c_size_t()
    L.spo_csv(arr, n, query_size, inter.ctypes.data, sizes.ctypes.data, int(jaccard), precision,
              min_threshold, C.byref(out), C.byref(ln))
    return _take(out, ln.value)


# ----------------------------------------------------------------- oracle/_ref

def ref_sketch_files(paths: Sequence[str], k=31, m=11, s=1000.0, threads: int = 1, workdir: Optional[str] = None,
                     abundance: int = 1) -> List[bytes]:
    """Run the compiled reference sub_sampler; returns the gunzipped sketch of
    each input, in input order.  Always </dev/null (the reference may cin.get())."""
    own = workdir is None
    wd = workdir or tempfile.mkdtemp(prefix="spsp_ref_")
    exe = os.path.join(REF_DIR, "sub_sampler")
    common = ["-k", str(k), "-m", str(m), "-s", repr(float(s)), "-v", "0", "-a", str(abundance)]
    if len(paths) == 1:
        cmd = [exe, "-i", os.path.abspath(paths[0])] + common
    else:
        fof = os.path.join(wd, "in_fof.txt")
        with open(fof, "w") as f:
            f.write("\n".join(os.path.abspath(p) for p in paths) + "\n")
        cmd = [exe, "-f", fof, "-t", str(threads)] + common
    subprocess.run(cmd, cwd=wd, stdin=subprocess.DEVNULL, stdout=subprocess.DEVNULL, check=True)
    out = []
    for p in paths:
        stem = os.path.basename(p).split(".")[0]
        with gzip.open(os.path.join(wd, "subsampled_" + stem + ".gz"), "rb") as f:
            out.append(f.read())
    if own:
        import shutil
        shutil.rmtree(wd, ignore_errors=True)
    return out


def ref_compare_files(sketch_paths: Sequence[str], query_paths: Sequence[str] = (), precision: int = 6,
                      min_threshold: float = 0.0, workdir: Optional[str] = None):
    """Run the compiled reference comparator on gz sketch files; returns
    (containment csv bytes, jaccard csv bytes, 'Comparisons lasted' seconds or None)."""
    own = workdir is None
    wd = workdir or tempfile.mkdtemp(prefix="spsp_refc_")
    fof = os.path.join(wd, "sk_fof.txt")
    with open(fof, "w") as f:
        f.write("\n".join(sketch_paths) + "\n")
    cmd = [os.path.join(REF_DIR, "comparator"), "-f", fof, "-p", str(precision), "-m", repr(float(min_threshold)),
           "-o", os.path.join(wd, "res")]
    if query_paths:
        qf = os.path.join(wd, "q_fof.txt")
        with open(qf, "w") as f:
            f.write("\n".join(query_paths) + "\n")
        cmd += ["-q", qf]
    r = subprocess.run(cmd, cwd=wd, stdin=subprocess.DEVNULL, stdout=subprocess.PIPE, check=True, text=True)
    secs = None
    for line in r.stdout.splitlines():
        if line.startswith("Comparisons lasted"):
            secs = float(line.split()[2])
    with gzip.open(os.path.join(wd, "res_containment.csv.gz"), "rb") as f:
        cont = f.read()
    with gzip.open(os.path.join(wd, "res_jaccard.csv.gz"), "rb") as f:
        jac = f.read()
    if own:
        import shutil
        shutil.rmtree(wd, ignore_errors=True)
    return cont, jac, secs
