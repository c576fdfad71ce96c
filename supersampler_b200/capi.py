"""ctypes bindings of include/spsp.h (device C ABI) and include/spsp_host.h."""
from __future__ import annotations

import collections.abc
import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(PKG, "lib")
BIN_DIR = os.path.join(PKG, "bin")
CSRC = os.path.join(PKG, "csrc")

HIT_DTYPE = np.dtype([("pos", "<u8"), ("canon", "<u4"), ("rev", "<u4")])

SCAN_AUTO, SCAN_DENSE, SCAN_FILTER = 0, 1, 2


class SpspError(RuntimeError):
    pass


class BatchResult(C.Structure):
    """spsp_batch_result (include/spsp.h)."""
    _fields_ = [("body", C.c_void_p), ("body_off", C.POINTER(C.c_uint64)), ("selected", C.POINTER(C.c_uint64)),
                ("elem_off", C.POINTER(C.c_uint64)), ("n_hits", C.c_uint64), ("n_elems", C.c_uint64),
                ("scan_ms", C.c_float), ("post_ms", C.c_float)]


def batch_layout(packed_list, rec_off_list):
    """Concatenate per-input packed buffers (each spsp_packed_words long, so every
    input starts on a 64-base boundary) -> (words, n_bases, rec_begin, rec_end, rec_input)."""
    base, begins, ends, inputs, off = [], [], [], [], 0
    for i, (w, ro) in enumerate(zip(packed_list, rec_off_list)):
        ro = np.asarray(ro, np.uint64)
        begins.append(ro[:-1] + np.uint64(off)); ends.append(ro[1:] + np.uint64(off))
        inputs.append(np.full(ro.size - 1, i, np.uint32))
        off += int(np.asarray(w).size) * 16
    words = np.concatenate(list(packed_list) + [np.zeros(64, np.uint32)])
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    return words, off, cat(begins, np.uint64), cat(ends, np.uint64), cat(inputs, np.uint32)


def build(force: bool = False) -> None:
    """Compile the CUDA device layer (sm_100a), the host layer and the CLIs in-tree."""
    args = ["make", "-C", CSRC, "all", "-j4"]
    if force:
        subprocess.run(["make", "-C", CSRC, "clean"], check=True, stdout=subprocess.DEVNULL)
    r = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise SpspError("build failed:\n" + r.stdout[-4000:])


_dev = None
_host = None


def _need(path: str) -> str:
    if not os.path.exists(path):
        raise SpspError(f"{path} is missing: run supersampler_b200.build() (needs nvcc); there is no fallback path")
    return path


def device_lib():
    global _dev
    if _dev is None:
        # SPSP_DEVICE_LIB: another build of the device layer (kernel experiments); the default is the in-tree library
        L = C.CDLL(_need(os.environ.get("SPSP_DEVICE_LIB") or os.path.join(LIB_DIR, "libspsp_b200.so")), mode=C.RTLD_GLOBAL)
        L.spsp_last_error.restype = C.c_char_p
        L.spsp_packed_words.restype = C.c_uint64
        L.spsp_packed_words.argtypes = [C.c_uint64]
        L.spsp_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.POINTER(C.c_void_p)]
        L.spsp_destroy.argtypes = [C.c_void_p]
        L.spsp_scan_config.argtypes = [C.c_void_p, C.c_int]
        L.spsp_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        L.spsp_host_free.argtypes = [C.c_void_p]
        L.spsp_scan_submit.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64]
        L.spsp_scan_collect.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.spsp_scan_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]
        L.spsp_sync.argtypes = [C.c_void_p, C.c_int]
        L.spsp_scan_last_kernel_ms.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        L.spsp_stream.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.spsp_cmp_load.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.spsp_cmp_load_device.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.spsp_cmp_run.argtypes = [C.c_void_p] + [C.c_uint32] * 4 + [C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64]
        L.spsp_cmp_run_device.argtypes = L.spsp_cmp_run.argtypes
        L.spsp_cmp_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.spsp_launch_count.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.spsp_scan_filter_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint64)]
        L.spsp_sketch_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_uint64, C.c_uint32, C.c_uint, C.POINTER(BatchResult)]
        L.spsp_sketch_batch_device.argtypes = L.spsp_sketch_batch.argtypes
        L.spsp_cmp_load_batch.argtypes = [C.c_void_p, C.c_int]
        L.spsp_batch_reserve.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
        L.spsp_batch_upload.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_uint64]
        L.spsp_sketch_batch_staged.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_uint64, C.c_uint32, C.c_uint, C.POINTER(BatchResult)]
        L.spsp_batch_text_reserve.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
        L.spsp_batch_text_upload.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_uint64]
        L.spsp_batch_upload_wait.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.spsp_batch_text_pack.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p]
        L.spsp_batch_text_last_ms.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        L.spsp_batch_text_records.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                              C.POINTER(C.c_uint64)]
        L.spsp_batch_download.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_uint64]
        L.spsp_dense_stats.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
        L.spsp_dense_stats_device.argtypes = L.spsp_dense_stats.argtypes
        L.spsp_dense_stats_staged.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
        L.spsp_nccl_unique_id.argtypes = [C.c_void_p]
        L.spsp_nccl_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.spsp_cmp_exchange_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32,
                                              C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
        L.spsp_cmp_exchange.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint32,
                                        C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
        L.spsp_batch_reserve.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
        L.spsp_batch_upload.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_uint64]
        L.spsp_sketch_batch_staged.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_uint64, C.c_uint32, C.c_uint, C.POINTER(BatchResult)]
        L.spsp_batch_elements.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        _dev = L
    return _dev


def host_lib():
    global _host
    if _host is None:
        device_lib()
        L = C.CDLL(_need(os.path.join(LIB_DIR, "libspsp_host.so")))
        L.spsph_last_error.restype = C.c_char_p
        L.spsph_free.argtypes = [C.c_void_p]
        L.spsph_threshold.restype = C.c_uint64
        L.spsph_threshold.argtypes = [C.c_int, C.c_int, C.c_double]
        L.spsph_sub_sampler_main.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
        L.spsph_comparator_main.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
        L.spsph_sort_csv_main.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
        L.spsph_pack_fasta.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64),
                                       C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        L.spsph_postpass.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.c_int,
                                     C.c_double, C.c_uint, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                     C.POINTER(C.c_uint64)]
        L.spsph_decode_sketch.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                          C.POINTER(C.c_uint64), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                          C.POINTER(C.c_void_p)]
        L.spsph_format_csv.argtypes = [C.POINTER(C.c_char_p), C.c_uint32, C.c_uint32, C.c_void_p, C.c_int, C.c_void_p,
                                       C.c_int, C.c_uint, C.c_double, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.spsph_sketch_buffers.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_uint, C.c_int, C.c_uint32,
                                           C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_int,
                                           C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_double),
                                           C.POINTER(C.c_uint64)]
        L.spsph_sketcher_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_uint, C.c_int, C.c_int,
                                            C.POINTER(C.c_void_p)]
        L.spsph_sketcher_destroy.argtypes = [C.c_void_p]
        L.spsph_sketcher_ctx.restype = C.c_void_p
        L.spsph_sketcher_ctx.argtypes = [C.c_void_p]
        L.spsph_sketcher_run.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t),
                                         C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_double),
                                         C.POINTER(C.c_uint64)]
        L.spsph_postpass_batch.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_double, C.c_uint, C.c_int,
                                           C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.spsph_comparer_create.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.spsph_comparer_destroy.argtypes = [C.c_void_p]
        L.spsph_comparer_run.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_char_p),
                                         C.POINTER(C.c_size_t), C.c_void_p, C.c_void_p, C.POINTER(C.c_int),
                                         C.POINTER(C.c_float), C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
        L.spsph_compare_buffers.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.POINTER(C.c_char_p),
                                            C.POINTER(C.c_size_t), C.c_void_p, C.c_void_p, C.POINTER(C.c_int),
                                            C.POINTER(C.c_float), C.POINTER(C.c_uint64)]
        L.spsph_pipeline_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_uint, C.c_int, C.POINTER(C.c_void_p)]
        L.spsph_pipeline_destroy.argtypes = [C.c_void_p]
        L.spsph_pipeline_ctx.restype = C.c_void_p
        L.spsph_pipeline_ctx.argtypes = [C.c_void_p]
        L.spsph_pipeline_set_max_batch_bases.argtypes = [C.c_void_p, C.c_uint64]
        L.spsph_pipeline_set_ingest.argtypes = [C.c_void_p, C.c_int]
        L.spsph_pipeline_sketch.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                            C.POINTER(C.c_char_p), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                            C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
        L.spsph_pipeline_pack.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                          C.POINTER(C.c_char_p)]
        L.spsph_pipeline_finish.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                            C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
        L.spsph_pipeline_elem_off.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        L.spsph_pipeline_compare.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(C.c_int),
                                             C.POINTER(C.c_float), C.POINTER(C.c_uint64)]
        _host = L
    return _host


def _hcheck(rc: int, what: str) -> None:
    if rc != 0:
        raise SpspError(f"{what}: {host_lib().spsph_last_error().decode()}")


def _dcheck(rc: int, what: str) -> None:
    if rc != 0:
        raise SpspError(f"{what}: {device_lib().spsp_last_error().decode()} (rc={rc})")


def _f32(s: float) -> float:
    """The reference parses -s with stof (SubSampler.cpp:699)."""
    return float(np.float32(s))


def packed_words(n_bases: int) -> int:
    return int(device_lib().spsp_packed_words(n_bases))


def threshold(k: int, m: int, s: float) -> int:
    return int(host_lib().spsph_threshold(k, m, _f32(s)))


def _take(ptr: C.c_void_p, nbytes: int, dtype) -> np.ndarray:
    out = np.frombuffer(C.string_at(ptr, nbytes), dtype=dtype).copy() if nbytes else np.zeros(0, dtype)
    host_lib().spsph_free(ptr)
    return out


def pack_fasta(fasta: bytes, min_len: int) -> Tuple[np.ndarray, int, np.ndarray]:
    """[cpu] FASTA text -> (packed words, n_bases, record offsets)."""
    L = host_lib()
    w, ro = C.c_void_p(), C.c_void_p()
    nb, nr = C.c_uint64(), C.c_uint64()
    _hcheck(L.spsph_pack_fasta(fasta, len(fasta), min_len, C.byref(w), C.byref(nb), C.byref(ro), C.byref(nr)), "pack_fasta")
    words = _take(w, packed_words(nb.value) * 4, np.uint32)
    offs = _take(ro, (nr.value + 1) * 8, np.uint64)
    return words, int(nb.value), offs


def postpass(words: np.ndarray, rec_off: np.ndarray, hits: np.ndarray, k: int, m: int, s: float,
             abundance: int = 1) -> Tuple[bytes, int]:
    """[cpu] hit list -> sketch bytes (before gzip), selected k-mer occurrences."""
    L = host_lib()
    words = np.ascontiguousarray(words, np.uint32)
    rec_off = np.ascontiguousarray(rec_off, np.uint64)
    hits = np.ascontiguousarray(hits, HIT_DTYPE)
    out, n, sel = C.c_void_p(), C.c_size_t(), C.c_uint64()
    _hcheck(L.spsph_postpass(words.ctypes.data, rec_off.ctypes.data, rec_off.size - 1, hits.ctypes.data, hits.size,
                             k, m, _f32(s), abundance, C.byref(out), C.byref(n), C.byref(sel)), "postpass")
    return _take(out, n.value, np.uint8).tobytes(), int(sel.value)


def decode_sketch(sketch: bytes):
    """[cpu] sketch bytes -> (k, m, minimizer u32[], kmer_lo u64[], kmer_hi u64[] or None)."""
    L = host_lib()
    k, m, n = C.c_int(), C.c_int(), C.c_uint64()
    a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
    _hcheck(L.spsph_decode_sketch(sketch, len(sketch), C.byref(k), C.byref(m), C.byref(n), C.byref(a), C.byref(b),
                                  C.byref(c)), "decode_sketch")
    mn = _take(a, n.value * 4, np.uint32)
    lo = _take(b, n.value * 8, np.uint64)
    hi = _take(c, n.value * 8, np.uint64) if c.value else None
    return k.value, m.value, mn, lo, hi


def format_csv(names: Sequence[str], query_size: int, inter: np.ndarray, full_rows: bool, sizes: np.ndarray,
               jaccard: bool, precision: int = 6, min_threshold: float = 0.0) -> bytes:
    L = host_lib()
    n = len(names)
    arr = (C.c_char_p * n)(*[x.encode() for x in names])
    inter = np.ascontiguousarray(inter, np.uint32)
    sizes = np.ascontiguousarray(sizes, np.uint64)
    out, ln = C.c_void_p(), C.c_size_t()
    _hcheck(L.spsph_format_csv(arr, n, query_size, inter.ctypes.data, int(full_rows), sizes.ctypes.data, int(jaccard),
                               precision, min_threshold, C.byref(out), C.byref(ln)), "format_csv")
    return _take(out, ln.value, np.uint8).tobytes()


def write_csv_gz(path: str, names: Sequence[str], query_size: int, inter: np.ndarray, full_rows: bool, sizes: np.ndarray,
                 jaccard: bool, precision: int = 6, min_threshold: float = 0.0, threads: int = 4) -> int:
    """format_csv streamed into a gzip file (parallel row blocks, consecutive gzip members) -> CSV bytes before compression."""
    L = host_lib()
    n = len(names)
    arr = (C.c_char_p * n)(*[x.encode() for x in names])
    inter = np.ascontiguousarray(inter, np.uint32)
    sizes = np.ascontiguousarray(sizes, np.uint64)
    tb = C.c_uint64()
    L.spsph_write_csv_gz.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_uint32, C.c_uint32, C.c_void_p, C.c_int, C.c_void_p,
                                     C.c_int, C.c_uint, C.c_double, C.c_int, C.POINTER(C.c_uint64)]
    _hcheck(L.spsph_write_csv_gz(os.fsencode(path), arr, n, query_size, inter.ctypes.data, int(full_rows), sizes.ctypes.data,
                                 int(jaccard), precision, min_threshold, threads, C.byref(tb)), "write_csv_gz")
    return int(tb.value)


def sketch_buffers(fastas: Sequence[bytes], k: int = 31, m: int = 11, s: float = 1000.0, abundance: int = 1,
                   device: int = 0, threads: int = 8, scan_mode: int = SCAN_AUTO, info: Optional[dict] = None) -> List[bytes]:
    """GPU: FASTA texts -> sketch bytes (before gzip), through Subsampler + the C ABI."""
    L = host_lib()
    n = len(fastas)
    arr = (C.c_char_p * n)(*fastas)
    lens = (C.c_size_t * n)(*[len(f) for f in fastas])
    outs = (C.c_void_p * n)()
    olens = (C.c_size_t * n)()
    tim = (C.c_double * 3)()
    nl = C.c_uint64()
    _hcheck(L.spsph_sketch_buffers(device, k, m, _f32(s), abundance, scan_mode, n, arr, lens, threads, outs, olens,
                                   tim, C.byref(nl)), "sketch_buffers")
    res = []
    for i in range(n):
        res.append(C.string_at(outs[i], olens[i]) if olens[i] else b"")
        L.spsph_free(outs[i])
    if info is not None:
        info.update(pack_s=tim[0], scan_s=tim[1], post_s=tim[2], launches=int(nl.value))
    return res


class Sketcher:
    """Persistent Subsampler workers on one device context (one stream per worker)."""

    def __init__(self, k: int = 31, m: int = 11, s: float = 1000.0, abundance: int = 1, device: int = 0,
                 threads: int = 8, scan_mode: int = SCAN_AUTO):
        self.L = host_lib()
        self.h = C.c_void_p()
        self.k, self.m, self.s = k, m, s
        _hcheck(self.L.spsph_sketcher_create(device, k, m, _f32(s), abundance, scan_mode, threads, C.byref(self.h)),
                "sketcher_create")

    def close(self):
        if self.h:
            self.L.spsph_sketcher_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, fastas: Sequence[bytes], info: Optional[dict] = None) -> List[bytes]:
        n = len(fastas)
        arr = (C.c_char_p * n)(*fastas)
        lens = (C.c_size_t * n)(*[len(f) for f in fastas])
        outs = (C.c_void_p * n)()
        olens = (C.c_size_t * n)()
        tim = (C.c_double * 3)()
        nl = C.c_uint64()
        _hcheck(self.L.spsph_sketcher_run(self.h, n, arr, lens, outs, olens, tim, C.byref(nl)), "sketcher_run")
        res = []
        for i in range(n):
            res.append(C.string_at(outs[i], olens[i]) if olens[i] else b"")
            self.L.spsph_free(outs[i])
        if info is not None:
            info.update(pack_s=tim[0], scan_s=tim[1], post_s=tim[2], launches=int(nl.value))
        return res

    def device_context(self) -> "DeviceContext":
        """Borrowed view of the sketcher's spsp_ctx (not destroyed by the view)."""
        return DeviceContext._borrow(self.L.spsph_sketcher_ctx(self.h), self.k, self.m)


class Pipeline:
    """Batch pipeline (csrc/host/pipeline.h): pack all inputs on host threads into one
    pinned staging buffer, one scan + device post-pass, compare from the device-resident
    elements.  Inputs are FASTA texts (bytes) or file paths (str)."""

    STAT_NAMES = ("prep_s", "pack_s", "device_s", "assemble_s", "scan_ms", "post_ms", "hits", "elems", "h2d_bytes",
                  "d2h_bytes", "batches", "bases", "text_inputs", "ingest_ms")
    INGEST = {"host": 0, "device": 1, "auto": 2}

    def __init__(self, k: int = 31, m: int = 11, s: float = 1000.0, abundance: int = 1, device: int = 0,
                 threads: int = 8, max_batch_bases: Optional[int] = None, ingest: Optional[str] = None):
        """ingest: who cleans + packs the FASTA text -- "host" (host threads, the default unless SPSP_INGEST says
        otherwise), "device" (raw text over PCIe + ingest kernels) or "auto" (both, on one work queue)."""
        self.L = host_lib()
        self.h = C.c_void_p()
        self.k, self.m, self.s = k, m, s
        _hcheck(self.L.spsph_pipeline_create(device, k, m, _f32(s), abundance, threads, C.byref(self.h)), "pipeline_create")
        if max_batch_bases is not None:
            _hcheck(self.L.spsph_pipeline_set_max_batch_bases(self.h, max_batch_bases), "pipeline_set_max_batch_bases")
        if ingest is not None:
            _hcheck(self.L.spsph_pipeline_set_ingest(self.h, self.INGEST[ingest]), "pipeline_set_ingest")
        self.n_last = 0

    def close(self):
        if self.h:
            self.L.spsph_pipeline_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _marshal(inputs):
        """inputs: FASTA text as bytes / PinnedBuffer, or a path (str / os.PathLike)."""
        n = len(inputs)
        mem = lambda x: isinstance(x, (bytes, PinnedBuffer))
        addr = lambda x: x.ptr if isinstance(x, PinnedBuffer) else C.cast(C.c_char_p(x), C.c_void_p).value
        data = (C.c_void_p * n)(*[addr(x) if mem(x) else None for x in inputs])
        lens = (C.c_size_t * n)(*[len(x) if mem(x) else 0 for x in inputs])
        paths = (C.c_char_p * n)(*[os.fsencode(x) if not mem(x) else None for x in inputs])
        return n, data, lens, paths

    def _collect(self, n, outs, olens, ok, st, nl, info):
        res = []
        for i in range(n):
            res.append((C.string_at(outs[i], olens[i]) if olens[i] else b"") if ok[i] else None)
            self.L.spsph_free(outs[i])
        self.n_last = n
        if info is not None:
            info.update({k_: float(st[i]) for i, k_ in enumerate(self.STAT_NAMES)})
            info["launches"] = int(nl.value)
        return res

    def sketch(self, inputs: Sequence, info: Optional[dict] = None) -> List[Optional[bytes]]:
        """-> sketch bytes (before gzip) per input; None for a file that cannot be opened."""
        n, data, lens, paths = self._marshal(inputs)
        outs, olens, ok = (C.c_void_p * n)(), (C.c_size_t * n)(), (C.c_int * n)()
        st, nl = (C.c_double * 14)(), C.c_uint64()
        _hcheck(self.L.spsph_pipeline_sketch(self.h, n, data, lens, paths, outs, olens, ok, st, C.byref(nl)),
                "pipeline_sketch")
        return self._collect(n, outs, olens, ok, st, nl, info)

    def pack(self, inputs: Sequence) -> None:
        """First half of sketch(): prepare, pack, queue the copies.  Returns when the host work is done."""
        n, data, lens, paths = self._marshal(inputs)
        _hcheck(self.L.spsph_pipeline_pack(self.h, n, data, lens, paths), "pipeline_pack")
        self._n_packed = n

    def finish(self, info: Optional[dict] = None) -> List[Optional[bytes]]:
        """Second half of sketch(): device phase of the job pack() started."""
        n = self._n_packed
        outs, olens, ok = (C.c_void_p * n)(), (C.c_size_t * n)(), (C.c_int * n)()
        st, nl = (C.c_double * 14)(), C.c_uint64()
        _hcheck(self.L.spsph_pipeline_finish(self.h, n, outs, olens, ok, st, C.byref(nl)), "pipeline_finish")
        return self._collect(n, outs, olens, ok, st, nl, info)

    def compare(self, query_size: Optional[int] = None, info: Optional[dict] = None):
        """Compare the sketches of the last sketch() call -> (inter[rows, n], sizes[n], full_rows)."""
        n = self.n_last
        q = n if query_size is None else query_size
        rows = n if q >= n else q
        inter = np.zeros((rows, n), np.uint32)
        sizes = np.zeros(n, np.uint64)
        full, ms, nl = C.c_int(), C.c_float(), C.c_uint64()
        _hcheck(self.L.spsph_pipeline_compare(self.h, q, inter.ctypes.data, sizes.ctypes.data, C.byref(full), C.byref(ms),
                                              C.byref(nl)), "pipeline_compare")
        if info is not None:
            info.update(kernel_ms=float(ms.value), launches=int(nl.value))
        return inter, sizes, bool(full.value)

    def elem_off(self) -> Tuple[np.ndarray, bool]:
        """(element offsets of the last sketch() call, still-on-device flag)."""
        off = np.zeros(self.n_last + 1, np.uint64)
        od = C.c_int()
        _hcheck(self.L.spsph_pipeline_elem_off(self.h, off.ctypes.data, C.byref(od)), "pipeline_elem_off")
        return off, bool(od.value)

    def device_context(self) -> "DeviceContext":
        return DeviceContext._borrow(self.L.spsph_pipeline_ctx(self.h), self.k, self.m)


class BatchSketches(collections.abc.Sequence):
    """Sketches of a device batch: the bytes of all inputs back to back + offsets; item i (header line + body, what the
    reference writes into the .gz file, SubSampler.cpp:459-504) is assembled on access.  Compares equal to any
    sequence of the same bytes."""

    def __init__(self, body: bytes, body_off: np.ndarray, selected: np.ndarray, head: str, tail: str):
        self.body, self.body_off, self.selected, self.head, self.tail = body, body_off, selected, head, tail

    def __len__(self):
        return int(self.selected.size)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        return (self.head + str(int(self.selected[i])) + self.tail).encode() + self.body[int(self.body_off[i]):int(self.body_off[i + 1])]

    def __eq__(self, other):
        try:
            return len(self) == len(other) and all(a == b for a, b in zip(self, other))
        except TypeError:
            return NotImplemented

    def __add__(self, other):
        return list(self) + list(other)

    def __radd__(self, other):
        return list(other) + list(self)

    @property
    def nbytes(self) -> int:
        """Bytes that came back from the device for this batch (bodies + offsets + counts)."""
        return len(self.body) + self.body_off.nbytes + self.selected.nbytes


class PinnedBuffer:
    """Bytes in page-locked host memory (cudaHostAlloc through the C ABI): FASTA text that the device-side ingest
    can copy asynchronously, straight from where it lies.  Keep it alive until the job that uses it has finished."""

    def __init__(self, data: bytes):
        self.L = device_lib()
        self.n = len(data)
        p = C.c_void_p()
        _dcheck(self.L.spsp_host_alloc(C.byref(p), max(1, self.n)), "spsp_host_alloc")
        self.ptr = p.value
        C.memmove(self.ptr, data, self.n)

    def __len__(self):
        return self.n

    def tobytes(self) -> bytes:
        return C.string_at(self.ptr, self.n)

    def close(self):
        if self.ptr:
            self.L.spsp_host_free(C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BatchStream:
    """A stream of batches through two `Pipeline`s (two device contexts on one GPU): the device phase (scan +
    post-pass) and the compare stage of batch i run on a background thread while the host threads clean, pack
    and copy batch i+1, so a long job is bound by its slower half instead of their sum.  `submit(inputs)` returns the finished result of the
    PREVIOUS batch (None for the first), `drain()` the last one.  A result is
    (sketches, (inter, sizes, full_rows), sketch_info, compare_info).

    compare_fn(pipeline, info) -> (inter, sizes, full_rows) replaces the single-GPU `pipeline.compare` (the
    multi-GPU exchange is plugged in here)."""

    def __init__(self, k: int = 31, m: int = 11, s: float = 1000.0, abundance: int = 1, device: int = 0,
                 threads: int = 8, compare_fn=None, pipelines: Optional[Sequence["Pipeline"]] = None):
        from concurrent.futures import ThreadPoolExecutor
        self.pipes = list(pipelines) if pipelines else [Pipeline(k, m, s, abundance, device, threads) for _ in range(2)]
        self._own = not pipelines
        self.compare_fn = compare_fn or (lambda pl, info: pl.compare(info=info))
        self.pool = ThreadPoolExecutor(1)
        self.pending = None
        self.n = 0

    def submit(self, inputs: Sequence):
        pl = self.pipes[self.n % 2]
        self.n += 1
        pl.pack(inputs)                                  # host half; the previous batch's device half + compare run meanwhile
        prev = self._finish()
        self.pending = self.pool.submit(self._back_half, pl)
        return prev

    def _back_half(self, pl):
        info, cinfo = {}, {}
        sks = pl.finish(info=info)
        res = self.compare_fn(pl, cinfo)
        return sks, res, info, cinfo

    def _finish(self):
        if self.pending is None:
            return None
        fut, self.pending = self.pending, None
        return fut.result()

    def drain(self):
        return self._finish()

    def close(self):
        self._finish()
        self.pool.shutdown()
        if self._own:
            for pl in self.pipes:
                pl.close()


def postpass_batch(packed: np.ndarray, base_off: np.ndarray, n_bases: np.ndarray, rec_off: np.ndarray,
                   rec_first: np.ndarray, hits: np.ndarray, k: int, m: int, s: float, abundance: int = 1,
                   threads: int = 8) -> List[bytes]:
    """[cpu] hits of one scan over several inputs packed back to back -> one sketch per input."""
    L = host_lib()
    n = int(np.asarray(base_off).size)
    packed = np.ascontiguousarray(packed, np.uint32)
    base_off = np.ascontiguousarray(base_off, np.uint64)
    n_bases = np.ascontiguousarray(n_bases, np.uint64)
    rec_off = np.ascontiguousarray(rec_off, np.uint64)
    rec_first = np.ascontiguousarray(rec_first, np.uint64)
    hits = np.ascontiguousarray(hits, HIT_DTYPE)
    outs = (C.c_void_p * n)()
    olens = (C.c_size_t * n)()
    _hcheck(L.spsph_postpass_batch(packed.ctypes.data, n, base_off.ctypes.data, n_bases.ctypes.data,
                                   rec_off.ctypes.data, rec_first.ctypes.data, hits.ctypes.data, hits.size, k, m,
                                   _f32(s), abundance, threads, outs, olens), "postpass_batch")
    res = []
    for i in range(n):
        res.append(C.string_at(outs[i], olens[i]) if olens[i] else b"")
        L.spsph_free(outs[i])
    return res


def compare_buffers(sketches: Sequence[bytes], query_size: Optional[int] = None, n_gpus: int = 1,
                    info: Optional[dict] = None):
    """GPU: sketch bytes -> (inter[rows, n] uint32, sizes[n] uint64, full_rows)."""
    L = host_lib()
    n = len(sketches)
    q = n if query_size is None else query_size
    arr = (C.c_char_p * n)(*sketches)
    lens = (C.c_size_t * n)(*[len(x) for x in sketches])
    rows = n if q >= n else q
    inter = np.zeros((rows, n), np.uint32)
    sizes = np.zeros(n, np.uint64)
    full, ms, nl = C.c_int(), C.c_float(), C.c_uint64()
    _hcheck(L.spsph_compare_buffers(n_gpus, n, q, arr, lens, inter.ctypes.data, sizes.ctypes.data, C.byref(full),
                                    C.byref(ms), C.byref(nl)), "compare_buffers")
    if info is not None:
        info.update(kernel_ms=float(ms.value), launches=int(nl.value))
    return inter, sizes, bool(full.value)


class Comparer:
    """Persistent Comparator (keeps its device contexts between calls)."""

    def __init__(self, n_gpus: int = 1, threads: int = 0):
        self.L = host_lib()
        self.h = C.c_void_p()
        _hcheck(self.L.spsph_comparer_create(n_gpus, threads, C.byref(self.h)), "comparer_create")

    def close(self):
        if self.h:
            self.L.spsph_comparer_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, sketches: Sequence[bytes], query_size: Optional[int] = None, info: Optional[dict] = None):
        n = len(sketches)
        q = n if query_size is None else query_size
        arr = (C.c_char_p * n)(*sketches)
        lens = (C.c_size_t * n)(*[len(x) for x in sketches])
        rows = n if q >= n else q
        inter = np.zeros((rows, n), np.uint32)
        sizes = np.zeros(n, np.uint64)
        full, ms, nl = C.c_int(), C.c_float(), C.c_uint64()
        tim = (C.c_double * 2)()
        _hcheck(self.L.spsph_comparer_run(self.h, n, q, arr, lens, inter.ctypes.data, sizes.ctypes.data,
                                          C.byref(full), C.byref(ms), C.byref(nl), tim), "comparer_run")
        if info is not None:
            info.update(kernel_ms=float(ms.value), launches=int(nl.value), decode_s=tim[0], device_s=tim[1])
        return inter, sizes, bool(full.value)


def _run_main(fn, argv: Sequence[str]) -> int:
    args = [a.encode() for a in argv]
    arr = (C.c_char_p * (len(args) + 1))(*args, None)
    return int(fn(len(args), arr))


def run_sub_sampler(args: Sequence[str]) -> int:
    """In-process `sub_sampler <args>` (writes into the CWD like the reference)."""
    return _run_main(host_lib().spsph_sub_sampler_main, ["sub_sampler", *args])


def run_sort_csv(args: Sequence[str]) -> int:
    """[cpu] In-process `sortCSV matrix.csv[.gz] out.csv names.txt`."""
    return _run_main(host_lib().spsph_sort_csv_main, ["sortCSV", *args])


def run_comparator(args: Sequence[str]) -> int:
    return _run_main(host_lib().spsph_comparator_main, ["comparator", *args])


class DeviceContext:
    """Thin RAII wrapper of spsp_ctx for tests / bench (device C ABI, raw pointers)."""

    def __init__(self, k: int, m: int, thr: int, device: int = 0, n_slots: int = 1):
        self.L = device_lib()
        self.h = C.c_void_p()
        _dcheck(self.L.spsp_create(device, k, m, thr, n_slots, C.byref(self.h)), "spsp_create")
        self.k, self.m, self.thr = k, m, thr

    @classmethod
    def _borrow(cls, handle: int, k: int, m: int) -> "DeviceContext":
        o = cls.__new__(cls)
        o.L = device_lib()
        o.h = C.c_void_p(handle)
        o.k, o.m, o.thr = k, m, None
        o._borrowed = True
        return o

    def close(self):
        if self.h and not getattr(self, "_borrowed", False):
            self.L.spsp_destroy(self.h)
        self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def config(self, mode: int):
        _dcheck(self.L.spsp_scan_config(self.h, mode), "spsp_scan_config")

    def filter_info(self) -> dict:
        """Filter table of this context: kind (0 bit, 1 byte, 2 bank-private bit + hash set, -1 none), g, n_selected."""
        kind, g, n = C.c_int(), C.c_int(), C.c_uint64()
        _dcheck(self.L.spsp_scan_filter_info(self.h, C.byref(kind), C.byref(g), C.byref(n)), "spsp_scan_filter_info")
        return {"kind": int(kind.value), "g": int(g.value), "n_selected": int(n.value)}

    def scan(self, words: np.ndarray, n_bases: int, slot: int = 0) -> np.ndarray:
        """Host buffer in, hits out (submit + collect)."""
        words = np.ascontiguousarray(words, np.uint32)
        assert words.size >= packed_words(n_bases)
        _dcheck(self.L.spsp_scan_submit(self.h, slot, words.ctypes.data, n_bases), "spsp_scan_submit")
        cap = 1 << 16
        n = C.c_uint64()
        hits = np.zeros(cap, HIT_DTYPE)
        rc = self.L.spsp_scan_collect(self.h, slot, hits.ctypes.data, cap, C.byref(n))
        if rc == -2:
            hits = np.zeros(n.value, HIT_DTYPE)
            rc = self.L.spsp_scan_collect(self.h, slot, hits.ctypes.data, n.value, C.byref(n))
        _dcheck(rc, "spsp_scan_collect")
        return hits[: n.value].copy()

    def scan_device(self, d_packed: int, n_bases: int, d_hits: int, cap: int, d_count: int, slot: int = 0):
        _dcheck(self.L.spsp_scan_device(self.h, slot, d_packed, n_bases, d_hits, cap, d_count), "spsp_scan_device")

    def sync(self, slot: int = 0):
        _dcheck(self.L.spsp_sync(self.h, slot), "spsp_sync")

    def scan_kernel_ms(self, slot: int = 0) -> float:
        ms = C.c_float()
        _dcheck(self.L.spsp_scan_last_kernel_ms(self.h, slot, C.byref(ms)), "spsp_scan_last_kernel_ms")
        return float(ms.value)

    def stream(self, slot: int = 0) -> int:
        p = C.c_void_p()
        _dcheck(self.L.spsp_stream(self.h, slot, C.byref(p)), "spsp_stream")
        return int(p.value or 0)

    def sketch_batch(self, words, n_bases: int, rec_begin, rec_end, rec_input, n_inputs: int, s: float,
                     abundance: int = 1, slot: int = 0, device_ptr: Optional[int] = None, info: Optional[dict] = None,
                     rec_device: Optional[Tuple[int, int, int, int]] = None):
        """Scan + device post-pass of a whole batch.  `words` is a host array (copied) unless
        device_ptr gives a device-resident packed buffer.  rec_device = (d_begin, d_end, d_input, n_rec)
        passes record tables that already live on the device.  Returns the sketch bytes of every input."""
        if rec_device is not None:
            class _P:                                     # raw device pointers behind the ndarray attributes used below
                def __init__(self, ptr, n):
                    self.ctypes = type("c", (), {"data": ptr})
                    self.size = n
            rec_begin, rec_end, rec_input = (_P(rec_device[i], rec_device[3]) for i in range(3))
        else:
            rec_begin = np.ascontiguousarray(rec_begin, np.uint64)
            rec_end = np.ascontiguousarray(rec_end, np.uint64)
            rec_input = np.ascontiguousarray(rec_input, np.uint32)
        res = BatchResult()
        if device_ptr is None:
            words = np.ascontiguousarray(words, np.uint32)
            assert words.size >= packed_words(n_bases)
            rc = self.L.spsp_sketch_batch(self.h, slot, words.ctypes.data, n_bases, rec_begin.ctypes.data,
                                          rec_end.ctypes.data, rec_input.ctypes.data, rec_begin.size, n_inputs,
                                          abundance, C.byref(res))
        else:
            rc = self.L.spsp_sketch_batch_device(self.h, slot, device_ptr, n_bases, rec_begin.ctypes.data,
                                                 rec_end.ctypes.data, rec_input.ctypes.data, rec_begin.size, n_inputs,
                                                 abundance, C.byref(res))
        _dcheck(rc, "spsp_sketch_batch")
        return self._batch_out(res, n_inputs, s, info)

    def _batch_out(self, res, n_inputs: int, s: float, info: Optional[dict]):
        # three bulk copies out of the library's result buffers; the per-input sketches (header line + body slice) are
        # put together when somebody asks for them: a caller that runs thousands of batches per second cannot spend
        # 0.2 ms of interpreter time per batch on 64 string concatenations
        u64 = lambda ptr, cnt: np.frombuffer(C.string_at(ptr, cnt * 8), np.uint64)
        body_off = u64(res.body_off, n_inputs + 1)
        total = int(body_off[n_inputs])
        out = BatchSketches(C.string_at(res.body, total) if total else b"", body_off, u64(res.selected, max(n_inputs, 1))[:n_inputs],
                            f"{2 * self.k - self.m} {self.m} ", " %f\n" % _f32(s))
        if info is not None:
            info.update(n_hits=int(res.n_hits), n_elems=int(res.n_elems), scan_ms=float(res.scan_ms),
                        post_ms=float(res.post_ms), elem_off=u64(res.elem_off, n_inputs + 1))
        return out

    def ingest_texts(self, texts: Sequence[bytes], slot: int = 0, info: Optional[dict] = None):
        """Device-side ingest of FASTA texts, the C ABI calls one by one (tests / diagnostics): every text gets a
        region of a staged batch sized for its byte length, the raw bytes are uploaded, spsp_batch_text_pack cleans
        and packs them.  -> (n_bases[i], words[i] (the region's packed words), base offset of every region,
        (rec_begin, rec_end, rec_input) in batch coordinates, total bases of the batch)."""
        n = len(texts)
        lens = np.array([len(t) for t in texts], np.uint64)
        w_off = np.zeros(n, np.uint64); t_off = np.zeros(n, np.uint64)
        tw = tt = 0
        for i in range(n):
            w_off[i], t_off[i] = tw, tt
            tw += packed_words(int(lens[i])); tt += (int(lens[i]) + 15) & ~15
        n_total = 16 * tw
        _dcheck(self.L.spsp_batch_reserve(self.h, slot, packed_words(n_total)), "spsp_batch_reserve")
        _dcheck(self.L.spsp_batch_text_reserve(self.h, slot, tt), "spsp_batch_text_reserve")
        for i, t in enumerate(texts):
            if len(t):
                _dcheck(self.L.spsp_batch_text_upload(self.h, slot, -1, int(t_off[i]), C.cast(C.c_char_p(t), C.c_void_p), len(t)),
                        "spsp_batch_text_upload")
        idx = np.arange(n, dtype=np.uint32)
        nb = np.zeros(max(n, 1), np.uint64); nr = np.zeros(max(n, 1), np.uint64)
        _dcheck(self.L.spsp_batch_text_pack(self.h, slot, n, t_off.ctypes.data, lens.ctypes.data, w_off.ctypes.data,
                                            idx.ctypes.data, nb.ctypes.data, nr.ctypes.data), "spsp_batch_text_pack")
        n_rec = int(nr[:n].sum())
        rb = np.zeros(max(n_rec, 1), np.uint64); re_ = np.zeros(max(n_rec, 1), np.uint64); ri = np.zeros(max(n_rec, 1), np.uint32)
        got = C.c_uint64()
        _dcheck(self.L.spsp_batch_text_records(self.h, slot, rb.ctypes.data, re_.ctypes.data, ri.ctypes.data, max(n_rec, 1),
                                               C.byref(got)), "spsp_batch_text_records")
        assert got.value == n_rec
        words = []
        for i in range(n):
            w = np.zeros(packed_words(int(nb[i])), np.uint32)
            _dcheck(self.L.spsp_batch_download(self.h, slot, int(w_off[i]), w.ctypes.data, w.size), "spsp_batch_download")
            words.append(w)
        if info is not None:
            ms = C.c_float()
            _dcheck(self.L.spsp_batch_text_last_ms(self.h, slot, C.byref(ms)), "spsp_batch_text_last_ms")
            info["ingest_ms"] = float(ms.value)
        return nb[:n].copy(), words, 16 * w_off, (rb[:n_rec], re_[:n_rec], ri[:n_rec]), n_total

    def sketch_staged(self, n_bases: int, rec_begin, rec_end, rec_input, n_inputs: int, s: float, abundance: int = 1,
                      slot: int = 0, info: Optional[dict] = None):
        """spsp_sketch_batch_staged on what was uploaded / ingested on the slot; the record arrays describe the
        host-packed inputs only (may be empty).  Returns the sketch bytes of every input."""
        rec_begin = np.ascontiguousarray(rec_begin, np.uint64)
        rec_end = np.ascontiguousarray(rec_end, np.uint64)
        rec_input = np.ascontiguousarray(rec_input, np.uint32)
        res = BatchResult()
        _dcheck(self.L.spsp_sketch_batch_staged(self.h, slot, n_bases, rec_begin.ctypes.data, rec_end.ctypes.data,
                                                rec_input.ctypes.data, rec_begin.size, n_inputs, abundance, C.byref(res)),
                "spsp_sketch_batch_staged")
        return self._batch_out(res, n_inputs, s, info)

    def dense_stats(self, words, n_bases: int, rec_begin, rec_end, rec_input, n_inputs: int, slot: int = 0,
                    device_ptr: Optional[int] = None, info: Optional[dict] = None):
        """Dense totals of a batch -> (total_superkmers[n_inputs], selected_kmers[n_inputs])."""
        rec_begin = np.ascontiguousarray(rec_begin, np.uint64)
        rec_end = np.ascontiguousarray(rec_end, np.uint64)
        rec_input = np.ascontiguousarray(rec_input, np.uint32)
        tot = np.zeros(max(n_inputs, 1), np.uint64); sel = np.zeros(max(n_inputs, 1), np.uint64)
        ms = C.c_float()
        if device_ptr is None:
            words = np.ascontiguousarray(words, np.uint32)
            assert words.size >= packed_words(n_bases)
            rc = self.L.spsp_dense_stats(self.h, slot, words.ctypes.data, n_bases, rec_begin.ctypes.data, rec_end.ctypes.data,
                                         rec_input.ctypes.data, rec_begin.size, n_inputs, tot.ctypes.data, sel.ctypes.data,
                                         C.byref(ms))
        else:
            rc = self.L.spsp_dense_stats_device(self.h, slot, device_ptr, n_bases, rec_begin.ctypes.data,
                                                rec_end.ctypes.data, rec_input.ctypes.data, rec_begin.size, n_inputs,
                                                tot.ctypes.data, sel.ctypes.data, C.byref(ms))
        _dcheck(rc, "spsp_dense_stats")
        if info is not None:
            info.update(dense_ms=float(ms.value))
        return tot[:n_inputs], sel[:n_inputs]

    def batch_elements(self, n_elems: int, slot: int = 0, want_hi: bool = False):
        """Host copy of the last batch's elements: (minimizer u32[], kmer_lo u64[], kmer_hi u64[] | None)."""
        mn = np.zeros(n_elems, np.uint32); lo = np.zeros(n_elems, np.uint64)
        hi = np.zeros(n_elems, np.uint64) if want_hi else None
        _dcheck(self.L.spsp_batch_elements(self.h, slot, mn.ctypes.data, lo.ctypes.data,
                                           hi.ctypes.data if want_hi else None, None, None, None), "spsp_batch_elements")
        return mn, lo, hi

    def batch_element_ptrs(self, slot: int = 0):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _dcheck(self.L.spsp_batch_elements(self.h, slot, None, None, None, C.byref(a), C.byref(b), C.byref(c)),
                "spsp_batch_elements")
        return a.value, b.value, c.value

    def nccl_init(self, rank: int, world: int, broadcast):
        """Join the ranks' contexts into one NCCL communicator.  `broadcast(buf: np.ndarray[128] uint8) ->
        np.ndarray` must return rank 0's buffer on every rank (any side channel, e.g. torch.distributed)."""
        buf = np.zeros(128, np.uint8)
        if rank == 0:
            _dcheck(self.L.spsp_nccl_unique_id(buf.ctypes.data), "spsp_nccl_unique_id")
        buf = np.ascontiguousarray(broadcast(buf), np.uint8)
        _dcheck(self.L.spsp_nccl_init(self.h, buf.ctypes.data, rank, world), "spsp_nccl_init")
        self._world = world

    def cmp_exchange_batch(self, cap_sketches: int, rank: int, slot: int = 0, info: Optional[dict] = None):
        """All-vs-all compare of the union of every rank's last batch (collective).
        -> (inter[N, N] uint32 (complete on rank 0, None elsewhere), sizes[N] uint64)."""
        inter = np.zeros((cap_sketches, cap_sketches), np.uint32) if rank == 0 else None
        sizes = np.zeros(cap_sketches, np.uint64)
        n, ms = C.c_uint32(), C.c_float()
        l0 = self.launches()
        _dcheck(self.L.spsp_cmp_exchange_batch(self.h, slot, inter.ctypes.data if inter is not None else None, cap_sketches,
                                               sizes.ctypes.data, cap_sketches, C.byref(n), C.byref(ms)),
                "spsp_cmp_exchange_batch")
        if info is not None:
            info.update(kernel_ms=float(ms.value), launches=self.launches() - l0)
        nt = int(n.value)
        return (inter[:nt, :nt] if inter is not None else None), sizes[:nt]

    def cmp_exchange(self, sizes_local, d_minim: int, d_klo: int, d_khi: Optional[int], n_queries_local: Optional[int],
                     cap_rows: int, cap_cols: int, rank: int, info: Optional[dict] = None):
        """Collective compare of the union of every rank's sketches (device-resident elements, host size list).
        n_queries_local = None: all-vs-all; else this rank's first n_queries_local sketches are queries and the
        result is queries x all, in the order [all queries rank-major | all references rank-major].
        -> (inter[rows, cols] uint32 on rank 0 (None elsewhere), sizes[cols] uint64)."""
        sizes_local = np.ascontiguousarray(sizes_local, np.uint64)
        n_local = int(sizes_local.size)
        sym = n_queries_local is None
        q_local = n_local if sym else int(n_queries_local)
        inter = np.zeros((cap_rows, cap_cols), np.uint32) if rank == 0 else None
        sizes = np.zeros(cap_cols, np.uint64)
        nr, nc, ms = C.c_uint32(), C.c_uint32(), C.c_float()
        l0 = self.launches()
        _dcheck(self.L.spsp_cmp_exchange(self.h, n_local, q_local, sizes_local.ctypes.data, d_minim, d_klo, d_khi, int(sym),
                                         inter.ctypes.data if inter is not None else None, cap_cols, sizes.ctypes.data,
                                         cap_rows, cap_cols, C.byref(nr), C.byref(nc), C.byref(ms)), "spsp_cmp_exchange")
        if info is not None:
            info.update(kernel_ms=float(ms.value), launches=self.launches() - l0)
        r_, c_ = int(nr.value), int(nc.value)
        return (np.ascontiguousarray(inter[:r_, :c_]) if inter is not None else None), sizes[:c_]

    def cmp_load_batch(self, slot: int = 0):
        _dcheck(self.L.spsp_cmp_load_batch(self.h, slot), "spsp_cmp_load_batch")

    def cmp_load(self, sk_off: np.ndarray, minim: np.ndarray, klo: np.ndarray, khi: Optional[np.ndarray] = None):
        sk_off = np.ascontiguousarray(sk_off, np.uint64)
        minim = np.ascontiguousarray(minim, np.uint32)
        klo = np.ascontiguousarray(klo, np.uint64)
        khi_p = None
        if khi is not None:
            khi = np.ascontiguousarray(khi, np.uint64)
            khi_p = khi.ctypes.data
        _dcheck(self.L.spsp_cmp_load(self.h, sk_off.size - 1, sk_off.ctypes.data, minim.ctypes.data, klo.ctypes.data,
                                     khi_p), "spsp_cmp_load")

    def cmp_load_device(self, sk_off: np.ndarray, d_minim: int, d_klo: int, d_khi: Optional[int] = None):
        sk_off = np.ascontiguousarray(sk_off, np.uint64)
        _dcheck(self.L.spsp_cmp_load_device(self.h, sk_off.size - 1, sk_off.ctypes.data, d_minim, d_klo, d_khi),
                "spsp_cmp_load_device")

    def cmp_run(self, rows: Tuple[int, int], cols: Tuple[int, int], symmetric: bool, rank: int = 0, ranks: int = 1) -> np.ndarray:
        out = np.zeros((rows[1] - rows[0], cols[1] - cols[0]), np.uint32)
        _dcheck(self.L.spsp_cmp_run(self.h, rows[0], rows[1], cols[0], cols[1], int(symmetric), rank, ranks,
                                    out.ctypes.data, out.shape[1]), "spsp_cmp_run")
        return out

    def cmp_run_device(self, rows, cols, symmetric: bool, rank: int, ranks: int, d_out: int, ld: int):
        _dcheck(self.L.spsp_cmp_run_device(self.h, rows[0], rows[1], cols[0], cols[1], int(symmetric), rank, ranks,
                                           d_out, ld), "spsp_cmp_run_device")

    def cmp_kernel_ms(self) -> float:
        ms = C.c_float()
        _dcheck(self.L.spsp_cmp_last_kernel_ms(self.h, C.byref(ms)), "spsp_cmp_last_kernel_ms")
        return float(ms.value)

    def launches(self) -> int:
        n = C.c_uint64()
        _dcheck(self.L.spsp_launch_count(self.h, C.byref(n)), "spsp_launch_count")
        return int(n.value)
