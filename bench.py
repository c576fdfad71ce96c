#!/usr/bin/env python
"""bench.py -- sketch Gbp/s (+ all-vs-all compare pairs/s) of the SuperSampler
hot path on B200, next to the reference's CPU path.

One *step* = one pass of the hot path over one batch: sketch G synthetic
genomes (default: BASELINE config 2, 64 x 5 Mbp, k31 m11 s1000) and compare
the G sketches all-vs-all.

  value : inputs (2-bit packed genomes) resident in HBM; per step: one scan
          launch over the batch, the exact post-pass on the device (hits ->
          super-k-mers -> buckets -> sketch bytes), sketch bytes D2H, compare
          kernel on the device-resident elements, matrix D2H.  --depth
          batches (default 4) are in flight on their own contexts / streams;
          the timed region ends when the last one has been retired.
  roofline : a third region, rank 0: the scan kernel alone over the rotating
          device-resident replicas, CUDA events on its launching stream
          (the in-step per-launch figures of both other regions are kept
          beside it in the JSON line).
  e2e   : the same work through the public host API with HOST buffers
          (supersampler_b200.Pipeline): FASTA text in host memory -> host
          threads clean/pack into pinned memory -> async H2D per input -> scan ->
          device post-pass -> sketch bytes D2H -> compare -> matrix D2H; all
          copies inside the timed region.
  --impl reference : the unmodified reference binaries (oracle/_ref, built from
          /root/reference in the build container) on this box's host cores,
          same files / parameters (64 x N genomes for --gpus N); C restatement
          if the binaries are absent.
  parity : after the timed regions rank 0 runs oracle/_ref on the job's own
          FASTA files (all 64 x N genomes) and compares every sketch and both
          CSV matrices byte for byte (`parity_vs_reference`).
  extra : BASELINE configs 3, 4, 5 at their stated sizes, once, outside the
          headline regions, inputs synthesised on the device, each with its own
          parity block against oracle/_ref (N > 1: configs 3 and 5, each as one
          fixed job over the ranks = the strong-scaling figures; config 5
          through the query-mode exchange).

Launch:  python bench.py --gpus N --steps K --warmup W        (N=1)
         python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
Rank 0 prints ONE JSON line on stdout.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "sketch_gbp_per_s"
UNIT = "Gbp/s"
ALL_CORES = os.sched_getaffinity(0)          # before a rank confines itself to its share of the cores


def _all_cores():
    """preexec_fn of the reference's processes: they get every core of the box, whatever this rank is bound to."""
    os.sched_setaffinity(0, ALL_CORES)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_traffic(bases_per_launch):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (same workload)."""
    for nm in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", nm)) as f:
                t = json.load(f)
            if abs(bases_per_launch - 320012288) < 1e6:
                return t["traffic_bytes_per_launch"], t["source"], t.get("pipes_pct_of_peak")
        except Exception:
            continue
    return None, None, None


def measured_compare_pipes():
    """ALU / LSU / issue-active percentages of hashjoin_kernel from the committed ncu capture (profiles/)."""
    for nm in ("r02_compare_pipes.json", "r01_compare_pipes.json"):
        try:
            with open(os.path.join(ROOT, "profiles", nm)) as f:
                return json.load(f)
        except Exception:
            continue
    return {}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ data

def _fasta_member(a):
    from supersampler_b200 import synth
    n_bases, seed, idx = a
    fam = _fasta_member.fam.get((n_bases, seed))
    if fam is None:
        fam = _fasta_member.fam[(n_bases, seed)] = synth.Family(n_bases, seed)
    return fam.fasta(idx)


_fasta_member.fam = {}


def make_fastas(n_genomes: int, n_bases: int, first: int, seed: int = 42, procs: int = 1):
    """Members first .. first+n-1 of the C2 family as FASTA text (numpy streams: the same genome
    whatever the rank layout).  procs > 1: worker processes (the reference arm of an N-GPU run
    needs 64 x N genomes on one rank)."""
    names = [f"g{first + i:05d}" for i in range(n_genomes)]
    jobs = [(n_bases, seed, first + i) for i in range(n_genomes)]
    if procs > 1 and n_genomes > 64:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(procs, n_genomes)) as pool:
            return pool.map(_fasta_member, jobs, chunksize=4), names
    return [_fasta_member(j) for j in jobs], names


def write_files(fastas, names, d):
    paths = []
    for fa, nm in zip(fastas, names):
        p = os.path.join(d, nm + ".fa")
        with open(p, "wb") as f:
            f.write(fa)
        paths.append(p)
    return paths


def scratch_dir():
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    return tempfile.mkdtemp(prefix="spsp_bench_", dir=base)


def scratch_free_bytes():
    base = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    try:
        st = os.statvfs(base)
        return st.f_bavail * st.f_frsize
    except OSError:
        return 0


# ------------------------------------------------------- reference CPU path

def ref_sketch(paths, k, m, s, wd, cores):
    """oracle/_ref/sub_sampler -f over `paths` -> (seconds, sketch gz paths in input order)."""
    from oracle import oracle as O
    fof = os.path.join(wd, "in.txt")
    with open(fof, "w") as f:
        f.write("\n".join(paths) + "\n")
    t0 = time.perf_counter()
    subprocess.run([os.path.join(O.REF_DIR, "sub_sampler"), "-f", fof, "-k", str(k), "-m", str(m), "-s", str(s),
                    "-t", str(cores), "-v", "0"], cwd=wd, stdin=subprocess.DEVNULL, stdout=subprocess.DEVNULL, check=True,
                   preexec_fn=_all_cores)
    t = time.perf_counter() - t0
    return t, [os.path.join(wd, "subsampled_" + os.path.basename(p).split(".")[0] + ".gz") for p in paths]


def ref_compare(sketch_paths, wd, tag="res", queries=None):
    """oracle/_ref/comparator -> (wall seconds, its own 'Comparisons lasted' seconds, csv paths)."""
    from oracle import oracle as O
    skf = os.path.join(wd, tag + "_sk.txt")
    with open(skf, "w") as f:
        f.write("\n".join(sketch_paths) + "\n")
    cmd = [os.path.join(O.REF_DIR, "comparator"), "-f", skf, "-o", os.path.join(wd, tag)]
    if queries:
        qf = os.path.join(wd, tag + "_q.txt")
        with open(qf, "w") as f:
            f.write("\n".join(queries) + "\n")
        cmd += ["-q", qf]
    t0 = time.perf_counter()
    r = subprocess.run(cmd, cwd=wd, stdin=subprocess.DEVNULL, stdout=subprocess.PIPE, text=True, check=True, preexec_fn=_all_cores)
    t = time.perf_counter() - t0
    lasted = None
    for ln in r.stdout.splitlines():
        if ln.startswith("Comparisons lasted"):
            lasted = float(ln.split()[2])
    return t, lasted, (os.path.join(wd, tag + "_containment.csv.gz"), os.path.join(wd, tag + "_jaccard.csv.gz"))


def run_reference_step(paths, args, wd, cores):
    """One pass of the reference's own path: sub_sampler -f ... -t cores, then comparator.
    Returns (sketch_seconds, compare_seconds_total, 'Comparisons lasted' seconds, kind)."""
    from oracle import oracle as O
    for f in os.listdir(wd):
        if f.startswith("subsampled_") or f.startswith("res_"):
            os.remove(os.path.join(wd, f))
    if O.have_ref():
        t_sk, sk = ref_sketch(paths, args.k, args.m, args.s, wd, cores)
        t_c, lasted, _ = ref_compare(sk, wd)
        return t_sk, t_c, lasted, "reference"
    # C restatement (single-threaded per call; ctypes drops the GIL so threads scale)
    from concurrent.futures import ThreadPoolExecutor
    datas = [open(p, "rb").read() for p in paths]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        sks = list(ex.map(lambda d: O.sketch(d, args.k, args.m, args.s)[0], datas))
    t1 = time.perf_counter()
    O.compare(sks)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1, t2 - t1, "port"


def sketches_identical(ref_gz_paths, ours):
    import gzip
    n = 0
    for p, sk in zip(ref_gz_paths, ours):
        with gzip.open(p, "rb") as f:
            n += int(f.read() == sk)
    return n


def csv_identical(csv_paths, names, query_size, inter, full_rows, sizes, S):
    """Our containment / Jaccard CSV text (`-p 6`) against the files the reference comparator wrote."""
    import gzip
    out = {}
    for path, tag, jac in ((csv_paths[0], "containment", False), (csv_paths[1], "jaccard", True)):
        with gzip.open(path, "rb") as f:
            ref = f.read()
        ours = S.format_csv(names, query_size, inter, full_rows, sizes, jac, 6, 0.0)
        out[f"{tag}_csv_identical"] = bool(ours == ref)
    return out


def reference_parity(wd, names, sketches, cmp_res, S):
    """Byte parity of this step's outputs with the files the unmodified reference just wrote for the same
    inputs: every gunzipped sketch, and both CSV matrices (`-p 6`).  Checker only, outside any timed region."""
    sk_paths = [os.path.join(wd, "subsampled_" + nm + ".gz") for nm in names]
    out = {"sketches_identical": sketches_identical(sk_paths, sketches), "sketches": len(names)}
    inter, sizes, full = cmp_res
    out.update(csv_identical((os.path.join(wd, "res_containment.csv.gz"), os.path.join(wd, "res_jaccard.csv.gz")),
                             sk_paths, len(names), inter, full, sizes, S))
    out["ok"] = bool(out["sketches_identical"] == out["sketches"] and out["containment_csv_identical"]
                     and out["jaccard_csv_identical"])
    return out


def cli_end_to_end(paths, names, args, wd, cores, ref_seconds):
    """The drop-in executables as PROCESSES (cold start: driver + context initialisation included) on the files the
    reference binaries just processed: bin/sub_sampler -f + bin/comparator, wall clock, outputs compared byte for
    byte (after gunzip) with the reference's."""
    import gzip
    from supersampler_b200 import capi
    d = os.path.join(wd, "cli")
    os.makedirs(d, exist_ok=True)
    fof = os.path.join(d, "in.txt")
    with open(fof, "w") as f:
        f.write("\n".join(paths) + "\n")
    t0 = time.perf_counter()
    subprocess.run([os.path.join(capi.BIN_DIR, "sub_sampler"), "-f", fof, "-k", str(args.k), "-m", str(args.m), "-s", str(args.s),
                    "-t", str(cores), "-v", "0"], cwd=d, stdin=subprocess.DEVNULL, stdout=subprocess.DEVNULL, check=True)
    t1 = time.perf_counter()
    sk = ["subsampled_" + n_ + ".gz" for n_ in names]
    with open(os.path.join(d, "mine.txt"), "w") as f:
        f.write("\n".join(os.path.join(d, x) for x in sk) + "\n")
    subprocess.run([os.path.join(capi.BIN_DIR, "comparator"), "-f", os.path.join(d, "mine.txt"), "-o", "res"], cwd=d,
                   stdin=subprocess.DEVNULL, stdout=subprocess.DEVNULL, check=True)
    t2 = time.perf_counter()
    gun = lambda p_: gzip.open(p_, "rb").read()
    same_sk = sum(gun(os.path.join(d, x)) == gun(os.path.join(wd, x)) for x in sk)
    strip = lambda b_: b_.replace((d + "/").encode(), b"").replace((wd + "/").encode(), b"")       # header row = file names
    same_csv = all(strip(gun(os.path.join(d, f"res_{t}.csv.gz"))) == strip(gun(os.path.join(wd, f"res_{t}.csv.gz")))
                   for t in ("containment", "jaccard"))
    ours = t2 - t0
    return {"workload": f"C2 files on tmpfs: {len(paths)} x {args.bases} bp, sub_sampler -f -t {cores} -v 0 + comparator, as processes",
            "ours_s": ours, "ours_sub_sampler_s": t1 - t0, "ours_comparator_s": t2 - t1, "reference_s": ref_seconds,
            "ratio": ref_seconds / ours, "sketches_identical": int(same_sk), "sketches": len(sk), "csv_identical": bool(same_csv),
            "note": "cold processes: CUDA driver + context start-up (about 1 s each) is inside ours_s"}


def reference_arm(args, rank, world):
    """The reference's own CPU implementation on the SAME job as the B200 arm at this N: 64 x N genomes
    sketched with every host core and compared all-vs-all (single-threaded by construction)."""
    if rank != 0:
        return
    from oracle import oracle as O
    O.build(with_ref=os.path.isdir(O.REFERENCE_SRC))
    cores = os.cpu_count() or 1
    n_gen = args.genomes * max(1, args.gpus)
    fastas, names = make_fastas(n_gen, args.bases, 0, procs=cores)
    total_bases = n_gen * args.bases
    wd = scratch_dir()
    try:
        paths = write_files(fastas, names, wd)
        del fastas
        for _ in range(args.warmup):
            run_reference_step(paths, args, wd, cores)
        ts, tc, tl, kind = [], [], [], "reference"
        t_all0 = time.perf_counter()
        for _ in range(args.steps):
            a, b, c, kind = run_reference_step(paths, args, wd, cores)
            ts.append(a); tc.append(b); tl.append(c if c is not None else b)
        t_all = time.perf_counter() - t_all0
    finally:
        shutil.rmtree(wd, ignore_errors=True)
    step = t_all / args.steps
    pairs = n_gen * (n_gen - 1) // 2
    value = total_bases / step / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, max(1, args.gpus)),
        "sketch_only_gbp_per_s": total_bases / statistics.mean(ts) / 1e9,
        "compare": {"pairs": pairs, "pairs_per_s": pairs / statistics.mean(tl), "seconds": statistics.mean(tl),
                    "note": "reference comparator is single-threaded; time = its own 'Comparisons lasted' line"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"full workload of the {max(1, args.gpus)}-GPU job: {n_gen} x {args.bases} bp, "
                                   f"sub_sampler -f -t {cores} + comparator, files on tmpfs"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": f"C2: file-of-files of {args.genomes} synthetic {args.bases} bp genomes per GPU "
                        f"(one ancestor, substitution rate 10^U(-3,-1)), sketch + all-vs-all",
            "k": args.k, "m": args.m, "s": args.s, "genomes_per_gpu": args.genomes, "bases_per_genome": args.bases,
            "ranks": world,
            "l2": "device-resident inputs rotate over replicas totalling > 126 MB (L2) between timed iterations"}


# ------------------------------------------------------- extras: C3 / C4 / C5 at full size

class _DevArray:
    """__cuda_array_interface__ view of a raw device pointer (no copy, no ownership)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class ResidentSet:
    """Sketches of a job that runs as several device batches: the sketch bytes go to the host, the compare
    elements of every batch are kept on the device (torch tensors: plumbing) for ONE compare at the end."""

    def __init__(self, ctx, k, m, s):
        self.ctx, self.k, self.m, self.s = ctx, k, m, s
        self.sketches, self.sizes = [], []
        self.mn, self.lo = [], []
        self.scan_ms = self.post_ms = self.sketch_s = 0.0
        self.hits = self.bases = self.batches = 0

    def add_batch(self, d_words, n_bases, rec_begin, rec_end, rec_input, n_inputs, rec_device=None):
        import torch
        info = {}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sks = self.ctx.sketch_batch(None, n_bases, rec_begin, rec_end, rec_input, n_inputs, self.s,
                                    device_ptr=d_words.data_ptr(), info=info, rec_device=rec_device)
        off = np.asarray(info["elem_off"], np.int64)
        e = int(off[-1])
        if e:
            a, b, _ = self.ctx.batch_element_ptrs()
            dev = d_words.device
            self.mn.append(torch.as_tensor(_DevArray(a, e, "<i4"), device=dev).clone())
            self.lo.append(torch.as_tensor(_DevArray(b, e, "<i8"), device=dev).clone())
        torch.cuda.synchronize()
        self.sketch_s += time.perf_counter() - t0
        self.sketches += sks
        self.sizes.append(np.diff(off))
        self.scan_ms += info["scan_ms"]; self.post_ms += info["post_ms"]
        self.hits += info["n_hits"]; self.bases += n_bases; self.batches += 1

    def elements(self):
        import torch
        sizes = np.concatenate(self.sizes) if self.sizes else np.zeros(0, np.int64)
        mn = torch.cat(self.mn) if self.mn else torch.zeros(1, dtype=torch.int32, device="cuda")
        lo = torch.cat(self.lo) if self.lo else torch.zeros(1, dtype=torch.int64, device="cuda")
        return sizes, mn, lo

    def compare(self, query_size=None):
        """-> (inter, sizes, seconds, kernel_ms); all-vs-all (upper triangle valid) or rows of the first
        query_size sketches against all."""
        import torch
        sizes, mn, lo = self.elements()
        n = sizes.size
        sk_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
        first = None
        for _ in range(2):       # warm context, like the sketch half: the first call also allocates the compare buffers
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            self.ctx.cmp_load_device(sk_off, mn.data_ptr(), lo.data_ptr(), None)
            if query_size is None:
                inter = self.ctx.cmp_run((0, n), (0, n), True)
            else:
                inter = self.ctx.cmp_run((0, query_size), (0, n), False)
            t = time.perf_counter() - t0
            first = t if first is None else first
        self.compare_first_call_s = first
        return inter, sizes.astype(np.uint64), t, self.ctx.cmp_kernel_ms()


def _kernel_name(finfo):
    return {2: "scan_rowbit_kernel", 1: "scan_filter_kernel(byte table)", 0: "scan_filter_kernel(bit table)"}.get(
        finfo["kind"], "scan_dense_kernel")


def extra_c3(S, SD, rank, world, dist, local_rank, cores, quick):
    """BASELINE config 3 at its stated size: 1 024 x 5 Mbp, k31 m11 s=100, all-vs-all; one fixed job dealt over
    the ranks (strong scaling): rank r sketches genomes [r*1024/W, (r+1)*1024/W), the elements are exchanged,
    the 32x32 tiles are dealt round-robin.  Parity: reference sub_sampler on every genome (N=1, when the
    scratch space allows) or on the first 128, reference comparator on the first 128 sketches."""
    import torch
    from supersampler_b200 import distributed as D
    from oracle import oracle as O
    k, m, s = 31, 11, 100.0
    G, nb = (128, 1_000_000) if quick else (1024, 5_000_000)
    per = G // world
    g0 = rank * per
    fam = SD.DeviceFamily(nb, seed=4242)
    ctx = S.DeviceContext(k, m, S.threshold(k, m, s), device=local_rank)
    rs = ResidentSet(ctx, k, m, s)
    bsz = max(1, min(200, (1 << 30) // (SD.words_per_input(nb) * 16)))
    # warm context, as in a long-running job: tables built and buffers grown to this job's batch shape by one untimed
    # batch (the first one) before the clock starts
    nw = min(bsz, per)
    b0 = fam.packed_batch(g0, nw)
    ctx.sketch_batch(None, *b0[1:], nw, s, device_ptr=b0[0].data_ptr())
    del b0
    for a in range(g0, g0 + per, bsz):
        cnt = min(bsz, g0 + per - a)
        buf, n_total, rb, re_, ri = fam.packed_batch(a, cnt)
        rs.add_batch(buf, n_total, rb, re_, ri, cnt)
        del buf
    t_sk = torch.tensor([rs.sketch_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t_sk, op=dist.ReduceOp.MAX)
    # compare
    cinfo = {}
    if dist is None:
        inter, sizes, t_cmp, cmp_ms = rs.compare()
    else:
        sizes_l, mn, lo = rs.elements()
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        all_sizes, g_mn, g_lo, _ = D.exchange_tensors(sizes_l, mn, lo, None, False, torch.device("cuda", local_rank))
        inter, sizes, _ = D.compare_gathered(all_sizes, g_mn, g_lo, None, rank, world, ctx, cinfo)
        torch.cuda.synchronize()
        tt = torch.tensor([time.perf_counter() - t0, cinfo.get("kernel_ms", 0.0)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_cmp, cmp_ms = float(tt[0]), float(tt[1])
    if rank != 0:
        ctx.close()
        return None
    pairs = G * (G - 1) // 2
    t_sk = float(t_sk.item())
    keycmp = float((sizes.astype(np.float64).sum() * (G - 1)))          # sum over pairs of |K_i| + |K_j|
    first_cmp = getattr(rs, "compare_first_call_s", None)
    out = {"workload": f"C3: {G} x {nb} bp genomes, k{k} m{m} s{int(s)}, all-vs-all, one fixed job over {world} GPU(s); warm context "
                       f"(one untimed batch of the job's shape first)",
           "scaling": "strong", "value_gbp_per_s": G * nb / (t_sk + t_cmp) / 1e9, "sketch_gbp_per_s": G * nb / t_sk / 1e9,
           "sketch_s": t_sk, "compare_s": t_cmp, "compare_first_call_s": first_cmp, "scan_ms": rs.scan_ms, "postpass_ms": rs.post_ms,
           "compare_kernel_ms": cmp_ms, "batches_per_rank": rs.batches, "kernel": _kernel_name(ctx.filter_info()),
           "hits_per_rank": rs.hits, "elements": int(sizes.sum()), "pairs": pairs, "pairs_per_s": pairs / t_cmp,
           "kernel_pairs_per_s": pairs / max(1e-9, cmp_ms * 1e-3), "key_comparisons_per_s": keycmp / max(1e-9, cmp_ms * 1e-3)}
    # ---- parity against oracle/_ref (checker only, not timed)
    if O.have_ref():
        n_par = per if (world == 1 and scratch_free_bytes() > 2.5 * G * nb and not quick) else min(128, per)
        n_cmp = min(128, n_par)
        wd = scratch_dir()
        try:
            paths = []
            for a in range(0, n_par, 64):
                fas = fam.fasta(a, min(64, n_par - a))
                paths += write_files(fas, [f"g{a + i:05d}" for i in range(len(fas))], wd)
                del fas
            t_ref, sk_paths = ref_sketch(paths, k, m, s, wd, cores)
            for p_ in paths:
                os.remove(p_)
            ident = sketches_identical(sk_paths, rs.sketches[:n_par])
            t_rc, lasted, csvs = ref_compare(sk_paths[:n_cmp], wd)
            sub = np.ascontiguousarray(inter[:n_cmp, :n_cmp])
            par = {"sketches": n_par, "sketches_identical": ident, "compare_subset": n_cmp,
                   "reference_sketch_s": t_ref, "reference_sketch_gbp_per_s": n_par * nb / t_ref / 1e9,
                   "reference_compare_s": lasted if lasted else t_rc,
                   "reference_pairs_per_s": n_cmp * (n_cmp - 1) / 2 / (lasted if lasted else t_rc)}
            par.update(csv_identical(csvs, sk_paths[:n_cmp], n_cmp, sub, False, sizes[:n_cmp], S))
            par["ok"] = bool(ident == n_par and par["containment_csv_identical"] and par["jaccard_csv_identical"])
            out["parity"] = par
        finally:
            shutil.rmtree(wd, ignore_errors=True)
    ctx.close()
    return out


def extra_c4(S, SD, local_rank, cores, threads, quick):
    """BASELINE config 4: 1 Gbp read sets of 150 bp reads (6 666 667 records each), k31 m11 s1000.
    `value`: reads resident in HBM (packed + record tables), one batch per read set.  `from_files`: the
    public Pipeline on the FASTA files (read, clean, pack, H2D, scan, post-pass).  Parity: every sketch
    against oracle/_ref sub_sampler on the same files (the reference uses one thread per file)."""
    import torch
    from oracle import oracle as O
    k, m, s = 31, 11, 1000.0
    n_sets, n_reads = (2, 200_000) if quick else (4, 6_666_667)
    L = 150
    rsrc = SD.DeviceReadSet(50_000_000 if not quick else 2_000_000, L, seed=7)
    ctx = S.DeviceContext(k, m, S.threshold(k, m, s), device=local_rank)
    rs = ResidentSet(ctx, k, m, s)
    wd = scratch_dir()
    try:
        paths = []
        for i in range(n_sets):
            codes = rsrc.codes(i, n_reads)
            buf, n_total, b, e, inp = rsrc.packed(codes)
            if i == 0:       # warm the context (tables, buffers) outside the timed calls
                ctx.sketch_batch(None, n_total, None, None, None, 1, s, device_ptr=buf.data_ptr(),
                                 rec_device=(b.data_ptr(), e.data_ptr(), inp.data_ptr(), n_reads))
            rs.add_batch(buf, n_total, None, None, None, 1,
                         rec_device=(b.data_ptr(), e.data_ptr(), inp.data_ptr(), n_reads))
            p = os.path.join(wd, f"reads{i:02d}.fa")
            with open(p, "wb") as f:
                f.write(rsrc.fasta(codes))
            paths.append(p)
            del codes, buf, b, e, inp
        torch.cuda.empty_cache()
        bases = n_sets * n_reads * L
        out = {"workload": f"C4: {n_sets} read sets of {n_reads} x {L} bp reads (2-line FASTA), k{k} m{m} s{int(s)}",
               "value_gbp_per_s": bases / rs.sketch_s / 1e9, "sketch_s": rs.sketch_s, "scan_ms": rs.scan_ms,
               "postpass_ms": rs.post_ms, "records_per_set": n_reads, "hits": rs.hits,
               "kernel": _kernel_name(ctx.filter_info()),
               "scan_tbp_per_s": rs.bases / max(1e-9, rs.scan_ms * 1e-3) / 1e12}
        ctx.close()
        # the public pipeline on the files: text cleaned + packed by host threads (one worker per file: 4 of them
        # have work here), and by the device (ingest kernels; the workers only read the files into pinned chunks)
        out["from_files"] = {}
        agree = True
        for mode in ("host", "device"):
            pl = S.Pipeline(k, m, s, device=local_rank, threads=threads, ingest=mode)
            pl.sketch(paths[:1])
            info = {}
            t0 = time.perf_counter()
            sks_files = pl.sketch(paths, info=info)
            t_files = time.perf_counter() - t0
            pl.close()
            out["from_files"][mode] = {"gbp_per_s": bases / t_files / 1e9, "seconds": t_files, "host_threads": threads,
                                       "pack_s": info.get("pack_s"), "device_s": info.get("device_s"),
                                       "ingest_kernels_ms": info.get("ingest_ms"),
                                       "api": f"supersampler_b200.Pipeline(ingest='{mode}').sketch(paths of the FASTA files on tmpfs)"}
            agree = agree and bool(sks_files == rs.sketches)
        out["routes_agree"] = agree
        if O.have_ref():
            t_ref, sk_paths = ref_sketch(paths, k, m, s, wd, cores)
            ident = sketches_identical(sk_paths, rs.sketches)
            out["parity"] = {"sketches": n_sets, "sketches_identical": ident, "reference_sketch_s": t_ref,
                             "reference_sketch_gbp_per_s": bases / t_ref / 1e9,
                             "reference_threads_used": min(cores, n_sets), "ok": bool(ident == n_sets and out["routes_agree"])}
    finally:
        shutil.rmtree(wd, ignore_errors=True)
    return out


def extra_c5(S, SD, rank, world, dist, local_rank, cores, quick):
    """BASELINE config 5: 100 query sketches vs 10 000 reference sketches, k31 m13 s200, `-q` mode:
    10 100 x 5 Mbp genomes sketched from HBM in 1 Gbp batches, Q x N compare on the device.  On several GPUs the
    job is dealt over the ranks (rank 0: the queries + its share of the references, the others their shares; the
    union is in the comparator's order, queries first) and compared by the query-mode exchange of the C ABI
    (spsp_cmp_exchange: only the Q x N rows exist anywhere).
    Parity: reference sub_sampler on the 100 queries + the first 156 references, reference
    `comparator -q` on those 256; our Q x 256 block of the full answer must give the same CSV bytes."""
    import torch
    from supersampler_b200 import distributed as D
    from oracle import oracle as O
    k, m, s = 31, 13, 200.0
    Q, R, nb = (10, 150, 1_000_000) if quick else (100, 10_000, 5_000_000)
    N = Q + R
    per = R // world
    g0 = 0 if rank == 0 else Q + rank * per                       # first genome of this rank
    n_loc = (Q + per if rank == 0 else per) + (R - per * world if rank == world - 1 else 0)
    fam = SD.DeviceFamily(nb, seed=555)
    ctx = S.DeviceContext(k, m, S.threshold(k, m, s), device=local_rank)
    rs = ResidentSet(ctx, k, m, s)
    bsz = max(1, min(200, (1 << 30) // (SD.words_per_input(nb) * 16)))
    nw = min(bsz, n_loc)                                 # warm context: one untimed batch of the job's shape
    b0 = fam.packed_batch(g0, nw)
    ctx.sketch_batch(None, *b0[1:], nw, s, device_ptr=b0[0].data_ptr())
    del b0
    for a in range(g0, g0 + n_loc, bsz):
        cnt = min(bsz, g0 + n_loc - a)
        buf, n_total, rb, re_, ri = fam.packed_batch(a, cnt)
        rs.add_batch(buf, n_total, rb, re_, ri, cnt)
        del buf
    t_sk = rs.sketch_s
    first_cmp = None
    if dist is None:
        inter, sizes, t_cmp, cmp_ms = rs.compare(query_size=Q)
        first_cmp = rs.compare_first_call_s
    else:
        D.join_contexts(ctx, rank, world)
        sizes_l, mn, lo = rs.elements()
        for _ in range(2):                               # warm context: the first call also sizes the exchange buffers
            cinfo = {}
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            inter, sizes = ctx.cmp_exchange(sizes_l, mn.data_ptr(), lo.data_ptr(), None, Q if rank == 0 else 0, Q, N, rank, cinfo)
            tt = torch.tensor([time.perf_counter() - t0, cinfo.get("kernel_ms", 0.0), t_sk], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            first_cmp = float(tt[0]) if first_cmp is None else first_cmp
        t_cmp, cmp_ms, t_sk = float(tt[0]), float(tt[1]), float(tt[2])
        if rank != 0:
            ctx.close()
            return None
    pairs = Q * R
    out = {"workload": f"C5: {Q} queries vs {R} references ({N} x {nb} bp genomes), k{k} m{m} s{int(s)}, -q mode"
                       + (f", one fixed job over {world} GPUs (query-mode exchange)" if world > 1 else ""),
           "sketch_gbp_per_s": N * nb / t_sk / 1e9, "sketch_s": t_sk, "scan_ms": rs.scan_ms,
           "postpass_ms": rs.post_ms, "batches": rs.batches, "kernel": _kernel_name(ctx.filter_info()),
           "elements": int(sizes.sum()), "compare_s": t_cmp, "compare_first_call_s": first_cmp,
           "compare_kernel_ms": cmp_ms, "query_ref_pairs": pairs,
           "pairs_per_s": pairs / t_cmp, "kernel_pairs_per_s": pairs / max(1e-9, cmp_ms * 1e-3),
           "value_gbp_per_s": N * nb / (t_sk + t_cmp) / 1e9}
    if world > 1:
        out["scaling"] = "strong"
    if O.have_ref():
        n_ref = min(156, R)
        sel = list(range(Q + n_ref))                    # queries first, then references: the comparator's order
        wd = scratch_dir()
        try:
            paths = []
            for a in range(0, len(sel), 64):
                fas = fam.fasta(a, min(64, len(sel) - a))
                paths += write_files(fas, [f"g{a + i:05d}" for i in range(len(fas))], wd)
                del fas
            t_ref, sk_paths = ref_sketch(paths, k, m, s, wd, cores)
            for p_ in paths:
                os.remove(p_)
            ident = sketches_identical(sk_paths, [rs.sketches[i] for i in sel])
            t_rc, lasted, csvs = ref_compare(sk_paths[Q:], wd, queries=sk_paths[:Q])
            sub = np.ascontiguousarray(inter[:, :len(sel)])
            par = {"sketches": len(sel), "of": N, "sketches_identical": ident, "compare_subset": f"{Q} x {len(sel)}",
                   "reference_sketch_s": t_ref, "reference_sketch_gbp_per_s": len(sel) * nb / t_ref / 1e9,
                   "reference_compare_s": lasted if lasted else t_rc}
            par.update(csv_identical(csvs, sk_paths, Q, sub, True, sizes[:len(sel)], S))
            par["ok"] = bool(ident == len(sel) and par["containment_csv_identical"] and par["jaccard_csv_identical"])
            out["parity"] = par
        finally:
            shutil.rmtree(wd, ignore_errors=True)
    ctx.close()
    return out


def run_extras(S, rank, world, dist, local_rank, cores, threads, args):
    """C3 / C4 / C5 of BASELINE.json at their stated sizes, run once, outside the headline timed regions.
    N > 1: C3 and C5, each as one fixed job over the ranks (the strong-scaling figures); C4 has no exchange."""
    import torch
    from supersampler_b200 import synth_device as SD
    extra = {}
    todo = [c for c in args.extras.split(",") if c]
    for name in todo:
        if name == "c4" and world > 1:               # (a sketch-only config: it shards by file, nothing to exchange)
            continue
        t0 = time.perf_counter()
        try:
            if name == "c3":
                r = extra_c3(S, SD, rank, world, dist, local_rank, cores, args.quick_extras)
            elif name == "c4":
                r = extra_c4(S, SD, local_rank, cores, threads, args.quick_extras)
            elif name == "c5":
                r = extra_c5(S, SD, rank, world, dist, local_rank, cores, args.quick_extras)
            else:
                continue
        except Exception as ex:                      # an extra must not take the headline line with it
            import traceback
            log(traceback.format_exc())
            r = {"error": f"{type(ex).__name__}: {ex}"}
            if dist is not None:
                raise
        torch.cuda.empty_cache()
        if r is not None:
            r["wall_s"] = time.perf_counter() - t0
            extra[name] = r
            log(f"[extra {name}] {json.dumps(r)}")
    return extra


# ------------------------------------------------------------- B200 arm

def b200_arm(args, rank, world, local_rank):
    import torch
    import supersampler_b200 as S
    from supersampler_b200 import distributed as D

    # host threads that wait for the device yield their core instead of spinning: measured equal to spinning with a
    # core per waiter (N=1, 16 cores) and ahead of it with 4 cores per rank (8 GPUs on a 32-core box)
    os.environ.setdefault("SPSP_SCHED", "yield")
    # several ranks on one box: each one (its pack workers, its pinned buffers) on its GPU's own NUMA node
    numa = {"bound": False}
    if world > 1 and not args.no_numa_bind:
        numa = D.bind_rank_to_gpu_cores(local_rank, world)
    torch.cuda.set_device(local_rank)
    dist = None
    saved_stdout = None
    if world > 1:
        # libraries (NCCL's version banner) write to fd 1: keep stdout clean for the ONE JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    cores = os.cpu_count() or 1
    threads = args.threads if args.threads > 0 else max(1, min(32, numa["n_cpus"] if numa.get("bound") else cores // world))
    k, m, s = args.k, args.m, args.s
    fastas, names = make_fastas(args.genomes, args.bases, rank * args.genomes)
    total_bases_rank = sum(args.bases for _ in fastas)
    n_in = len(fastas)

    # Both paths run the batches of consecutive steps one-deep pipelined, as a long job would: two device
    # contexts on the GPU; while one sketches batch i+1 the other finishes the compare stage of batch i on a
    # background thread (BatchStream).  Every step still is one full pass (sketch + compare) over its batch; the
    # last compare is drained inside the timed region.  --no-pipelining runs the stages back to back.
    from concurrent.futures import ThreadPoolExecutor
    depth = 1 if args.no_pipelining else 2
    rdepth = 1 if args.no_pipelining else max(2, args.depth)     # device-resident path: batches in flight
    if args.cmp_depth <= 0:
        args.cmp_depth = 2 if world < 4 else min(4, rdepth)
    w_req = args.warmup
    args.warmup = max(args.warmup, rdepth)      # every context has run (tables, buffers) before the clock starts
    # host-buffer path: the public pipeline (pack on host threads -> pinned -> H2D -> scan -> post-pass -> compare)
    pipes = [S.Pipeline(k, m, s, device=local_rank, threads=threads, ingest=args.ingest) for _ in range(depth)]
    # host buffers of the e2e path: page-locked when part of the text goes to the device as it is (asynchronous
    # copies straight from where it lies); plain bytes when host threads pack all of it
    fastas_e2e = [S.PinnedBuffer(fa) for fa in fastas] if args.ingest != "host" else fastas
    pctxs = [p_.device_context() for p_ in pipes]
    # device-resident path: contexts of their own; all genomes packed back to back, R replicas in HBM
    dctxs = [S.DeviceContext(k, m, S.threshold(k, m, s), device=local_rank) for _ in range(rdepth)]
    ws, ros = [], []
    for fa in fastas:
        w, nb, offs = S.pack_fasta(fa, k)
        ws.append(w); ros.append(offs)
    packed, n_total, rec_begin, rec_end, rec_input = S.batch_layout(ws, ros)
    del ws
    packed_bytes = packed.size * 4
    replicas = max(2, int(np.ceil(160e6 / packed_bytes)) + 1)
    h_packed = torch.from_numpy(packed.view(np.int32))
    d_packed = [h_packed.cuda() for _ in range(replicas)]
    torch.cuda.synchronize()

    stats = {"scan_ms": [], "post_ms": [], "cmp_ms": [], "hits": 0, "launches": 0, "sketch_s": [], "compare_s": [],
             "d2h": 0, "e2e": []}

    if dist is not None:
        for c_ in dctxs + pctxs:
            D.join_contexts(c_, rank, world)
    use_native = os.environ.get("SPSP_BENCH_TORCH_EXCHANGE", "0") != "1"

    def compare_device(ctx, elem_off, cinfo):
        """Compare stage from the elements the batch left on `ctx`'s device."""
        t0 = time.perf_counter()
        if dist is not None and use_native:
            res = D.native_exchange_compare(ctx, n_in, rank, world, cinfo)
        elif dist is None:
            l0 = ctx.launches()
            ctx.cmp_load_batch()
            inter = ctx.cmp_run((0, n_in), (0, n_in), True)
            cinfo.update(kernel_ms=ctx.cmp_kernel_ms(), launches=ctx.launches() - l0)
            res = inter, np.diff(np.asarray(elem_off, np.uint64)), False
        else:
            res = D.allgather_compare_device(elem_off, ctx, rank, world, cinfo)
        cinfo["seconds"] = time.perf_counter() - t0
        return res

    class Resident:
        """Device-resident steps, `rdepth` batches in flight: the sketch of batch i runs on context i % rdepth from
        a pool of host threads (batches on different streams overlap on the device); the compare stages run in
        step order on `cdepth` threads (each context has its own NCCL communicator, so exchanges of different
        contexts may be in flight together).  A context is reused only after its previous batch has been retired."""

        def __init__(self):
            self.sk_pool = ThreadPoolExecutor(rdepth)
            self.cmp_pool = ThreadPoolExecutor(max(1, args.cmp_depth))
            self.inflight = []

        def _sketch(self, i):
            ctx = dctxs[i % rdepth]
            t0 = time.perf_counter()
            info = {}
            l0 = ctx.launches()
            sks = ctx.sketch_batch(None, n_total, rec_begin, rec_end, rec_input, n_in, s,
                                   device_ptr=d_packed[i % replicas].data_ptr(), info=info)
            return sks, info, ctx.launches() - l0, time.perf_counter() - t0

        def _compare(self, i, fut):
            sks, info, nl, t_sk = fut.result()
            cinfo = {}
            res = compare_device(dctxs[i % rdepth], info["elem_off"], cinfo)
            return sks, res, info, cinfo, nl, t_sk

        def _retire(self, record):
            sks, res, info, cinfo, nl, t_sk = self.inflight.pop(0).result()
            if record:
                stats["scan_ms"].append(info["scan_ms"]); stats["post_ms"].append(info["post_ms"])
                stats["cmp_ms"].append(cinfo.get("kernel_ms", 0.0))
                stats["hits"] = info["n_hits"]
                stats["launches"] += nl + cinfo.get("launches", 0)
                stats["sketch_s"].append(t_sk); stats["compare_s"].append(cinfo["seconds"])
                stats["d2h"] = sks.nbytes + (res[0].size * 4 if res[0] is not None else 0)
            return sks, res

        def step(self, i, record):
            out = self._retire(record) if len(self.inflight) >= rdepth else None
            fut = self.sk_pool.submit(self._sketch, i)
            self.inflight.append(self.cmp_pool.submit(self._compare, i, fut))
            return out

        def finish(self, record):
            out = None
            while self.inflight:
                out = self._retire(record)
            return out

    class HostBuffers:
        """e2e steps through the public API: BatchStream over the two pipelines."""

        def __init__(self):
            def cmp_fn(pl, cinfo):
                if dist is None:
                    return pl.compare(info=cinfo)
                off, on_dev = pl.elem_off()
                assert on_dev
                return compare_device(pctxs[pipes.index(pl)], off, cinfo)
            self.stream = S.BatchStream(k, m, s, compare_fn=cmp_fn, pipelines=pipes + (pipes if depth == 1 else []))

        @staticmethod
        def _note(done, record):
            if done is None:
                return None
            sks, res, info, cinfo = done
            if record:
                info["cmp_kernel_ms"] = cinfo.get("kernel_ms"); info["launches"] += cinfo.get("launches", 0)
                info["d2h_bytes"] += res[0].size * 4 if res[0] is not None else 0
                stats["e2e"].append(info)
            return sks, res

        def step(self, i, record):
            out = self._note(self.stream.submit(fastas_e2e), record)
            if depth == 1:
                out = self._note(self.stream.drain(), record)
            return out

        def finish(self, record):
            return self._note(self.stream.drain(), record)

    def timed(runner, steps, warmup, ctx, min_total=0.5, max_regions=400):
        """Timed regions of EXACTLY `steps` steps each, every one bracketed by barrier + synchronize, CUDA events on
        a stream the kernels are launched on, the pipeline drained inside the region (every step's results have
        reached the host), max over ranks.  Regions are repeated until >= min_total seconds have been measured
        (a 20-step region of this workload lasts a few milliseconds); the caller reports the median region."""
        ext = torch.cuda.ExternalStream(ctx.stream(0))
        for i in range(warmup):
            runner.step(i, False)
        runner.finish(False)
        regions, out, n_done, total = [], None, warmup, 0.0
        while True:
            barrier(); torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            ev0.record(ext)
            for i in range(steps):
                o = runner.step(n_done + i, True)
                out = o if o is not None else out
            o = runner.finish(True)
            out = o if o is not None else out
            ev1.record(ext)
            torch.cuda.synchronize(); barrier()
            wall = time.perf_counter() - t0
            dev = ev0.elapsed_time(ev1) / 1e3
            t = torch.tensor([max(wall, dev)], dtype=torch.float64, device="cuda")
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            n_done += steps
            regions.append(float(t.item()))              # identical on every rank: so is the decision to stop
            total += regions[-1]
            if total >= min_total or len(regions) >= max_regions:
                return regions, out

    # nvidia-smi needs ~0.2 s to start reporting: sample over both timed regions (device busy throughout)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    reg_res, (sks_res, cmp_res) = timed(Resident(), args.steps, args.warmup, dctxs[0], args.min_seconds)
    reg_e2e, (sks_e2e, cmp_e2e) = timed(HostBuffers(), args.steps, args.warmup, pctxs[0], args.min_seconds)
    t_res, t_e2e = statistics.median(reg_res), statistics.median(reg_e2e)
    # third region: the scan kernel alone (what roofline.achieved is quoted on): one launch per replica in turn
    # (> L2 between launches), CUDA events on the launching stream around each launch
    scan_alone = []
    if rank == 0:
        cap = int(n_total * S.threshold(k, m, s) / 2.0 ** 64 * 1.5) + 65536
        d_hits = torch.empty(cap * 16, dtype=torch.uint8, device="cuda")
        d_cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
        n_alone = max(10, args.steps)
        for i in range(args.warmup + n_alone):
            dctxs[0].scan_device(d_packed[i % replicas].data_ptr(), n_total, d_hits.data_ptr(), cap, d_cnt.data_ptr())
            dctxs[0].sync()
            if i >= args.warmup:
                scan_alone.append(dctxs[0].scan_kernel_ms())
        assert int(d_cnt.item()) == stats["hits"], "scan alone and scan inside the step disagree on the hit count"
        del d_hits
    clocks = sampler.stop() if rank == 0 else None

    # both paths must produce the same bytes / counts
    assert sks_res == sks_e2e, "device-resident and host-buffer paths disagree"
    assert np.array_equal(cmp_res[1], cmp_e2e[1])
    if rank == 0:
        assert np.array_equal(np.triu(cmp_res[0], 1), np.triu(cmp_e2e[0], 1))
    if dist is not None and use_native:
        # the exchange inside the C ABI must give what the torch.distributed exchange gives
        last = pipes[0]
        chk = D.allgather_compare_device(last.elem_off()[0], pctxs[0], rank, world, {})
        assert np.array_equal(chk[1], cmp_res[1])
        if rank == 0:
            assert np.array_equal(np.triu(chk[0], 1), np.triu(cmp_res[0], 1)), "native and torch exchange disagree"

    # ---- parity against the unmodified reference at this N (checker only, outside every timed region):
    # every rank's FASTA files go to one scratch directory, rank 0 runs oracle/_ref over the 64 x N genomes and
    # holds every sketch of the job and the N x N matrix of the last step against the reference's files.
    parity, cpu_base, cli_e2e = None, None, None
    from oracle import oracle as O
    if not args.no_cpu_baseline and (O.have_ref() or world == 1):
        wd_box = [scratch_dir() if rank == 0 else None]
        if dist is not None:
            dist.broadcast_object_list(wd_box, src=0)
        wd = wd_box[0]
        try:
            my_paths = write_files(fastas, names, wd)
            all_sks, all_names = [list(sks_res)], [names]
            if dist is not None:
                all_sks, all_names = [None] * world, [None] * world
                dist.all_gather_object(all_sks, list(sks_res))
                dist.all_gather_object(all_names, names)
                barrier()
            if rank == 0:
                names_all = [n_ for part in all_names for n_ in part]
                paths = [os.path.join(wd, n_ + ".fa") for n_ in names_all]
                a, b, c, kind = run_reference_step(paths, args, wd, cores)
                if kind == "reference":
                    parity = reference_parity(wd, names_all, [x for part in all_sks for x in part], cmp_res, S)
                    parity["n_gpus"] = world
                    parity["checked"] = (f"all {len(names_all)} sketches of the {world}-GPU job and the "
                                         f"{len(names_all)} x {len(names_all)} matrix of the last timed step")
                if kind == "reference" and world == 1:
                    cli_e2e = cli_end_to_end(paths, names_all, args, wd, cores, a + b)
                tb = args.bases * len(names_all)
                prs = len(names_all) * (len(names_all) - 1) // 2
                cpu_base = {"value": tb / (a + b) / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
                            "sample": f"full workload once: {len(names_all)} x {args.bases} bp, sub_sampler -f -t {cores} "
                                      f"({a:.2f} s) + comparator ({b:.2f} s, single-threaded by construction)",
                            "sketch_gbp_per_s": tb / a / 1e9, "compare_pairs_per_s": prs / (c if c else b)}
            barrier()
        finally:
            if rank == 0:
                shutil.rmtree(wd, ignore_errors=True)

    finfo = dctxs[0].filter_info()
    # free the headline's device state before the extras
    del d_packed
    for p_ in pipes:
        p_.close()
    for c_ in dctxs:
        c_.close()
    torch.cuda.empty_cache()
    extra = run_extras(S, rank, world, dist, local_rank, cores, threads, args) if args.extras else {}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    n_gen_total = args.genomes * world
    total_bases = total_bases_rank * world
    pairs = n_gen_total * (n_gen_total - 1) // 2
    step_s = t_res / args.steps
    val = lambda t: total_bases / (t / args.steps) / 1e9
    # the scan kernel's launch duration (CUDA events on its launching stream): quoted on the kernel-alone region;
    # inside the two pipelined regions it shares the SMs / the copy engines with other batches, those per-launch
    # figures are kept beside it
    scan_ms_value_region = statistics.mean(stats["scan_ms"])
    scan_ms_e2e_region = statistics.mean(x["scan_ms"] for x in stats["e2e"])
    scan_ms = statistics.mean(scan_alone)
    scan_kernel_name = {2: "scan_rowbit_kernel (bank-private bit table + hashed m-mer table, DESIGN.md 3.3b)",
                        1: "scan_filter_kernel (byte table of phase masks, DESIGN.md 3.3)",
                        0: "scan_filter_kernel (bit table, DESIGN.md 3.3)"}.get(finfo["kind"], "scan_dense_kernel")
    algo_bytes = n_total / 4 + 16 * stats["hits"]
    achieved = algo_bytes / (scan_ms * 1e-3) / 1e9
    sizes = cmp_res[1]
    e2 = stats["e2e"]
    mean = lambda key: statistics.mean(x[key] for x in e2)
    cmp_kernel_ms = statistics.mean(stats["cmp_ms"])
    keycmp = float(sizes.astype(np.float64).sum()) * (n_gen_total - 1)        # sum over pairs of |K_i| + |K_j|
    cmp_ncu = measured_compare_pipes()
    line = {
        "metric": METRIC, "value": val(t_res), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": w_req, "warmup_done": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": workload_config(args, world),
        "clocks": clocks,
        "regions": {"note": f"timed regions of exactly {args.steps} steps each, repeated until >= {args.min_seconds} s were "
                            "measured; value / e2e / ms_per_step are the MEDIAN region",
                    "value": {"n": len(reg_res), "median": val(t_res), "min": val(max(reg_res)), "max": val(min(reg_res))},
                    "e2e": {"n": len(reg_e2e), "median": val(t_e2e), "min": val(max(reg_e2e)), "max": val(min(reg_e2e))}},
        "e2e": {"value": val(t_e2e), "unit": UNIT,
                "h2d_bytes_per_step": int(mean("h2d_bytes")), "d2h_bytes_per_step": int(mean("d2h_bytes")),
                "ms_per_step": t_e2e / args.steps * 1e3, "host_threads": threads, "numa_binding": numa,
                "ingest": {"mode": args.ingest, "inputs_cleaned_on_device_per_step": mean("text_inputs"),
                           "of": n_in, "ingest_kernels_ms": mean("ingest_ms"),
                           "note": "host = host threads clean + pack to 2-bit (0.25 B per base over PCIe); device = raw FASTA "
                                   "text over PCIe, clean + pack kernels (csrc/device/ingest.cu); auto = both on one work queue"},
                "phases_ms": {"pack_and_h2d": mean("pack_s") * 1e3, "device": mean("device_s") * 1e3,
                              "assemble": mean("assemble_s") * 1e3, "scan_kernel": mean("scan_ms"),
                              "postpass_device": mean("post_ms"), "compare_kernel": mean("cmp_kernel_ms")},
                "gpu_launches": int(sum(x["launches"] for x in e2) / max(1, len(reg_e2e))),
                "api": "supersampler_b200.BatchStream.submit(FASTA bytes in host memory) over two Pipelines "
                       "(.sketch() + .compare(), the compare of batch i overlapping the sketch of batch i+1)"},
        "pipelining": "off" if depth == 1 else
                      f"value: {rdepth} batches in flight (sketches on {rdepth} contexts / streams, compare stages on "
                      f"{max(1, args.cmp_depth)} thread(s)); "
                      "e2e: one batch deep (compare(i) on a background thread / second context while sketch(i+1) runs)",
        "gpu_launches": int(stats["launches"] / max(1, len(reg_res))),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": measured_traffic(n_total)[0], "traffic_source": measured_traffic(n_total)[1],
                     "algorithmic_bytes": int(algo_bytes), "peak_source": peak_src, "kernel": scan_kernel_name,
                     "kernel_ms": scan_ms, "bases_per_launch": int(n_total), "hits_per_launch": int(stats["hits"]),
                     "kernel_tbp_per_s": n_total / (scan_ms * 1e-3) / 1e12,
                     "kernel_ms_timed_in": f"kernel-alone region: {len(scan_alone)} launches over the rotating replicas, CUDA events on the launching stream",
                     "kernel_ms_in_value_region": scan_ms_value_region, "kernel_ms_in_e2e_region": scan_ms_e2e_region,
                     "kernel_share_of_step": scan_ms / (step_s * 1e3),
                     "step_frac": algo_bytes * world / step_s / 1e9 / (peak * world),
                     "step_frac_note": "the same algorithmic bytes over the whole `value` step (scan + post-pass + compare, "
                                       "per GPU) instead of over the scan kernel alone",
                     "ncu_pipes_pct_of_peak": measured_traffic(n_total)[2],
                     "note": "the kernel that streams every input byte; the rest of the step works on n/s-sized data "
                             "(latency-bound post-pass, INT/LSU-bound compare), see DESIGN.md section 6"},
        "phases_ms": {"sketch": statistics.mean(stats["sketch_s"]) * 1e3,
                      "compare": statistics.mean(stats["compare_s"]) * 1e3,
                      "scan_kernel": scan_ms_value_region, "postpass_device": statistics.mean(stats["post_ms"]),
                      "compare_kernel": cmp_kernel_ms},
        "d2h_bytes_per_step": int(stats["d2h"]),
        "compare": {"pairs": pairs, "pairs_per_s": pairs / statistics.mean(stats["compare_s"]),
                    "kernel_pairs_per_s": pairs / max(1e-9, cmp_kernel_ms * 1e-3),
                    "elements": int(sizes.sum()),
                    "roofline": {"bound": "INT/LSU (shared-memory hash join), not HBM",
                                 "key_comparisons_per_s": keycmp / max(1e-9, cmp_kernel_ms * 1e-3),
                                 "unit_note": "one pair = |K_i| + |K_j| key comparisons (SURVEY 8d); kernel time = max over "
                                              "ranks of hashjoin_kernel on this step's tiles",
                                 **cmp_ncu}},
        "parity_vs_reference": parity,
    }
    if cpu_base is not None and world == 1:
        line["cpu_baseline"] = cpu_base
    elif cpu_base is not None:
        line["reference_same_job"] = cpu_base
    if cli_e2e is not None:
        line["cli_e2e"] = cli_e2e
    if extra:
        line["extra"] = extra
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--genomes", type=int, default=64)
    ap.add_argument("--bases", type=int, default=5_000_000)
    ap.add_argument("-k", type=int, default=31)
    ap.add_argument("-m", type=int, default=11)
    ap.add_argument("-s", type=float, default=1000.0)
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the reference run (cpu_baseline + parity)")
    ap.add_argument("--no-pipelining", action="store_true", help="run sketch and compare of a step back to back")
    ap.add_argument("--depth", type=int, default=4, help="device-resident path: batches in flight")
    ap.add_argument("--cmp-depth", type=int, default=0,
                    help="device-resident path: compare stages in flight (0 = 2 on one or two GPUs, 4 from four ranks on: an "
                         "exchange is latency-bound and every context has its own communicator)")
    ap.add_argument("--ingest", default=os.environ.get("SPSP_INGEST", "auto"), choices=["host", "device", "auto"],
                    help="e2e path: who cleans + packs the FASTA text (host threads, the device, or both on one work queue)")
    ap.add_argument("--no-numa-bind", action="store_true", help="N > 1: do not confine a rank to the cores local to its GPU")
    ap.add_argument("--threads", type=int, default=0, help="host packing threads per rank (default: cores / ranks)")
    ap.add_argument("--min-seconds", type=float, default=0.5, help="repeat the K-step timed region until this much was measured")
    ap.add_argument("--extras", default="c3,c4,c5", help="full-size BASELINE configs run once after the headline ('' = none)")
    ap.add_argument("--quick-extras", action="store_true", help="small shapes of the extras (smoke run)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    import supersampler_b200 as S
    if rank == 0:
        S.build()
    b200_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
