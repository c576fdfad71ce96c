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
          same files / parameters; C restatement if the binaries are absent.

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


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_traffic(bases_per_launch):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (same workload)."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        with open(p) as f:
            t = json.load(f)
        if abs(bases_per_launch - 320012288) < 1e6:
            return t["traffic_bytes_per_launch"], t["source"], t.get("pipes_pct_of_peak")
    except Exception:
        pass
    return None, None, None


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

def make_fastas(n_genomes: int, n_bases: int, first: int, seed: int = 42):
    from supersampler_b200 import synth
    fam = synth.Family(n_bases, seed)
    return [fam.fasta(first + i) for i in range(n_genomes)], [f"g{first + i:05d}" for i in range(n_genomes)]


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


# ------------------------------------------------------- reference CPU path

def run_reference_step(paths, args, wd, cores):
    """One pass of the reference's own path: sub_sampler -f ... -t cores, then comparator.
    Returns (sketch_seconds, compare_seconds_total, 'Comparisons lasted' seconds, kind)."""
    from oracle import oracle as O
    for f in os.listdir(wd):
        if f.startswith("subsampled_") or f.startswith("res_"):
            os.remove(os.path.join(wd, f))
    if O.have_ref():
        fof = os.path.join(wd, "in.txt")
        with open(fof, "w") as f:
            f.write("\n".join(paths) + "\n")
        t0 = time.perf_counter()
        subprocess.run([os.path.join(O.REF_DIR, "sub_sampler"), "-f", fof, "-k", str(args.k), "-m", str(args.m),
                        "-s", str(args.s), "-t", str(cores), "-v", "0"], cwd=wd, stdin=subprocess.DEVNULL,
                       stdout=subprocess.DEVNULL, check=True)
        t1 = time.perf_counter()
        sk = [os.path.join(wd, "subsampled_" + os.path.basename(p).split(".")[0] + ".gz") for p in paths]
        skf = os.path.join(wd, "sk.txt")
        with open(skf, "w") as f:
            f.write("\n".join(sk) + "\n")
        r = subprocess.run([os.path.join(O.REF_DIR, "comparator"), "-f", skf, "-o", os.path.join(wd, "res")], cwd=wd,
                           stdin=subprocess.DEVNULL, stdout=subprocess.PIPE, text=True, check=True)
        t2 = time.perf_counter()
        lasted = None
        for ln in r.stdout.splitlines():
            if ln.startswith("Comparisons lasted"):
                lasted = float(ln.split()[2])
        return t1 - t0, t2 - t1, lasted, "reference"
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


def reference_parity(wd, names, sketches, cmp_res, S):
    """Byte parity of this step's outputs with the files the unmodified reference just wrote for the same
    inputs: every gunzipped sketch, and both CSV matrices (`-p 6`).  Checker only, outside any timed region."""
    import gzip
    out = {"sketches_identical": 0, "sketches": len(names)}
    for nm, sk in zip(names, sketches):
        with gzip.open(os.path.join(wd, "subsampled_" + nm + ".gz"), "rb") as f:
            out["sketches_identical"] += int(f.read() == sk)
    inter, sizes, full = cmp_res
    csv_names = [os.path.join(wd, "subsampled_" + nm + ".gz") for nm in names]
    for tag, jac in (("containment", False), ("jaccard", True)):
        with gzip.open(os.path.join(wd, f"res_{tag}.csv.gz"), "rb") as f:
            ref = f.read()
        ours = S.format_csv(csv_names, len(names), inter, full, sizes, jac, 6, 0.0)
        out[f"{tag}_csv_identical"] = bool(ours == ref)
    out["ok"] = bool(out["sketches_identical"] == out["sketches"] and out["containment_csv_identical"]
                     and out["jaccard_csv_identical"])
    return out


def reference_arm(args, rank, world):
    if rank != 0:
        return
    from oracle import oracle as O
    O.build(with_ref=os.path.isdir(O.REFERENCE_SRC))
    cores = os.cpu_count() or 1
    fastas, names = make_fastas(args.genomes, args.bases, 0)
    total_bases = args.genomes * args.bases
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
    pairs = args.genomes * (args.genomes - 1) // 2
    value = total_bases / step / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, 1),
        "sketch_only_gbp_per_s": total_bases / statistics.mean(ts) / 1e9,
        "compare": {"pairs": pairs, "pairs_per_s": pairs / statistics.mean(tl), "seconds": statistics.mean(tl),
                    "note": "reference comparator is single-threaded; time = its own 'Comparisons lasted' line"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"full workload: {args.genomes} x {args.bases} bp, sub_sampler -f -t {cores} + comparator, files on tmpfs"},
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


# ------------------------------------------------------------- B200 arm

def b200_arm(args, rank, world, local_rank):
    import torch
    import supersampler_b200 as S
    from supersampler_b200 import distributed as D

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
    threads = args.threads if args.threads > 0 else max(1, min(32, cores // world))
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
    w_req = args.warmup
    args.warmup = max(args.warmup, rdepth)      # every context has run (tables, buffers) before the clock starts
    # host-buffer path: the public pipeline (pack on host threads -> pinned -> H2D -> scan -> post-pass -> compare)
    pipes = [S.Pipeline(k, m, s, device=local_rank, threads=threads) for _ in range(depth)]
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
        a pool of host threads (the post-pass is a chain of small latency-bound kernels, so batches on different
        streams overlap on the device); the compare stages run one at a time in step order on one thread (one
        NCCL exchange in flight).  A context is reused only after its previous batch has been retired."""

        def __init__(self):
            self.sk_pool = ThreadPoolExecutor(rdepth)
            self.cmp_pool = ThreadPoolExecutor(1)
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
                stats["d2h"] = sum(len(x) for x in sks) + (res[0].size * 4 if res[0] is not None else 0)
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
            out = self._note(self.stream.submit(fastas), record)
            if depth == 1:
                out = self._note(self.stream.drain(), record)
            return out

        def finish(self, record):
            return self._note(self.stream.drain(), record)

    def timed(runner, steps, warmup, ctx):
        """K steps bracketed by barrier + synchronize; CUDA events on a stream the kernels are launched on;
        the pipeline is drained inside the timed region (every step's results have reached the host)."""
        ext = torch.cuda.ExternalStream(ctx.stream(0))
        for i in range(warmup):
            runner.step(i, False)
        runner.finish(False)
        barrier(); torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record(ext)
        out = None
        for i in range(steps):
            o = runner.step(warmup + i, True)
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
        return float(t.item()), out

    # nvidia-smi needs ~0.2 s to start reporting: sample over both timed regions (device busy throughout)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    t_res, (sks_res, cmp_res) = timed(Resident(), args.steps, args.warmup, dctxs[0])
    t_e2e, (sks_e2e, cmp_e2e) = timed(HostBuffers(), args.steps, args.warmup, pctxs[0])
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
    clocks = sampler.stop() if rank == 0 else None

    # both paths must produce the same bytes / counts
    assert sks_res == sks_e2e, "device-resident and host-buffer paths disagree"
    assert np.array_equal(cmp_res[1], cmp_e2e[1])
    if rank == 0:
        assert np.array_equal(cmp_res[0], cmp_e2e[0])
    if dist is not None and use_native:
        # the exchange inside the C ABI must give what the torch.distributed exchange gives
        last = pipes[(args.warmup + args.steps - 1) % depth]
        chk = D.allgather_compare_device(last.elem_off()[0], pctxs[pipes.index(last)], rank, world, {})
        assert np.array_equal(chk[1], cmp_res[1])
        if rank == 0:
            assert np.array_equal(np.triu(chk[0], 1), np.triu(cmp_res[0], 1)), "native and torch exchange disagree"

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    n_gen_total = args.genomes * world
    total_bases = total_bases_rank * world
    pairs = n_gen_total * (n_gen_total - 1) // 2
    step_s = t_res / args.steps
    # the scan kernel's launch duration (CUDA events on its launching stream): quoted on the kernel-alone region;
    # inside the two pipelined regions it shares the SMs / the copy engines with other batches, those per-launch
    # figures are kept beside it
    scan_ms_value_region = statistics.mean(stats["scan_ms"])
    scan_ms_e2e_region = statistics.mean(x["scan_ms"] for x in stats["e2e"])
    scan_ms = statistics.mean(scan_alone)
    finfo = dctxs[0].filter_info()
    scan_kernel_name = {2: "scan_rowbit_kernel (bank-private bit table + hashed m-mer table, DESIGN.md 3.3b)",
                        1: "scan_filter_kernel (byte table of phase masks, DESIGN.md 3.3)",
                        0: "scan_filter_kernel (bit table, DESIGN.md 3.3)"}.get(finfo["kind"], "scan_dense_kernel")
    algo_bytes = n_total / 4 + 16 * stats["hits"]
    achieved = algo_bytes / (scan_ms * 1e-3) / 1e9
    sizes = cmp_res[1]
    e2 = stats["e2e"]
    mean = lambda key: statistics.mean(x[key] for x in e2)
    line = {
        "metric": METRIC, "value": total_bases / step_s / 1e9, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": w_req, "warmup_done": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": workload_config(args, world),
        "clocks": clocks,
        "e2e": {"value": total_bases / (t_e2e / args.steps) / 1e9, "unit": UNIT,
                "h2d_bytes_per_step": int(mean("h2d_bytes")), "d2h_bytes_per_step": int(mean("d2h_bytes")),
                "ms_per_step": t_e2e / args.steps * 1e3, "host_threads": threads,
                "phases_ms": {"pack_and_h2d": mean("pack_s") * 1e3, "device": mean("device_s") * 1e3,
                              "assemble": mean("assemble_s") * 1e3, "scan_kernel": mean("scan_ms"),
                              "postpass_device": mean("post_ms"), "compare_kernel": mean("cmp_kernel_ms")},
                "gpu_launches": int(sum(x["launches"] for x in e2)),
                "api": "supersampler_b200.BatchStream.submit(FASTA bytes in host memory) over two Pipelines "
                       "(.sketch() + .compare(), the compare of batch i overlapping the sketch of batch i+1)"},
        "pipelining": "off" if depth == 1 else
                      f"value: {rdepth} batches in flight (sketches on {rdepth} contexts / streams, compare stages serialised on one thread); "
                      "e2e: one batch deep (compare(i) on a background thread / second context while sketch(i+1) runs)",
        "gpu_launches": stats["launches"],
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": measured_traffic(n_total)[0], "traffic_source": measured_traffic(n_total)[1],
                     "algorithmic_bytes": int(algo_bytes), "peak_source": peak_src, "kernel": scan_kernel_name,
                     "kernel_ms": scan_ms, "bases_per_launch": int(n_total), "hits_per_launch": int(stats["hits"]),
                     "kernel_tbp_per_s": n_total / (scan_ms * 1e-3) / 1e12,
                     "kernel_ms_timed_in": f"kernel-alone region: {len(scan_alone)} launches over the rotating replicas, CUDA events on the launching stream",
                     "kernel_ms_in_value_region": scan_ms_value_region, "kernel_ms_in_e2e_region": scan_ms_e2e_region,
                     "kernel_share_of_step": scan_ms / (step_s * 1e3),
                     "ncu_pipes_pct_of_peak": measured_traffic(n_total)[2],
                     "note": "the kernel that streams every input byte; the rest of the step works on n/s-sized data "
                             "(latency-bound post-pass, INT/LSU-bound compare), see DESIGN.md section 6"},
        "phases_ms": {"sketch": statistics.mean(stats["sketch_s"]) * 1e3,
                      "compare": statistics.mean(stats["compare_s"]) * 1e3,
                      "scan_kernel": scan_ms_value_region, "postpass_device": statistics.mean(stats["post_ms"]),
                      "compare_kernel": statistics.mean(stats["cmp_ms"])},
        "d2h_bytes_per_step": int(stats["d2h"]),
        "compare": {"pairs": pairs, "pairs_per_s": pairs / statistics.mean(stats["compare_s"]),
                    "kernel_pairs_per_s": pairs / max(1e-9, statistics.mean(stats["cmp_ms"]) * 1e-3),
                    "elements": int(sizes.sum())},
    }
    # CPU baseline on this box's host cores (rank 0, N=1 only): the reference binaries on the same workload
    if world == 1 and not args.no_cpu_baseline:
        wd = scratch_dir()
        wd_keep = wd
        try:
            paths = write_files(fastas, names, wd)
            a, b, c, kind = run_reference_step(paths, args, wd, cores)
            for p_ in paths:
                os.remove(p_)
        except Exception:
            shutil.rmtree(wd, ignore_errors=True)
            raise
        # the same run is the parity check at full size: the reference's files against this step's results
        parity = None
        if kind == "reference":
            parity = reference_parity(wd_keep, names, sks_res, cmp_res, S)
        line["parity_vs_reference"] = parity
        shutil.rmtree(wd, ignore_errors=True)
        line["cpu_baseline"] = {"value": total_bases / (a + b) / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
                                "sample": f"full workload once: {args.genomes} x {args.bases} bp, sub_sampler -f -t {cores} "
                                          f"({a:.2f} s) + comparator ({b:.2f} s, single-threaded by construction)",
                                "sketch_gbp_per_s": total_bases / a / 1e9,
                                "compare_pairs_per_s": pairs / (c if c else b)}
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipelining", action="store_true", help="run sketch and compare of a step back to back")
    ap.add_argument("--depth", type=int, default=4, help="device-resident path: batches in flight")
    ap.add_argument("--threads", type=int, default=0, help="host packing threads per rank (default: cores / ranks)")
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
