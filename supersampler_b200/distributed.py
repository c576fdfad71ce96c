"""Multi-GPU plumbing (one process per GPU, torch.distributed): the compare
stage's exchange step.  Sketching shards by input file and needs no
collective; comparing needs every rank to hold every sketch, so the ranks
all-gather their sketches' elements (NCCL over NVLink on GPUs; gloo on CPU in
the tests), each rank computes the tiles dealt to it and the count matrix is
sum-reduced to rank 0 (the tiles are disjoint, so the sum is a gather).

The exchange logic is backend-agnostic (CPU tensors + gloo in tests/); only
`allgather_compare` touches the device C ABI.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np


RT = 8   # row tiles that share one column-table build (csrc/device/compare.cuh CMP_RT; 4 when k > 32)


def units_for_rank(n_rows: int, n_cols: int, symmetric: bool, rank: int, ranks: int, rt: int = RT):
    """The work units spsp_cmp_run / spsp_cmp_exchange deal to `rank`: (jb, ib0, n_ib) = column tile jb with row
    tiles ib0 .. ib0 + n_ib - 1 (same enumeration as csrc/device/capi.cu for_each_unit and
    exchange_plan_kernel: column-tile major, runs of rt row tiles, only row tiles <= jb when symmetric,
    round-robin over the ranks)."""
    n_i, n_j = (n_rows + 31) // 32, (n_cols + 31) // 32
    out, idx = [], 0
    for jb in range(n_j):
        lim = min(jb + 1, n_i) if symmetric else n_i
        for ib0 in range(0, lim, rt):
            if idx % ranks == rank:
                out.append((jb, ib0, min(rt, lim - ib0)))
            idx += 1
    return out


def tiles_for_rank(n_rows: int, n_cols: int, symmetric: bool, rank: int, ranks: int, rt: int = RT) -> List[Tuple[int, int]]:
    """The 32x32 tiles (ib, jb) inside the units dealt to `rank`."""
    return [(ib, jb) for jb, ib0, n_ib in units_for_rank(n_rows, n_cols, symmetric, rank, ranks, rt)
            for ib in range(ib0, ib0 + n_ib)]


def global_order(n_locals: Sequence[int], q_locals: Sequence[int]):
    """Where the sketches of every rank end up in the union that spsp_cmp_exchange compares: all queries
    rank-major, then all references rank-major (the reference's comparator lists the `-q` files first,
    Comparator.cpp:7-21, 512-515).  -> list over ranks of index arrays (local sketch -> global index)."""
    q_tot = int(sum(q_locals))
    q_off = np.concatenate([[0], np.cumsum(q_locals)])
    r_off = np.concatenate([[0], np.cumsum(np.asarray(n_locals) - np.asarray(q_locals))])
    out = []
    for r, (n, q) in enumerate(zip(n_locals, q_locals)):
        idx = np.empty(n, np.int64)
        idx[:q] = q_off[r] + np.arange(q)
        idx[q:] = q_tot + r_off[r] + np.arange(n - q)
        out.append(idx)
    return out


def local_elements(sketches: Sequence[bytes]):
    """Decode this rank's sketches -> (k, m, sizes[G], minim, klo, khi|None)."""
    from . import capi
    sizes, mn, lo, hi = [], [], [], []
    k = m = 0
    for sk in sketches:
        k, m, a, b, c = capi.decode_sketch(sk)
        sizes.append(a.size); mn.append(a); lo.append(b)
        if c is not None:
            hi.append(c)
    cat = lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dt)
    return k, m, np.array(sizes, np.int64), cat(mn, np.uint32), cat(lo, np.uint64), (cat(hi, np.uint64) if hi else None)


def _gather_sizes(sizes: np.ndarray, device) -> List[np.ndarray]:
    """Per-rank size lists of ranks that may hold different numbers of sketches: the counts go first, the lists
    are padded to the longest one (collectives need equal shapes on every rank)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    sizes = np.ascontiguousarray(sizes, np.int64)
    cnt = torch.tensor([sizes.size], dtype=torch.int64, device=device)
    cnts = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(cnts, cnt)
    cnts = cnts.cpu().numpy()
    n_max = max(int(cnts.max()), 1)
    pad = np.zeros(n_max, np.int64)
    pad[: sizes.size] = sizes
    out = torch.empty(world * n_max, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, torch.from_numpy(pad).to(device))
    out = out.cpu().numpy().reshape(world, n_max)
    return [out[r, : int(cnts[r])].copy() for r in range(world)]


def exchange_elements(sizes: np.ndarray, mn: np.ndarray, lo: np.ndarray, hi: Optional[np.ndarray], device,
                      has_hi: Optional[bool] = None):
    """All-gather variable-length element lists (ranks may hold different numbers of sketches, or none).
    Returns torch tensors on `device`: (sizes_all int64 (cpu numpy, rank-major), minim int32, klo int64,
    khi int64|None), elements of rank 0's sketches first.  has_hi (k > 32) must be the same on every rank: it is
    taken from `hi` only when not given."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    if has_hi is None:
        has_hi = hi is not None
    all_sizes = _gather_sizes(sizes, device)
    e_rank = [int(x.sum()) for x in all_sizes]
    e_max = max(max(e_rank), 1)

    def gather(arr: Optional[np.ndarray], np_view, t_dtype):
        buf = torch.zeros(e_max, dtype=t_dtype, device=device)
        if arr is not None and arr.size:
            buf[: arr.size] = torch.from_numpy(arr.view(np_view).copy()).to(device)
        outs = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(outs, buf)
        return torch.cat([o[: e_rank[r]] for r, o in enumerate(outs)]).contiguous()

    d_mn = gather(mn, np.int32, torch.int32)
    d_lo = gather(lo, np.int64, torch.int64)
    d_hi = gather(hi, np.int64, torch.int64) if has_hi else None
    return np.concatenate(all_sizes), d_mn, d_lo, d_hi


class _DevArray:
    """__cuda_array_interface__ view of a raw device pointer (no copy, no ownership)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def exchange_device_elements(sizes: np.ndarray, d_mn: int, d_lo: int, d_hi: Optional[int], device):
    """exchange_tensors on raw device pointers (spsp_batch_elements)."""
    import torch
    e = int(np.asarray(sizes).sum())
    wrap = lambda ptr, ts: torch.as_tensor(_DevArray(ptr, e, ts), device=device) if e else None
    return exchange_tensors(sizes, wrap(d_mn, "<i4"), wrap(d_lo, "<i8"), wrap(d_hi, "<i8") if d_hi else None,
                            d_hi is not None and d_hi != 0, device)


def exchange_tensors(sizes: np.ndarray, t_mn, t_lo, t_hi, has_hi: bool, device):
    """All-gather of the element arrays of this rank's sketches (tensors on
    `device`: int32 minimizers, int64 k-mer words; None when the rank has no
    element): sizes first, then ONE collective for the payload (klo | khi |
    minimizer packed into one padded byte buffer per rank) -- NCCL over NVLink on
    GPUs, gloo in the CPU tests.  Same return value as exchange_elements."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    all_sizes = _gather_sizes(sizes, device)
    e_rank = [int(x.sum()) for x in all_sizes]
    e_mine = int(np.asarray(sizes).sum())
    e_max = (max(max(e_rank), 1) + 1) & ~1                    # even: the 4-byte section stays 8-byte aligned
    n64 = 2 if has_hi else 1
    stride = e_max * (8 * n64 + 4)                            # bytes per rank: [klo][khi][minimizer]
    buf = torch.zeros(stride, dtype=torch.uint8, device=device)
    if e_mine:
        buf[: e_mine * 8].view(torch.int64).copy_(t_lo)
        if has_hi:
            buf[e_max * 8: e_max * 8 + e_mine * 8].view(torch.int64).copy_(t_hi)
        o = e_max * 8 * n64
        buf[o: o + e_mine * 4].view(torch.int32).copy_(t_mn)
    out = torch.empty(world * stride, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(out, buf)

    def section(byte_off, width, t_dtype):
        parts = [out[r * stride + byte_off: r * stride + byte_off + e_rank[r] * width].view(t_dtype) for r in range(world)]
        return torch.cat(parts).contiguous()

    g_lo = section(0, 8, torch.int64)
    g_hi = section(e_max * 8, 8, torch.int64) if has_hi else None
    g_mn = section(e_max * 8 * n64, 4, torch.int32)
    return np.concatenate(all_sizes), g_mn, g_lo, g_hi


def compare_gathered(all_sizes, d_mn, d_lo, d_hi, rank: int, world: int, dctx, info: Optional[dict] = None):
    """Tiles dealt to this rank on the gathered elements, then sum-reduce to rank 0."""
    import torch
    import torch.distributed as dist
    dev = d_mn.device
    n = int(all_sizes.size)
    sk_off = np.concatenate([[0], np.cumsum(all_sizes)]).astype(np.uint64)
    d_out = torch.zeros(n * n, dtype=torch.int32, device=dev)
    torch.cuda.current_stream().synchronize()
    l0 = dctx.launches()
    dctx.cmp_load_device(sk_off, d_mn.data_ptr(), d_lo.data_ptr(), d_hi.data_ptr() if d_hi is not None else None)
    dctx.cmp_run_device((0, n), (0, n), True, rank, world, d_out.data_ptr(), n)   # synchronises its stream
    if info is not None:
        info.update(kernel_ms=dctx.cmp_kernel_ms(), launches=dctx.launches() - l0)
    dist.reduce(d_out, dst=0, op=dist.ReduceOp.SUM)
    inter = d_out.cpu().numpy().view(np.uint32).reshape(n, n)
    return inter, all_sizes.astype(np.uint64), False


def allgather_compare_device(elem_off, dctx, rank: int, world: int, info: Optional[dict] = None, slot: int = 0):
    """Compare stage of a multi-GPU job whose sketches were just built by a device
    batch on every rank (elements still resident): all-gather + tiles + reduce."""
    import torch
    dev = torch.device("cuda", torch.cuda.current_device())
    sizes = np.diff(np.asarray(elem_off, np.int64))
    d_mn, d_lo, d_hi = dctx.batch_element_ptrs(slot)
    all_sizes, g_mn, g_lo, g_hi = exchange_device_elements(sizes, d_mn, d_lo, d_hi if dctx.k > 32 else None, dev)
    return compare_gathered(all_sizes, g_mn, g_lo, g_hi, rank, world, dctx, info)


def join_contexts(dctx, rank: int, world: int):
    """spsp_nccl_init on every rank: rank 0's NCCL id travels over torch.distributed."""
    import torch
    import torch.distributed as dist

    def bcast(buf):
        on_gpu = dist.get_backend() == "nccl"
        t = torch.from_numpy(buf.copy())
        if on_gpu:
            t = t.cuda()
        dist.broadcast(t, src=0)
        return t.cpu().numpy()

    dctx.nccl_init(rank, world, bcast)


def native_exchange_compare(dctx, n_local: int, rank: int, world: int, info: Optional[dict] = None, slot: int = 0):
    """The compare stage of a multi-GPU job entirely inside the C ABI (spsp_cmp_exchange_batch):
    NCCL all-gather of the device-resident elements + tiles + reduce, no tensor library in between.
    Same return value as allgather_compare_device (inter is None on ranks other than 0)."""
    inter, sizes = dctx.cmp_exchange_batch(n_local * world, rank, slot, info)
    return inter, sizes, False


def allgather_compare(sketches: Sequence[bytes], k: int, m: int, rank: int, world: int, dctx, info: Optional[dict] = None):
    """All-vs-all compare of the union of all ranks' sketches (rank-major order).
    Returns (inter[N,N] uint32 -- complete on rank 0 --, sizes[N] uint64, False)."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device())
    _, _, sizes, mn, lo, hi = local_elements(sketches)
    all_sizes, d_mn, d_lo, d_hi = exchange_elements(sizes, mn, lo, hi, dev, has_hi=k > 32)
    n = int(all_sizes.size)
    sk_off = np.concatenate([[0], np.cumsum(all_sizes)]).astype(np.uint64)
    d_out = torch.zeros(n * n, dtype=torch.int32, device=dev)
    torch.cuda.current_stream().synchronize()
    l0 = dctx.launches()
    dctx.cmp_load_device(sk_off, d_mn.data_ptr(), d_lo.data_ptr(), d_hi.data_ptr() if d_hi is not None else None)
    dctx.cmp_run_device((0, n), (0, n), True, rank, world, d_out.data_ptr(), n)   # synchronises its stream
    if info is not None:
        info.update(kernel_ms=dctx.cmp_kernel_ms(), launches=dctx.launches() - l0)
    dist.reduce(d_out, dst=0, op=dist.ReduceOp.SUM)
    inter = d_out.cpu().numpy().view(np.uint32).reshape(n, n)
    return inter, all_sizes.astype(np.uint64), False


def _parse_cpulist(txt: str) -> List[int]:
    out: List[int] = []
    for part in txt.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out


def bind_rank_to_gpu_cores(local_rank: int, local_world: int, pci_bus_ids: Optional[Sequence[str]] = None) -> dict:
    """One process per GPU on a multi-socket box: confine this rank (and every thread it starts: pack workers,
    pinned allocations by first touch) to its share of the cores that are LOCAL to its GPU (sysfs `local_cpulist`
    of the GPU's PCI device = the NUMA node behind its root complex), split evenly among the ranks whose GPUs sit
    on the same node.  Without it the ranks' packers read text and write pinned memory across the socket link.
    Returns what it did (for the bench line); a no-op when the topology cannot be read."""
    import os
    info = {"bound": False}
    try:
        if pci_bus_ids is None:
            import subprocess
            q = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader"], stdout=subprocess.PIPE,
                               stderr=subprocess.DEVNULL, text=True, timeout=20).stdout.split()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                q = [q[int(x)] for x in vis.split(",") if x.strip().isdigit() and int(x) < len(q)]
            pci_bus_ids = q
        if len(pci_bus_ids) < local_world:
            return info
        allowed = sorted(os.sched_getaffinity(0))

        def local_cpus(bus_id: str) -> List[int]:
            b = bus_id.lower()
            if b.count(":") == 2 and len(b.split(":")[0]) == 8:          # nvidia-smi prints an 8-digit domain
                b = b[4:]
            with open(f"/sys/bus/pci/devices/{b}/local_cpulist") as f:
                cpus = [c for c in _parse_cpulist(f.read()) if c in allowed]
            return cpus or allowed

        sets = [tuple(local_cpus(pci_bus_ids[r])) for r in range(local_world)]
        mine = sets[local_rank]
        peers = [r for r in range(local_world) if sets[r] == mine]            # ranks that share my node
        share = len(mine) // len(peers)
        if share < 1:
            return info
        j = peers.index(local_rank)
        cpus = list(mine[j * share:(j + 1) * share])
        os.sched_setaffinity(0, cpus)
        info.update(bound=True, cpus=f"{cpus[0]}-{cpus[-1]}" if cpus == list(range(cpus[0], cpus[-1] + 1)) else cpus,
                    n_cpus=len(cpus), ranks_on_node=len(peers), node_cpus=len(mine))
    except Exception as ex:                      # topology not readable: stay unbound
        info["error"] = f"{type(ex).__name__}: {ex}"
    return info
