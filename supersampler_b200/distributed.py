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


def tiles_for_rank(n_rows: int, n_cols: int, symmetric: bool, rank: int, ranks: int) -> List[Tuple[int, int]]:
    """The 32x32 tiles spsp_cmp_run deals to `rank` (same enumeration as
    csrc/device/capi.cu cmp_run_impl: row-major, upper triangle when symmetric,
    round-robin)."""
    n_i, n_j = (n_rows + 31) // 32, (n_cols + 31) // 32
    out, t = [], 0
    for ib in range(n_i):
        for jb in range(ib if symmetric else 0, n_j):
            if t % ranks == rank:
                out.append((ib, jb))
            t += 1
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


def exchange_elements(sizes: np.ndarray, mn: np.ndarray, lo: np.ndarray, hi: Optional[np.ndarray], device):
    """All-gather variable-length element lists.  Returns torch tensors on
    `device`: (sizes_all[world*G] int64 (cpu numpy), minim int32, klo int64, khi int64|None),
    elements of rank 0's sketches first."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    t_sizes = torch.from_numpy(sizes.copy()).to(device)
    all_sizes = [torch.empty_like(t_sizes) for _ in range(world)]
    dist.all_gather(all_sizes, t_sizes)
    all_sizes = [x.cpu().numpy() for x in all_sizes]
    e_rank = [int(x.sum()) for x in all_sizes]
    e_max = max(max(e_rank), 1)

    def gather(arr: np.ndarray, np_view, t_dtype):
        buf = torch.zeros(e_max, dtype=t_dtype, device=device)
        if arr.size:
            buf[: arr.size] = torch.from_numpy(arr.view(np_view).copy()).to(device)
        outs = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(outs, buf)
        return torch.cat([o[: e_rank[r]] for r, o in enumerate(outs)]).contiguous()

    d_mn = gather(mn, np.int32, torch.int32)
    d_lo = gather(lo, np.int64, torch.int64)
    d_hi = gather(hi, np.int64, torch.int64) if hi is not None else None
    return np.concatenate(all_sizes), d_mn, d_lo, d_hi


def allgather_compare(sketches: Sequence[bytes], k: int, m: int, rank: int, world: int, dctx, info: Optional[dict] = None):
    """All-vs-all compare of the union of all ranks' sketches (rank-major order).
    Returns (inter[N,N] uint32 -- complete on rank 0 --, sizes[N] uint64, False)."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device())
    _, _, sizes, mn, lo, hi = local_elements(sketches)
    all_sizes, d_mn, d_lo, d_hi = exchange_elements(sizes, mn, lo, hi, dev)
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
