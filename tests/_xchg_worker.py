"""Worker of tests/test_gpu_multi.py: one rank of an N-GPU NCCL group.  Every rank sketches its own (unevenly
sized) share of a genome family on its GPU, then the ranks run spsp_cmp_exchange in all-vs-all and in query mode;
rank 0 holds the answers against the oracle run on the union of the sketches in the exchange's global order."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import supersampler_b200 as S
    from supersampler_b200 import distributed as D, synth
    from oracle import oracle as O

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    out_dir = sys.argv[1]
    k, m, s = int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # uneven split: rank r holds 5 + 3 r genomes (+ 37 on the last rank so that several tiles exist), its first
    # 1 + r are queries in query mode
    counts = [5 + 3 * r + (37 if r == world - 1 else 0) for r in range(world)]
    queries = [1 + r for r in range(world)]
    first = sum(counts[:rank])
    fam = synth.Family(60_000, seed=99)
    fastas = [fam.fasta(first + i) for i in range(counts[rank])]
    pl = S.Pipeline(k, m, s, device=local, threads=2)
    sks = pl.sketch(fastas)
    ctx = pl.device_context()
    D.join_contexts(ctx, rank, world)
    off, on_dev = pl.elem_off()
    assert on_dev
    sizes = np.diff(off.astype(np.int64)).astype(np.uint64)
    d_mn, d_lo, d_hi = ctx.batch_element_ptrs()
    n_tot = sum(counts)
    res = {}
    all_sks = [None] * world
    dist.all_gather_object(all_sks, sks)
    for mode in ("all", "query", "all_again"):
        q = None if mode != "query" else queries[rank]
        rows_cap = n_tot if q is None else sum(queries)
        info = {}
        inter, gsizes = ctx.cmp_exchange(sizes, d_mn, d_lo, d_hi if k > 32 else None, q, rows_cap, n_tot, rank, info)
        if rank == 0:
            order = D.global_order(counts, counts if q is None else queries)
            union = [None] * n_tot
            for r in range(world):
                for j, g in enumerate(order[r]):
                    union[int(g)] = all_sks[r][j]
            o_inter, o_sizes, _, _ = O.compare(union, None if q is None else sum(queries))
            ok = bool(np.array_equal(gsizes, o_sizes))
            if q is None:
                ok = ok and bool(np.array_equal(np.triu(inter, 1), np.triu(o_inter, 1)))
            else:
                Q = sum(queries)
                want = o_inter[:Q].copy().astype(np.int64)
                got = inter.astype(np.int64)
                # the oracle fills pair (i, j) at [min, max]: symmetrise its query rows
                full = np.triu(o_inter.astype(np.int64), 1)
                full = full + full.T
                want = full[:Q]
                np.fill_diagonal(got[:, :Q], 0)
                ok = ok and bool(np.array_equal(got, want))
            res[mode] = {"ok": ok, "shape": list(inter.shape), "nonzero": int((inter > 0).sum()), "kernel_ms": info["kernel_ms"]}
    # the batch form (all-vs-all of the last batches) must agree with the general one
    inter_b, sizes_b = ctx.cmp_exchange_batch(n_tot, rank)
    if rank == 0:
        res["batch_form_ok"] = bool(np.array_equal(np.triu(inter_b, 1), np.triu(inter, 1)) and np.array_equal(sizes_b, gsizes))
        with open(os.path.join(out_dir, "result.json"), "w") as f:
            json.dump(res, f)
    dist.barrier()
    pl.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
