"""Worker of tests/test_distributed_gloo.py: one rank of a world_size-2 gloo
group running the compare stage's exchange step on CPU tensors.  The device
kernel is replaced by a numpy intersection over the rank's tiles, everything
else (decode, variable-length all-gather, tile dealing, sum-reduce) is the
product's code path."""
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
    from supersampler_b200 import distributed as D
    from oracle import oracle as O
    from tests.golden_inputs import build_input
    from tests.test_host_logic import oracle_hits

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    out_dir = sys.argv[1]
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{os.environ['MASTER_PORT']}", rank=rank,
                            world_size=world)
    k, m, s = 31, 11, 100
    # 71 sketches in total (3 tiles per side), unevenly split (36 / 35): sketching shards by input file
    n_total = 71
    mine = [i for i in range(n_total) if i % world == rank]
    names = ([f"fam12_{i % 12}" for i in range(n_total)])
    sks = []
    thr = S.threshold(k, m, s)
    cache = {}
    for i in mine:
        nm = names[i]
        if nm not in cache:
            fa = build_input(nm)
            words, nb, offs = S.pack_fasta(fa, k)
            hits = oracle_hits(O, words, nb, m, thr)
            cache[nm] = S.postpass(words, offs, hits, k, m, s)[0]
        sks.append(cache[nm])
    kk, mm, sizes, mn, lo, hi = D.local_elements(sks)
    all_sizes, d_mn, d_lo, d_hi = D.exchange_elements(sizes, mn, lo, hi, torch.device("cpu"))
    # the fused single-collective exchange used on the GPUs gives the same arrays
    t = lambda a, v: torch.from_numpy(a.view(v).copy()) if a.size else None
    s2, f_mn, f_lo, f_hi = D.exchange_tensors(sizes, t(mn, np.int32), t(lo, np.int64), None, False, torch.device("cpu"))
    assert np.array_equal(s2, all_sizes) and torch.equal(f_mn, d_mn) and torch.equal(f_lo, d_lo) and f_hi is None
    n = all_sizes.size
    off = np.concatenate([[0], np.cumsum(all_sizes)])
    el = np.zeros(int(off[-1]), dtype=[("m", "<u4"), ("l", "<u8")])
    el["m"] = d_mn.numpy().view(np.uint32)
    el["l"] = d_lo.numpy().view(np.uint64)
    inter = np.zeros((n, n), np.int64)
    for ib, jb in D.tiles_for_rank(n, n, True, rank, world):
        for i in range(ib * 32, min(n, ib * 32 + 32)):
            for j in range(jb * 32, min(n, jb * 32 + 32)):
                if i < j:
                    inter[i, j] = np.intersect1d(el[off[i]:off[i + 1]], el[off[j]:off[j + 1]]).size
    t = torch.from_numpy(inter)
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        # rank-major order of the union: rank 0's sketches first
        order = [i for r in range(world) for i in range(n_total) if i % world == r]
        union = [None] * n_total
        all_sk = {}
        for i in range(n_total):
            nm = names[i]
            if nm not in all_sk:
                all_sk[nm] = O.sketch(build_input(nm), k, m, s)[0]
        union = [all_sk[names[i]] for i in order]
        o_inter, o_sizes, _, _ = O.compare(union)
        ok = bool(np.array_equal(np.triu(t.numpy(), 1), np.triu(o_inter.astype(np.int64), 1))
                  and np.array_equal(all_sizes.astype(np.uint64), o_sizes))
        with open(os.path.join(out_dir, "result.json"), "w") as f:
            json.dump({"ok": ok, "n": int(n), "elements": int(off[-1]), "pairs_nonzero": int((t.numpy() > 0).sum())}, f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
