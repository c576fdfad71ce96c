"""World_size-2 gloo run (CPU) of the compare stage's exchange step."""
import json
import os
import socket
import subprocess
import sys

from supersampler_b200 import distributed as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tile_dealing_is_a_partition():
    for n, sym in ((70, True), (100, False), (33, True), (1, True)):
        for ranks in (1, 2, 3, 8):
            seen = []
            for r in range(ranks):
                seen += D.tiles_for_rank(n, n, sym, r, ranks)
            nb = (n + 31) // 32
            want = [(i, j) for i in range(nb) for j in range(i if sym else 0, nb)]
            assert sorted(seen) == want


def test_allgather_compare_gloo_world2(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_gloo_worker.py"), str(tmp_path)],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    res = json.load(open(tmp_path / "result.json"))
    assert res["ok"] and res["n"] == 71 and res["pairs_nonzero"] > 1000
