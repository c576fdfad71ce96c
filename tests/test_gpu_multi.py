"""Multi-GPU exchange on hardware (needs >= 2 GPUs on the box; skipped otherwise): spsp_cmp_exchange in
all-vs-all and query mode over NCCL against the oracle, uneven shares per rank."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("k,m,s", [(31, 11, 20.0), (63, 15, 10.0)])
def test_exchange_all_and_query_mode(tmp_path, k, m, s):
    n = min(_gpus(), 4)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_xchg_worker.py"), str(tmp_path), str(k), str(m), str(s)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:]
    res = json.load(open(tmp_path / "result.json"))
    for mode in ("all", "query", "all_again"):
        assert res[mode]["ok"], res
        assert res[mode]["nonzero"] > 0
    assert res["batch_form_ok"], res
