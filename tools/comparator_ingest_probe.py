#!/usr/bin/env python
"""bin/comparator on thousands of gzip sketch files: N sketches of 1 Mbp genomes at k31 m13 s200 are
written as .gz files, `comparator -q` reads + decodes them on all host threads (while the driver starts on another
thread), compares Q x N on the GPU and streams both CSVs; SPSP_TRACE shows the phases.  The reference comparator
runs on a 256-file subset (it is single-threaded and keeps every file open) and must give the same CSV bytes."""
import gzip, os, subprocess, sys, tempfile, time
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import supersampler_b200 as S
from supersampler_b200 import capi, synth_device as SD
from oracle import oracle as O


def main():
    S.build(); O.build(with_ref=False)
    n, q = (int(sys.argv[1]) if len(sys.argv) > 1 else 4000), 100
    k, m, s, nb = 31, 13, 200.0, 1_000_000
    wd = tempfile.mkdtemp(prefix="spsp_cmpin_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    fam = SD.DeviceFamily(nb, seed=99)
    ctx = S.DeviceContext(k, m, S.threshold(k, m, s))
    names = []
    t0 = time.perf_counter()
    with ThreadPoolExecutor(os.cpu_count()) as ex:
        for a in range(0, n, 500):
            cnt = min(500, n - a)
            buf, n_total, rb, re_, ri = fam.packed_batch(a, cnt)
            sks = list(ctx.sketch_batch(None, n_total, rb, re_, ri, cnt, s, device_ptr=buf.data_ptr()))
            paths = [os.path.join(wd, f"sk{a + i:05d}.gz") for i in range(cnt)]
            list(ex.map(lambda pr: open(pr[0], "wb").write(gzip.compress(pr[1], 1)), zip(paths, sks)))
            names += paths
    ctx.close()
    print(f"{n} sketches written in {time.perf_counter() - t0:.1f} s, {sum(os.path.getsize(p) for p in names) / 1e6:.0f} MB of .gz")
    open(os.path.join(wd, "q.txt"), "w").write("\n".join(names[:q]) + "\n")
    open(os.path.join(wd, "r.txt"), "w").write("\n".join(names[q:]) + "\n")
    env = dict(os.environ, SPSP_TRACE="1")
    t0 = time.perf_counter()
    r = subprocess.run([os.path.join(capi.BIN_DIR, "comparator"), "-f", "r.txt", "-q", "q.txt", "-o", "ours"], cwd=wd, env=env,
                       stdin=subprocess.DEVNULL, capture_output=True, text=True)
    t_ours = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr[-2000:]
    print(r.stderr.strip())
    print(f"ours: comparator -q {q} x {n - q} files: {t_ours:.2f} s; containment CSV {os.path.getsize(os.path.join(wd, 'ours_containment.csv.gz')) / 1e6:.1f} MB gz")
    if O.have_ref():
        sub = 256
        open(os.path.join(wd, "r2.txt"), "w").write("\n".join(names[q:sub]) + "\n")
        t0 = time.perf_counter()
        subprocess.run([os.path.join(O.REF_DIR, "comparator"), "-f", "r2.txt", "-q", "q.txt", "-o", "ref"], cwd=wd,
                       stdin=subprocess.DEVNULL, stdout=subprocess.DEVNULL, check=True)
        t_ref = time.perf_counter() - t0
        subprocess.run([os.path.join(capi.BIN_DIR, "comparator"), "-f", "r2.txt", "-q", "q.txt", "-o", "ours2"], cwd=wd,
                       stdin=subprocess.DEVNULL, stdout=subprocess.DEVNULL, check=True)
        same = all(gzip.open(os.path.join(wd, f"ours2_{t}.csv.gz")).read() == gzip.open(os.path.join(wd, f"ref_{t}.csv.gz")).read()
                   for t in ("containment", "jaccard"))
        print(f"reference comparator on {q} x {sub - q} files: {t_ref:.2f} s; CSVs identical with ours on the same files: {same}")
    subprocess.run(["rm", "-rf", wd])


if __name__ == "__main__":
    main()
