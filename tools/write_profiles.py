#!/usr/bin/env python
"""gpurun_out/ev_* (tools/collect_evidence.sh) -> profiles/ (tracked, judged)."""
import io, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
R = sys.argv[1] if len(sys.argv) > 1 else "r01"

def run(*a):
    return subprocess.run(list(a), capture_output=True, text=True).stdout

def main():
    os.makedirs(P, exist_ok=True)
    for src, dst in (("ev_bench_n1.json", f"{R}_bench_n1.json"), ("ev_bench_ref.json", f"{R}_bench_reference_arm.json"),
                     ("ev_launches.csv", f"{R}_bench_launches.csv"), ("ev_pipeline_probe.txt", f"{R}_pipeline_probe.txt"),
                     ("ev_gpu.txt", f"{R}_box.txt"), ("ev_tests.txt", f"{R}_gpu_tests.txt"), ("ev_smoke.txt", f"{R}_smoke.txt"),
                     ("ev_ingest.txt", f"{R}_ingest.txt")):
        if os.path.exists(os.path.join(G, src)):
            shutil.copy(os.path.join(G, src), os.path.join(P, dst))
    with open(os.path.join(P, f"{R}_bench_launch_summary.md"), "w") as f:
        f.write("# ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (per-launch times are cold and serialised: compare shares)\n\n")
        f.write(run(sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), os.path.join(G, "ev_launches.csv")))
    if os.path.exists(os.path.join(G, "ev_batch_launches.csv")):
        with open(os.path.join(P, f"{R}_batch_launches.md"), "w") as f:
            f.write("# ncu launch list of the last batch of `python tools/profile_batch.py 64 1000 2` (scan + post-pass + compare of one "
                    "64 x 5 Mbp batch; per-launch times are cold and serialised)\n\n```\n")
            f.write(run(sys.executable, os.path.join(ROOT, "tools", "launch_list.py"), os.path.join(G, "ev_batch_launches.csv")))
            f.write("```\n")
    for rep, name, title in (("ev_prof_scan", "scan_ncu", "scan kernel of the bench workload (scan_rowbit_kernel at k31 m11 s1000)"),
                             ("ev_prof_scan_s100", "scan_s100_ncu", "scan_filter_kernel at k31 m11 s100 (C3's table: denser threshold), 320 Mbp"),
                             ("ev_prof_scan_m13", "scan_m13_ncu", "scan_filter_kernel at k31 m13 s200 (C5's table), 320 Mbp"),
                             ("ev_prof_post", "postpass_compare_ncu", "pp_sort_small / pp_replay / pp_bucket / hashjoin kernels of one 64 x 5 Mbp batch"),
                             ("ev_prof_ingest", "ingest_ncu", "ingest kernels (summary, carry, totals, zero, write) on 64 x 5 Mbp of FASTA text"),
                             ("ev_prof_cmp_c3", "compare_c3_ncu", "hashjoin_kernel, all-vs-all of 256 sketches of 52 k elements (C3 shape)"),
                             ("ev_prof_dense", "dense_ncu", "dense_rows_kernel / dense_segments_kernel, bench workload")):
        path = os.path.join(G, rep + ".ncu-rep")
        if os.path.exists(path):
            with open(os.path.join(P, f"{R}_{name}.md"), "w") as f:
                f.write(f"# ncu --set full --clock-control none: {title}\n\n")
                f.write(run(sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), path))
    # traffic of the dominant streaming kernel for bench.py's roofline.traffic
    path = os.path.join(G, "ev_prof_scan.ncu-rep")
    if os.path.exists(path):
        import csv
        rows = list(csv.reader(io.StringIO(run("ncu", "-i", path, "--page", "raw", "--csv"))))
        hdr, units = rows[0], rows[1]
        def val(name, row):
            i = hdr.index(name); v = float(row[i]); u = units[i].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        rd, wr = val("dram__bytes_read.sum", rows[2]), val("dram__bytes_write.sum", rows[2])
        with open(os.path.join(P, f"{R}_traffic.json"), "w") as f:
            def pct(name):
                return round(float(rows[2][hdr.index(name)]), 1) if name in hdr else None
            json.dump({"kernel": rows[2][hdr.index("Kernel Name")][:60], "workload": "C2 batch: 320012288 bases, k31 m11 s1000",
                       "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "traffic_bytes_per_launch": int(rd + wr),
                       "pipes_pct_of_peak": {
                           "lsu_wavefronts": pct("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                           "issue_active": pct("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                           "alu": pct("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
                           "fma": pct("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                           "dram": pct("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")},
                       "source": f"ncu --set full --clock-control none, profiles/{R}_scan_ncu.md launch 0"}, f, indent=1)
    import glob
    for t in glob.glob(os.path.join(G, "ev_batch_n*.txt")):
        shutil.copy(t, os.path.join(P, f"{R}_{os.path.basename(t)[3:]}"))
    # compare kernel pipes for bench.py's compare.roofline
    path = os.path.join(G, "ev_prof_cmp_c3.ncu-rep")
    if os.path.exists(path):
        import csv
        rows = list(csv.reader(io.StringIO(run("ncu", "-i", path, "--page", "raw", "--csv"))))
        hdr = rows[0]
        def pc(name):
            return round(float(rows[2][hdr.index(name)]), 1) if name in hdr else None
        with open(os.path.join(P, f"{R}_compare_pipes.json"), "w") as f:
            json.dump({"ncu_workload": "hashjoin_kernel, 256 sketches of 52 k elements (C3 shape), one GPU",
                       "issue_active_pct": pc("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                       "alu_pct": pc("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
                       "lsu_pct": pc("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                       "dram_pct": pc("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                       "source": f"ncu --set full --clock-control none, profiles/{R}_compare_c3_ncu.md"}, f, indent=1)
    print("profiles/ updated")

if __name__ == "__main__":
    main()
