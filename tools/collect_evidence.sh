#!/bin/bash
# Runs on the GPU box (gpurun): bench (both arms), ncu launch lists, full captures of the dominant kernels,
# per-batch timings.  Everything lands in gpurun_out/ev_*; tools/write_profiles.py turns it into profiles/.
# Numbers printed under ncu are never used as bench values; every ncu command runs after the same command exited 0.
set -u
O=gpurun_out
NCU="ncu --set full --import-source on --clock-control none"
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/ev_bench_ref.json 2> $O/ev_bench_ref.err
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/ev_bench_n1.json 2> $O/ev_bench_n1.err || { tail -5 $O/ev_bench_n1.err; exit 1; }
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/ev_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/ev_smoke.txt 2>&1
for cfg in "64 1000 3 11" "64 100 3 11" "200 100 3 11" "64 200 3 13"; do
  set -- $cfg
  python tools/profile_batch.py $cfg > $O/ev_batch_n$1_s$2_m$4.txt 2>&1
done
python tools/profile_ingest.py 64 3 > $O/ev_ingest.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/ev_batch_launches.csv \
    python tools/profile_batch.py 64 1000 2 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/ev_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --extras= --min-seconds 0.01 > $O/ev_ncu_launches.log 2>&1
$NCU -k regex:"scan_filter|scan_rowbit" -c 2 -f -o $O/ev_prof_scan python tools/profile_batch.py 64 1000 2 > $O/ev_ncu_scan.log 2>&1
$NCU -k regex:"scan_filter|scan_rowbit" -s 1 -c 1 -f -o $O/ev_prof_scan_s100 python tools/profile_batch.py 64 100 2 > /dev/null 2>&1
$NCU -k regex:"scan_filter|scan_rowbit" -s 1 -c 1 -f -o $O/ev_prof_scan_m13 python tools/profile_batch.py 64 200 2 13 > /dev/null 2>&1
$NCU -k regex:"pp_bucket|pp_sort_small|pp_replay|hashjoin" -s 6 -c 6 -f -o $O/ev_prof_post python tools/profile_batch.py 64 1000 2 > $O/ev_ncu_post.log 2>&1
$NCU -k regex:"ingest_" -s 5 -c 5 -f -o $O/ev_prof_ingest python tools/profile_ingest.py 64 2 > /dev/null 2>&1
$NCU -k regex:"hashjoin" -s 2 -c 1 -f -o $O/ev_prof_cmp_c3 python tools/compare_probe.py 256 100 0 > /dev/null 2>&1
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem --format=csv > $O/ev_gpu.txt
lscpu | grep -E "Model name|^CPU\(s\)|Thread|L3" >> $O/ev_gpu.txt
echo done
