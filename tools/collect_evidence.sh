#!/bin/bash
# Runs on the GPU box (gpurun): bench (both arms), ncu launch list of the bench command, full captures of the
# dominant kernels, dense-totals timing.  Everything lands in gpurun_out/ev_*; tools/write_profiles.py turns
# it into profiles/.  Numbers printed under ncu are never used as bench values.
set -u
O=gpurun_out
python bench.py --impl reference --steps 2 --warmup 1 > $O/ev_bench_ref.json 2> $O/ev_bench_ref.err
python bench.py --steps 5 --warmup 3 > $O/ev_bench_n1.json 2> $O/ev_bench_n1.err || { tail -5 $O/ev_bench_n1.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/ev_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ev_ncu_launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"scan_filter|scan_rowbit" -c 2 -o $O/ev_prof_scan \
    python tools/profile_batch.py 64 1000 2 > $O/ev_ncu_scan.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"pp_chain|hashjoin|pp_emit" -c 4 -o $O/ev_prof_post \
    python tools/profile_batch.py 64 1000 1 > $O/ev_ncu_post.log 2>&1
ncu --set full --clock-control none -k regex:"dense_rows|dense_segments" -c 2 -o $O/ev_prof_dense \
    python tools/profile_batch.py 64 1000 1 > $O/ev_ncu_dense.log 2>&1
python tools/profile_batch.py 64 1000 3 > $O/ev_batch_s1000.txt 2>&1
python tools/profile_batch.py 64 100 3 > $O/ev_batch_s100.txt 2>&1
python tools/pipeline_probe.py 64 4,8,16 > $O/ev_pipeline_probe.txt 2>&1
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem --format=csv > $O/ev_gpu.txt
lscpu | grep -E "Model name|^CPU\(s\)|Thread|L3" >> $O/ev_gpu.txt
echo done
