#!/bin/bash
# post-pass A/B: bucket-kernel scale variants (SPSP_DEVICE_LIB) on the C2 batch and the s=100 1 Gbp batch
for v in "" _s1 _s4; do
  lib=$GRAFT_REPO_ROOT/supersampler_b200/lib/libspsp_b200$v.so
  [ -f $lib ] || continue
  for cfg in "64 1000" "200 100"; do
    echo "variant '$v' cfg $cfg: $(SPSP_DEVICE_LIB=$lib python tools/profile_batch.py $cfg 4 2>&1 | grep post_ms)"
  done
done
SPSP_PP_DEBUG=1 python tools/profile_batch.py 200 100 2 2>&1 | grep -E "^\[pp\]" | tail -1
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "postpass or pipeline or random or golden or sketch" 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/pp4_s1000.csv python tools/profile_batch.py 64 1000 2 > /dev/null 2>&1
python tools/launch_list.py gpurun_out/pp4_s1000.csv
