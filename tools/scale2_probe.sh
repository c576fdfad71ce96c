#!/bin/bash
# 2-GPU probe: exchange parity tests, then the weak-scaling step with all cores and with 4 cores per rank
# (what a rank of an 8-GPU job gets on a 32-core box), for several in-flight depths.
N=2
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
run() {
  tag=$1; pre=$2; shift 2
  timeout 300 $pre python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
    bench.py --gpus $N --steps 20 --warmup 5 --extras= --no-cpu-baseline --min-seconds 0.3 "$@" > gpurun_out/s2_$tag.json 2> gpurun_out/s2_$tag.err
  echo "rc=$? $tag: $(python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s2_$tag.json").read().strip().splitlines()[-1])
    print(round(d["value"]), round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["phases_ms"].items()}, "e2e", round(d["e2e"]["value"],1), d["e2e"]["ingest"]["inputs_cleaned_on_device_per_step"])
except Exception as e:
    print("fail", e)
PY
)"
}
nproc
run all_d4c2 "" --depth 4 --cmp-depth 2
run c8_d4c2 "taskset -c 0-7" --depth 4 --cmp-depth 2 --threads 4
run c8_d8c4 "taskset -c 0-7" --depth 8 --cmp-depth 4 --threads 4
run c8_d6c3 "taskset -c 0-7" --depth 6 --cmp-depth 3 --threads 4
