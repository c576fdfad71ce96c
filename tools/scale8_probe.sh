#!/bin/bash
N=${1:-8}
run() {
  tag=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
    bench.py --gpus $N --steps 20 --warmup 5 --extras= --no-cpu-baseline --min-seconds 0.3 "$@" > gpurun_out/s8_$tag.json 2> gpurun_out/s8_$tag.err
  echo "rc=$? $tag: $(python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s8_$tag.json").read().strip().splitlines()[-1])
    e=d["e2e"]
    print(round(d["value"]), round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["phases_ms"].items()}, "| e2e", round(e["value"],1), round(e["ms_per_step"],2), {k:round(v,2) for k,v in e["phases_ms"].items()}, e["ingest"]["inputs_cleaned_on_device_per_step"], e["host_threads"], e["numa_binding"])
except Exception as ex:
    print("fail", ex)
PY
)"
}
nvidia-smi topo -m 2>/dev/null | head -14
run bind
run bind_d8 --depth 8 --cmp-depth 4
