#!/bin/bash
# e2e with the three ingest modes, all cores and 4 cores (what a rank of an 8-GPU job gets on a 32-core box)
run() {
  tag=$1; shift
  "$@" > gpurun_out/ing_$tag.json 2> gpurun_out/ing_$tag.err
  echo "rc=$? $tag: $(python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ing_$tag.json").read().strip().splitlines()[-1])
    e=d["e2e"]
    print("value", round(d["value"]), "e2e", round(e["value"],1), "ms", round(e["ms_per_step"],2), {k:round(v,3) for k,v in e["phases_ms"].items()}, e["ingest"]["inputs_cleaned_on_device_per_step"], round(e["ingest"]["ingest_kernels_ms"],3), "thr", e["host_threads"], d["regions"]["e2e"])
except Exception as ex:
    print("fail", ex)
PY
)"
}
nproc
B="python bench.py --steps 20 --warmup 5 --extras= --no-cpu-baseline"
run host $B --ingest host
run auto $B --ingest auto
run device $B --ingest device
run c4_host taskset -c 0-3 $B --threads 4 --ingest host
run c4_auto taskset -c 0-3 $B --threads 4 --ingest auto
run c4_device taskset -c 0-3 $B --threads 4 --ingest device
