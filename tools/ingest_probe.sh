#!/bin/bash
# e2e with the ingest modes, all cores and 4 cores (what a rank of an 8-GPU job gets on a 32-core box)
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
for mode in ${MODES:-host auto}; do
  run $mode $B --ingest $mode
  run c4_$mode taskset -c 0-3 $B --threads 4 --ingest $mode
done
