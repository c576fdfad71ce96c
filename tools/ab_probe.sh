#!/bin/bash
# A/B of two builds of the device layer on one box: the `value` region of the bench (device-resident path)
for rep in 1 2; do
for v in "" _lb; do
  lib=$GRAFT_REPO_ROOT/supersampler_b200/lib/libspsp_b200$v.so
  SPSP_DEVICE_LIB=$lib python bench.py --steps 20 --warmup 5 --extras= --no-cpu-baseline > gpurun_out/ab$v.json 2>/dev/null
  echo "variant '$v': $(python - <<PY
import json
d=json.loads(open("gpurun_out/ab$v.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"],4), {k:round(x,3) for k,x in d["phases_ms"].items()}, d["regions"]["value"]["min"], d["regions"]["value"]["max"])
PY
)"
done
done
