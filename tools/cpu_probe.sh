#!/bin/bash
# How much of the N-GPU step is host-CPU scarcity?  N=1 value region with the process confined to few cores.
run() {
  tag=$1; shift
  "$@" > gpurun_out/cpu_$tag.json 2> gpurun_out/cpu_$tag.err
  echo "rc=$? $tag: $(python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/cpu_$tag.json").read().strip().splitlines()[-1])
    print(round(d["value"]), round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["phases_ms"].items()}, "e2e", round(d["e2e"]["value"],1), d["e2e"]["host_threads"])
except Exception as e:
    print("fail", e)
PY
)"
}
nproc
B="python bench.py --steps 20 --warmup 5 --extras= --no-cpu-baseline"
run all $B
run c4_spin taskset -c 0-3 $B --threads 4
SPSP_SCHED=yield run c4_yield taskset -c 0-3 $B --threads 4
SPSP_SCHED=yield run c4_yield_d2 taskset -c 0-3 $B --threads 4 --depth 2
SPSP_SCHED=yield run all_yield $B
