#!/bin/bash
# N-GPU probe of the weak-scaling step: exchange timing breakdown + in-flight depth + host wait mode.
# usage: tools/scale_probe.sh N  (writes gpurun_out/probe_N_*.json / .err)
N=${1:-8}
run() {
  tag=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
    bench.py --gpus $N --steps 20 --warmup 5 --extras "" --no-cpu-baseline "$@" > gpurun_out/probe_${N}_$tag.json 2> gpurun_out/probe_${N}_$tag.err
  echo "rc=$? $tag: $(python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/probe_${N}_$tag.json").read().strip().splitlines()[-1])
    print(round(d["value"]), round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["phases_ms"].items()}, "e2e", round(d["e2e"]["value"],1))
except Exception as e:
    print("fail", e)
PY
)"
}
nproc
SPSP_XCHG_TIMING=1 run base
grep "\[xchg\]" gpurun_out/probe_${N}_base.err | tail -4
run cd4 --cmp-depth 4
SPSP_SCHED=yield run yield_cd2
SPSP_SCHED=yield run yield_cd4 --cmp-depth 4
run d2cd2 --depth 2 --cmp-depth 2
