#!/bin/bash
# sweep one env knob: tools/gpu_sweep.sh NAME v1 v2 ...
mkdir -p gpurun_out
name=$1; shift
for v in "$@"; do
  env $name=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/sweep_$v.json 2> gpurun_out/sweep_$v.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/sweep_$v.json").read().strip().splitlines()[-1])
    print("$name=$v", round(d["ms_per_step"], 1), d["stage_ms_per_step"])
except Exception as e:
    print("$name=$v failed", e)
PY
done
