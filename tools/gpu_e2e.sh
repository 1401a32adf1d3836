#!/bin/bash
# e2e sweep: env settings given as "A=1 B=2" strings
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/e2e_$i.json 2> gpurun_out/e2e_$i.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/e2e_$i.json").read().strip().splitlines()[-1])
    print("$cfg", "value ms", round(d["ms_per_step"], 1), "e2e ms", round(d["e2e"]["ms_per_step"], 1))
except Exception as e:
    print("$cfg failed", e, open("gpurun_out/e2e_$i.err").read()[-500:])
PY
done
