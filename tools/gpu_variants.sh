#!/bin/bash
# A/B of library variants and env knobs on the GPU box: tools/gpu_variants.sh "label:ENV=val,ENV2=val:libpath" ...
mkdir -p gpurun_out
for spec in "$@"; do
  label=${spec%%:*}; rest=${spec#*:}; envs=${rest%%:*}; lib=${rest#*:}
  (
    if [ "$lib" != "default" ] && [ -n "$lib" ]; then export SSFM_LIB_PATH=$PWD/$lib; fi
    IFS=',' read -ra kv <<< "$envs"; for e in "${kv[@]}"; do [ -n "$e" ] && [ "$e" != "-" ] && export "$e"; done
    timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/var_$label.json 2> gpurun_out/var_$label.err
  )
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/var_$label.json").read().strip().splitlines()[-1])
    print("$label", "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["ms_per_step"], 1), {k: round(v, 1) for k, v in d["stage_ms_per_step"].items()}, "roofline", round(d["roofline"]["frac"], 3))
except Exception as e:
    print("$label failed", e, open("gpurun_out/var_$label.err").read()[-400:])
PY
done
