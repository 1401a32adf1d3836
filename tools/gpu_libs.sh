#!/bin/bash
# A/B of prebuilt library variants: tools/gpu_libs.sh path1 path2 ...  ("default" = the in-tree library)
mkdir -p gpurun_out
for lib in "$@"; do
  if [ "$lib" == "default" ]; then unset SSFM_LIB_PATH; else export SSFM_LIB_PATH=$PWD/$lib; fi
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/lib_run.json 2> gpurun_out/lib_run.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/lib_run.json").read().strip().splitlines()[-1])
    print("$lib", round(d["ms_per_step"], 1), d["stage_ms_per_step"])
except Exception as e:
    print("$lib failed", e, open("gpurun_out/lib_run.err").read()[-400:])
PY
done
