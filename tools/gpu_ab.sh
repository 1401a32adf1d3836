#!/bin/bash
# A/B on the GPU box: parity tests, then the default bench with and without an env knob ($1, e.g. SSFM_NO_PREFILTER).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
env $1=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
python - <<'PY'
import json
for f in ("a", "b"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["ms_per_step"], 1), {k: v for k, v in d.items() if k in ("stages", "stage_ms", "run_stats")})
    except Exception as e:
        print(f, "failed", e, open(f"gpurun_out/bench_{f}.err").read()[-800:])
PY
