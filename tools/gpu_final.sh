#!/bin/bash
# Round-end evidence: default bench (both arms), then the ncu passes.
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 600 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
bash tools/gpu_profile.sh
