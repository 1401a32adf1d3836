#!/bin/bash
# One GPU visit: parity tests, bench (small + full C3), ncu launch list + one full capture of the hot kernel.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py --pairs 8192 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; tail -c 1500 gpurun_out/bench_small.json
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 3000 gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2>&1; tail -c 1200 gpurun_out/bench_ref.json
if [ "$1" == "ncu" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --pairs 8192 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_score_rounds -s 2 -c 2 -o gpurun_out/prof_score python bench.py --pairs 8192 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_chain -s 2 -c 2 -o gpurun_out/prof_chain python bench.py --pairs 8192 --steps 1 --warmup 3 --no-cpu >> gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
fi
