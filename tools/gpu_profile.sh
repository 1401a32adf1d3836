#!/bin/bash
# ncu evidence for profiles/: launch list + full captures of the dominant kernels (never a bench number).
mkdir -p gpurun_out
CMD="python bench.py --pairs 8192 --steps 1 --warmup 3 --no-cpu"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_score_rounds -s 8 -c 2 -o gpurun_out/r01_prof_score $CMD > gpurun_out/ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_chain -s 20 -c 2 -o gpurun_out/r01_prof_chain $CMD >> gpurun_out/ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_refit_small -s 12 -c 2 -o gpurun_out/r01_prof_refit $CMD >> gpurun_out/ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sample_solve -s 8 -c 2 -o gpurun_out/r01_prof_solve $CMD >> gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -8
