#!/bin/bash
# Round-2 ncu evidence on the full C3 workload (124 750 pairs): launch list of the default bench command's kernels, and
# --set full captures of the BIG launch of every stage (ordinals taken from the launch list: step 2 of the warm-up).
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-extras"
if [ "$1" != "nolist" ]; then
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 1200 --csv --log-file gpurun_out/r02_launches_c3_full.csv $CMD > gpurun_out/ncu_launch.log 2>&1
fi
cap() {  # name, kernel regex, skip, count
  local name=$1 rx=$2 skip=$3 cnt=$4
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o gpurun_out/r02_prof_$name $CMD >> gpurun_out/ncu_full.log 2>&1
  python tools/ncu_summary.py gpurun_out/r02_prof_$name.ncu-rep gpurun_out/r02_${name}_kernel_ncu.md "$name kernel (round 2, full C3: 124 750 pairs)" > /dev/null 2>&1
}
cap score k_score_rounds 5 2
cap solve k_sample_solve 5 2
cap chain k_chain 32 2
cap refit_small k_refit_small 24 1
cap refit_big k_refit_big 26 1
rm -f gpurun_out/r02_prof_solve.ncu-rep gpurun_out/r02_prof_chain.ncu-rep gpurun_out/r02_prof_refit_small.ncu-rep gpurun_out/r02_prof_refit_big.ncu-rep
ls -la gpurun_out | tail -12
