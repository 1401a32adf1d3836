#!/bin/bash
# ncu evidence for profiles/: launch list + full captures of the dominant kernels (never a bench number).
# Each capture is summarised on the box (raw-page CSV + markdown); only the two reports read in detail afterwards
# travel back (gpurun_out is capped at 64 MiB).
mkdir -p gpurun_out
CMD="python bench.py --pairs 8192 --steps 1 --warmup 3 --no-cpu"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o gpurun_out/r01_prof_$name "$@" >> gpurun_out/ncu_full.log 2>&1
  ncu -i gpurun_out/r01_prof_$name.ncu-rep --page raw --csv > gpurun_out/r01_prof_${name}_raw.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/r01_prof_$name.ncu-rep gpurun_out/r01_${name}_kernel_ncu.md "$name kernel (round 1)" > /dev/null 2>&1
}
cap score k_score_rounds 8 2 $CMD
cap chain k_chain 20 2 $CMD
cap refit k_refit_small 12 2 $CMD
cap solve k_sample_solve 8 2 $CMD
cap sixpt k_sixpt_sample_solve 1 1 python tools/c4_probe.py 512
rm -f gpurun_out/r01_prof_chain.ncu-rep gpurun_out/r01_prof_refit.ncu-rep gpurun_out/r01_prof_solve.ncu-rep
ls -la gpurun_out | tail -20
