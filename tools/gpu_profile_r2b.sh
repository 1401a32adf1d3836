#!/bin/bash
# Round-2 (session 3) evidence: the fixed pre-filter test, refreshed launch list + --set full captures of the 3-point
# kernels (tools/gpu_profile_r2.sh), and --set full captures of the other paths' dominant kernels through their probes.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "near_noise_free or c3_slice or c1_against" > gpurun_out/r2s3_gputest2.log 2>&1; tail -3 gpurun_out/r2s3_gputest2.log
bash tools/gpu_profile_r2.sh
capp() {  # name, kernel regex, skip, title, command...
  local name=$1 rx=$2 skip=$3 title=$4; shift 4
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o gpurun_out/r02_prof_$name "$@" >> gpurun_out/ncu_full_b.log 2>&1
  python tools/ncu_summary.py gpurun_out/r02_prof_$name.ncu-rep gpurun_out/r02_${name}_kernel_ncu.md "$title" > /dev/null 2>&1
}
rm -f gpurun_out/r02_prof_match*.ncu-rep
capp match k_match_2nn 1 "descriptor-matching kernel k_match_2nn (round 2: 2016 pairs x 4000 x 4000)" python tools/match_probe.py 64 4000 2016
capp sixpt k_sixpt_sample_solve 2 "six-point solver kernel (round 2: 2000 pairs x 1000 corr, config C4)" python tools/c4_probe.py 2000
capp tri k_retriangulate 0 "Retriangulate kernel (round 2: 200 000 points)" python tools/tri_probe.py 200000
rm -f gpurun_out/r02_prof_tri.ncu-rep
ls -la gpurun_out | grep r02_
