#!/bin/bash
# Round-2 final evidence: GPU tests, default bench (both arms), launch list + --set full capture of the scoring kernel.
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -x -q > gpurun_out/r2s3_gputest10.log 2>&1; tail -3 gpurun_out/r2s3_gputest10.log
timeout 500 python bench.py > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; tail -c 200 gpurun_out/r02_bench_c3.json; tail -2 gpurun_out/r02_bench_c3.err
timeout 300 python bench.py --impl reference > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; tail -c 200 gpurun_out/r02_bench_ref.json
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-extras"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 1200 --csv --log-file gpurun_out/r02_launches_c3_full.csv $CMD > gpurun_out/ncu_launch.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_score_rounds -s 5 -c 2 -o gpurun_out/r02_prof_score $CMD > gpurun_out/ncu_full.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_prof_score.ncu-rep gpurun_out/r02_score_kernel_ncu.md "score kernel (round 2, full C3: 124 750 pairs)" > /dev/null 2>&1
ncu -i gpurun_out/r02_prof_score.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys,json
rows=list(csv.reader(sys.stdin)); h=rows[0]; r=rows[2]
g=lambda k: float(r[h.index(k)].replace(',',''))
print(json.dumps({'dram_bytes_read': g('dram__bytes_read.sum'), 'dram_bytes_write': g('dram__bytes_write.sum'), 'units': rows[1][h.index('dram__bytes_read.sum')], 'grid': r[h.index('launch__grid_size')], 'ms': g('gpu__time_duration.sum')}))" > gpurun_out/score_kernel_traffic_new.json 2>/dev/null
ls -la gpurun_out | grep "r02_\|traffic"
