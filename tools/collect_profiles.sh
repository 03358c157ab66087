#!/bin/bash
# Run on the GPU box (gpurun): bench lines, the launch list of one timed step and ncu --set full captures of every kernel
# family, into gpurun_out/.  usage: tools/collect_profiles.sh TAG
tag=$1
set -x
timeout 300 python bench.py > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${tag}_ref.json 2>> gpurun_out/bench_${tag}_n1.log
timeout 300 python bench.py --mode spec --steps 3 --no-cpu-baseline --no-extras > gpurun_out/bench_${tag}_spec.json 2>> gpurun_out/bench_${tag}_n1.log
# launch list of ONE timed resident step (bench.py brackets the timed steps with cudaProfilerStart/Stop)
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_launches_${tag}.log 2>&1
# full sections: the first launches of every family of that step (wave 0 = I pictures, wave 1.. = P pictures)
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_(inter|intra|residual|digest|zero)' -c 12 \
    -o gpurun_out/prof_step_${tag} -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_step_${tag}.log 2>&1
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:'k_(blk_scan|stream_expand)' -c 4 \
    -o gpurun_out/prof_scan_${tag} -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_scan_${tag}.log 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:'k_(deblock|pad|inter|intra)' -s 6 -c 8 \
    -o gpurun_out/prof_spec_${tag} -f python bench.py --mode spec --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_spec_${tag}.log 2>&1
ls -la gpurun_out/*${tag}*
