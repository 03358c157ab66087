#!/bin/bash
# usage: tools/sweep.sh TAG "ENV1=.. ENV2=.." ...   -- one short bench run per environment set; prints the key numbers
tag=$1; shift
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/sweep_${tag}_$i.json 2> gpurun_out/sweep_${tag}_$i.log
  python - "$envs" gpurun_out/sweep_${tag}_$i.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().split('\n')[-1])
    k=d['kernel_ms_per_step']; lp=d.get('launch_profile',{})
    print("%-50s value %8.0f (%.2f ms)  e2e %8.0f  1lane %.2f inter %.2f intra %.2f dig %.2f | inter/launch %.3f intra rest %.3f" % (sys.argv[1], d['value'], d['ms_per_step'], d['e2e']['value'], k['step_on_one_lane'], k['inter'], k['intra'], k['digest'], lp.get('inter',{}).get('rest_avg_ms',0), lp.get('intra',{}).get('rest_avg_ms',0)))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
