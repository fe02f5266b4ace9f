#!/bin/bash
# A/B of two builds of libwld.so (same ABI) on the same box: alternates them per workload.
#   tools/ab_bench.sh tools/ab/libwldA.so tools/ab/libwldB.so "c3 c5 c3ld" tag
A=$1; B=$2; W=${3:-"c3 c5"}; TAG=${4:-ab}
for w in $W; do
  for rep in 1 2; do
    for lib in $A $B; do
      WLD_LIBRARY=$PWD/$lib python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > gpurun_out/${TAG}_tmp.json 2> gpurun_out/${TAG}_tmp.err
      python - "$w" "$lib" "$rep" <<'PY'
import json, sys
d = json.loads(open("gpurun_out/%s_tmp.json" % "TAG").read().strip().splitlines()[-1]) if False else None
PY
      python -c "
import json,sys
d=json.loads(open('gpurun_out/${TAG}_tmp.json').read().strip().splitlines()[-1])
print('$w','$lib','rep$rep', 'step %.3f'%d['ms_per_step'], 'pair %.3f'%d['stages_ms']['pair'], 'alg %d'%d['roofline']['achieved'], d['clocks']['sm_mhz'])
" | tee -a gpurun_out/${TAG}.log
    done
  done
done
