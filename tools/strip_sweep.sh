#!/bin/bash
# Sweeps the rasterisation strip width of the pair-stage tile schedule (WLD_STRIP) on one workload.
W=${1:-c5}
for s in 2 4 8 16 32; do
  WLD_STRIP=$s timeout 600 python bench.py --workload $W --steps 4 --warmup 3 --no-cpu > gpurun_out/strip_${W}_$s.json 2> gpurun_out/strip_${W}_$s.err
  python - <<PY
import json
d=json.load(open("gpurun_out/strip_${W}_$s.json"))
print("strip $s", "$W", "pair_ms %.2f"%d["stages_ms"]["pair"], "value %.3e"%d["value"], d["clocks"]["sm_mhz"])
PY
done
