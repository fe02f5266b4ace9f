#!/bin/bash
O=gpurun_out; T=${1:-r02e}
timeout 300 python -m pytest tests/test_screen_refine.py -m gpu -q -x > $O/${T}_screen_tests.log 2>&1; echo rc=$? >> $O/${T}_screen_tests.log
python bench.py --no-cpu --steps 5 > $O/${T}_bench_c5.json 2> $O/${T}_bench_c5.err
python bench.py --workload c3 --no-cpu --steps 10 > $O/${T}_bench_c3.json 2> $O/${T}_bench_c3.err
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:'pair_refine|expand_screen' -c 4 --csv --log-file $O/${T}_refine.csv python bench.py --steps 1 --warmup 1 --profile --no-cpu > $O/${T}_ncu.log 2>&1
