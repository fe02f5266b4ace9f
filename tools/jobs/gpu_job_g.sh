#!/bin/bash
O=gpurun_out; T=${1:-r02g}
timeout 300 python -m pytest tests/test_screen_refine.py -m gpu -q -x > $O/${T}_screen_tests.log 2>&1; echo rc=$? >> $O/${T}_screen_tests.log
python bench.py --no-cpu --steps 5 > $O/${T}_bench_c5.json 2> $O/${T}_bench_c5.err
python bench.py --workload c4 --no-cpu --steps 3 > $O/${T}_bench_c4.json 2> $O/${T}_bench_c4.err
python bench.py --workload c3 --no-cpu --steps 10 > $O/${T}_bench_c3.json 2> $O/${T}_bench_c3.err
