#!/bin/bash
# Round-2 GPU job B: screen + refine — new tests, then the whole GPU suite, then bench per workload (auto vs never)
O=gpurun_out; T=${1:-r02b}
( time timeout 900 python -m pytest tests/test_screen_refine.py -m gpu -x -q --durations=8 ) > $O/${T}_screen_tests.log 2>&1
echo "pytest rc=$?" >> $O/${T}_screen_tests.log
for w in c5 c3; do
  python bench.py --workload $w --no-cpu --steps 5 > $O/${T}_bench_${w}_auto.json 2> $O/${T}_bench_${w}_auto.err
  python bench.py --workload $w --no-cpu --steps 5 --screen never > $O/${T}_bench_${w}_never.json 2> $O/${T}_bench_${w}_never.err
done
python bench.py --workload c4 --no-cpu --steps 3 > $O/${T}_bench_c4_auto.json 2> $O/${T}_bench_c4_auto.err
python bench.py --workload c3ld --no-cpu --steps 5 > $O/${T}_bench_c3ld_auto.json 2> $O/${T}_bench_c3ld_auto.err
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/${T}_gputest.log 2>&1
echo "pytest rc=$?" >> $O/${T}_gputest.log
