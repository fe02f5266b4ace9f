#!/bin/bash
O=gpurun_out; T=${1:-r02h}
python bench.py --no-cpu --steps 5 > $O/${T}_bench_c5.json 2> $O/${T}_bench_c5.err
python bench.py --workload c3 --no-cpu --steps 10 > $O/${T}_bench_c3.json 2> $O/${T}_bench_c3.err
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=10 ) > $O/${T}_gputest.log 2>&1
echo "pytest rc=$?" >> $O/${T}_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo rc=$? >> $O/${T}_smoke.log
