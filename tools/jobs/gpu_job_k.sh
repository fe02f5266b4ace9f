#!/bin/bash
O=gpurun_out; T=${1:-r02k}
nproc > $O/${T}_nproc.txt
python bench.py --no-cpu --steps 5 > $O/${T}_bench_c5.json 2> $O/${T}_bench_c5.err
python tools/cli_scale.py --workload c5 --repeat 3 > $O/${T}_cli_c5.json 2>> $O/${T}_cli.err
python tools/cli_scale.py --workload c4 --repeat 2 > $O/${T}_cli_c4.json 2>> $O/${T}_cli.err
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_cli_gpu.py -m gpu -q -x > $O/${T}_tests.log 2>&1; echo rc=$? >> $O/${T}_tests.log
