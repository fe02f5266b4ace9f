#!/bin/bash
O=gpurun_out; T=${1:-r02i}; N=${2:-2}
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > $O/${T}_mgpu_tests.log 2>&1; echo rc=$? >> $O/${T}_mgpu_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > $O/${T}_bench_c5_g$N.json 2> $O/${T}_bench_c5_g$N.err
