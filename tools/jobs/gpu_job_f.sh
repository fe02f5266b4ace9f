#!/bin/bash
# multi-GPU job: N = $2 GPUs
O=gpurun_out; T=${1:-r02f}; N=${2:-2}
timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_cli_gpu.py -m gpu -q -x > $O/${T}_mgpu_tests.log 2>&1; echo rc=$? >> $O/${T}_mgpu_tests.log
for w in c5 c4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload $w --steps 5 --warmup 3 > $O/${T}_bench_${w}_g$N.json 2> $O/${T}_bench_${w}_g$N.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --impl reference --steps 2 --warmup 1 > $O/${T}_bench_ref_g$N.json 2> $O/${T}_bench_ref_g$N.err
python tools/cli_scale.py --workload c5 --gpus $N --repeat 2 > $O/${T}_cli_c5_g$N.json 2>> $O/${T}_cli.err
