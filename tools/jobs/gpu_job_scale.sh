#!/bin/bash
# strong scaling on one multi-GPU box: config 5 at the listed GPU counts, config 4 at the largest
O=gpurun_out; T=${1:-r02s}; shift; NS=${@:-"1 4 8"}
for N in $NS; do
  if [ "$N" = "1" ]; then
    python bench.py --no-cpu --steps 5 > $O/${T}_c5_g1.json 2> $O/${T}_c5_g1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus $N --steps 10 --warmup 3 > $O/${T}_c5_g$N.json 2> $O/${T}_c5_g$N.err
  fi
  LAST=$N
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $LAST --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus $LAST --workload c4 --steps 5 --warmup 3 > $O/${T}_c4_g$LAST.json 2> $O/${T}_c4_g$LAST.err
nvidia-smi --query-gpu=index,name,power.limit --format=csv > $O/${T}_smi.txt
