#!/bin/bash
# final single-GPU evidence of the round: default bench line, reference arm, launch list, full capture of the dominant kernels
O=gpurun_out; T=${1:-r02z}
python bench.py > $O/${T}_bench_c5.json 2> $O/${T}_bench_c5.err
python bench.py --impl reference > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_c5.csv python bench.py --steps 2 --warmup 1 --profile --no-cpu > $O/${T}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'pair_umma|pair_refine|expand_screen' -c 4 -s 4 -o $O/${T}_pair_c5 -f python bench.py --steps 1 --warmup 1 --profile --no-cpu > $O/${T}_ncu_full.log 2>&1
python bench.py --workload c4 --no-cpu --steps 3 > $O/${T}_bench_c4.json 2> $O/${T}_bench_c4.err
python bench.py --workload c3 --no-cpu --steps 10 > $O/${T}_bench_c3.json 2> $O/${T}_bench_c3.err
python bench.py --workload c3ld --no-cpu --steps 10 > $O/${T}_bench_c3ld.json 2> $O/${T}_bench_c3ld.err
