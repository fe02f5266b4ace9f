#!/bin/bash
# Round-2 GPU job A: whole -m gpu suite (incl. full-size configs), default bench + reference arm, launch list, full capture
O=gpurun_out; T=r02a
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 ) > $O/${T}_gputest.log 2>&1
echo "pytest rc=$?" >> $O/${T}_gputest.log
python bench.py > $O/${T}_bench_c5.json 2> $O/${T}_bench_c5.err
python bench.py --impl reference > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err
python bench.py --workload c4 --no-cpu > $O/${T}_bench_c4.json 2> $O/${T}_bench_c4.err
python bench.py --workload c3 --steps 10 --no-cpu > $O/${T}_bench_c3.json 2> $O/${T}_bench_c3.err
python bench.py --workload c3ld --steps 10 --no-cpu > $O/${T}_bench_c3ld.json 2> $O/${T}_bench_c3ld.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_c5.csv python bench.py --steps 2 --warmup 1 --profile --no-cpu > $O/${T}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pair_umma -c 1 -s 2 -o $O/${T}_pair_umma_c5 -f python bench.py --steps 1 --warmup 2 --profile --no-cpu > $O/${T}_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'hist_vec16|gather_kernel|accumulate|expand_|quantize' -c 12 -o $O/${T}_hbm_c5 -f python bench.py --steps 1 --warmup 0 --profile --no-cpu > $O/${T}_ncu_hbm.log 2>&1
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm --format=csv > $O/${T}_smi.txt
