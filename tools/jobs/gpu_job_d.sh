#!/bin/bash
# Round-2 GPU job D: screen + refine after tuning — bench, ncu launch list + full captures, CLI runs
O=gpurun_out; T=${1:-r02d}
python bench.py --no-cpu --steps 5 > $O/${T}_bench_c5.json 2> $O/${T}_bench_c5.err
python bench.py --workload c3 --no-cpu --steps 10 > $O/${T}_bench_c3.json 2> $O/${T}_bench_c3.err
timeout 300 python -m pytest tests/test_screen_refine.py -m gpu -q > $O/${T}_screen_tests.log 2>&1; echo rc=$? >> $O/${T}_screen_tests.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_c5.csv python bench.py --steps 2 --warmup 1 --profile --no-cpu > $O/${T}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'pair_umma|pair_refine' -c 3 -s 3 -o $O/${T}_pair_c5 -f python bench.py --steps 1 --warmup 1 --profile --no-cpu > $O/${T}_ncu_full.log 2>&1
for w in c5 c3ld c4; do python tools/cli_scale.py --workload $w --repeat 3 > $O/${T}_cli_$w.json 2>> $O/${T}_cli.err; done
