#!/bin/bash
# A/B of the in-tree libwld.so against tools/ab/libwldB.so on the same box: pair-kernel time, with and without epilogue
O=gpurun_out; T=${1:-exp}
for rep in 1 2; do
  python tools/exp_skip_epilogue.py c3 c5 > $O/${T}_A_$rep.json 2> $O/${T}_A_$rep.err
  WLD_LIBRARY=$PWD/tools/ab/libwldB.so python tools/exp_skip_epilogue.py c3 c5 > $O/${T}_B_$rep.json 2> $O/${T}_B_$rep.err
done
WLD_LIBRARY=$PWD/tools/ab/libwldB.so timeout 300 python -m pytest tests/test_screen_refine.py tests/test_gpu_parity.py -m gpu -q -x > $O/${T}_B_tests.log 2>&1; echo rc=$? >> $O/${T}_B_tests.log
