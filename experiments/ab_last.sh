#!/bin/bash
# interleaved A/B of two variants on ONE box (box-to-box spread is ~2 %): bash experiments/ab_last.sh fw4 fw1
mkdir -p gpurun_out
for rep in 1 2; do
  for v in "$@"; do
    export SRB200_LIB=$PWD/experiments/ab/libsrb200_$v.so
    echo "=== $v" | tee -a gpurun_out/ab_last.log
    timeout 200 python benchmarks/bench_score.py --steps 10 2>&1 | tail -1 | tee -a gpurun_out/ab_last.log
  done
done
export SRB200_LIB=$PWD/experiments/ab/libsrb200_$2.so
timeout 300 python -m pytest tests/test_gmm_tc_gpu.py -x -q -m gpu 2>&1 | tail -2 | tee -a gpurun_out/ab_last.log
