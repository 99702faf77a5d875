#!/bin/bash
# A/B timing of libsrb200 variants (experiments/ab_build.py) on one GPU: scorer-only bench at C2 / real parameters / K = 256,
# and the tensor-route parity tests for every variant.  Usage: bash experiments/ab_run.sh base diet stag ...
mkdir -p gpurun_out
for v in "$@"; do
  export SRB200_LIB=$PWD/experiments/ab/libsrb200_$v.so
  echo "=== $v" | tee -a gpurun_out/ab.log
  timeout 300 python benchmarks/bench_score.py --steps 10 2>&1 | tail -1 | tee -a gpurun_out/ab.log
  timeout 300 python benchmarks/bench_score.py --steps 10 --real 2>&1 | tail -1 | tee -a gpurun_out/ab.log
  timeout 300 python benchmarks/bench_score.py --steps 5 --S 200 --K 256 --utts 8 2>&1 | tail -1 | tee -a gpurun_out/ab.log
  if [ "$v" != "base" ]; then
    timeout 600 python -m pytest tests/test_gmm_tc_gpu.py tests/test_gmm_gpu.py -x -q -m gpu 2>&1 | tail -3 | tee -a gpurun_out/ab.log
  fi
done
