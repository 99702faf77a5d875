#!/bin/bash
# timing-only A/B (no parity tests): bash experiments/ab_time.sh base fake1 ...
mkdir -p gpurun_out
for v in "$@"; do
  export SRB200_LIB=$PWD/experiments/ab/libsrb200_$v.so
  echo "=== $v" | tee -a gpurun_out/ab.log
  timeout 300 python benchmarks/bench_score.py --steps 10 2>&1 | tail -1 | tee -a gpurun_out/ab.log
done
