#!/bin/bash
# tools/sweep.sh [N] [batches]: quick_bench of every variants/*.so (tuning experiments on the GPU box)
N=${1:-512}; B=${2:-1024}
mkdir -p gpurun_out
for v in variants/*.so; do
  echo "== $v" | tee -a gpurun_out/sweep.txt
  CHS_B200_LIB=$PWD/$v timeout 300 python tools/quick_bench.py $N $B 2>&1 | tail -3 | tee -a gpurun_out/sweep.txt
done
