#!/bin/bash
# tools/ncu_mix.sh [N] [B]: ncu --set full of one k_mix launch (mixed column + row CTAs); text summary in gpurun_out/
N=${1:-512}; B=${2:-256}
mkdir -p gpurun_out
python tools/quick_bench.py $N $B > gpurun_out/plain_mix.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_mix' -s 30 -c 1 -f -o gpurun_out/prof_mix \
    python tools/quick_bench.py $N $B > gpurun_out/ncu_mix.log 2>&1
tail -1 gpurun_out/plain_mix.log
python tools/ncu_summary.py gpurun_out/prof_mix.ncu-rep 40 > gpurun_out/sum_mix.txt 2>&1
[ "$KEEP" = "1" ] || rm -f gpurun_out/prof_mix.ncu-rep
