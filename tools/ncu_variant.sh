#!/bin/bash
# tools/ncu_variant.sh NAME [N] [B]: ncu --set full of one k_col<STEP> + one k_row<STEP> launch of variants/NAME.so;
# leaves text summaries in gpurun_out/ (the .ncu-rep is kept only with KEEP=1: 28 MB each, gpurun_out is capped at 64 MiB)
v=$1; N=${2:-512}; B=${3:-256}
mkdir -p gpurun_out
[ -f variants/$v.so ] && export CHS_B200_LIB=$PWD/variants/$v.so
python tools/quick_bench.py $N $B > gpurun_out/plain_$v.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_(row|col)' -s 60 -c 2 -f -o gpurun_out/prof_$v \
    python tools/quick_bench.py $N $B > gpurun_out/ncu_$v.log 2>&1
tail -1 gpurun_out/plain_$v.log
python tools/ncu_summary.py gpurun_out/prof_$v.ncu-rep 30 > gpurun_out/sum_$v.txt 2>&1
python tools/ncu_opmix.py gpurun_out/prof_$v.ncu-rep > gpurun_out/opmix_$v.txt 2>&1
(echo '== k_row'; python tools/ncu_lines.py gpurun_out/prof_$v.ncu-rep k_row 40; echo '== k_col'; python tools/ncu_lines.py gpurun_out/prof_$v.ncu-rep k_col 40) > gpurun_out/lines_$v.txt 2>&1
[ "$KEEP" = "1" ] || rm -f gpurun_out/prof_$v.ncu-rep
