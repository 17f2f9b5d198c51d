# per-launch durations of a single simulation's step kernels (fixed-latency probe)
for N in 128 512; do
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_(row|col)' -s 100 -c 8 --csv python tools/quick_bench.py $N 1 2>/dev/null | grep -E "k_row|k_col" | awk -F'","' '{print $5, $NF}' | head -8
done
