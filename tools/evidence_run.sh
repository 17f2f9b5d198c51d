#!/bin/bash
# tools/evidence_run.sh TAG: the round's evidence set on one B200 -> gpurun_out/TAG_*
# (GPU tests, the contract bench with the driver's arguments, ncu launch list of a short bench run, traffic.json)
T=${1:-rX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/${T}_pytest.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python bench.py --steps 4 --warmup 3 --batch 256 --no-cpu > gpurun_out/${T}_small.json 2>/dev/null &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_ncu_launches.csv \
    python bench.py --steps 4 --warmup 3 --batch 256 --no-cpu > gpurun_out/${T}_ncu.log 2>&1
python tools/quick_bench.py 512 256 > gpurun_out/plain.log 2>&1 && python tools/make_traffic.py > gpurun_out/${T}_traffic.log 2>&1
tail -2 gpurun_out/${T}_pytest.txt; tail -c 600 gpurun_out/${T}_bench.json
