#!/bin/bash
# runs tools/quick_bench.py against every alternative build in tools/_libs (tuning experiments)
for f in tools/_libs/*.so; do
  echo "== $f"
  CHS_B200_LIB=$PWD/$f python tools/quick_bench.py 512 ${1:-256} 2>&1 | grep N=
done
