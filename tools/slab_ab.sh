#!/bin/bash
# tools/slab_ab.sh NGPU [lib ...]: slab path timing (N = 8192, 16384) with the given library builds / CHS_SLAB_BULK=0
G=${1:-2}; shift
run() { if [ "$G" = "1" ]; then python tools/slab_check.py $1 30; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29533 tools/slab_check.py $1 30; fi 2>&1 | grep -E "\"slab\"|Error|error" | tail -2 | cut -c1-130; }
for N in 8192 16384; do
  echo "== stores (CHS_SLAB_BULK=0) N=$N"; CHS_SLAB_BULK=0 run $N
  echo "== bulk, default build N=$N"; run $N
  for v in "$@"; do echo "== bulk, variants/$v.so N=$N"; CHS_B200_LIB=$PWD/variants/$v.so run $N; done
done
