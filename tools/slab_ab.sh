#!/bin/bash
# tools/slab_ab.sh NGPU [N ...]: slab path timing, SM-driven exchange vs copy-engine exchange (CHS_SLAB_CE=1) with 1/2/4 row chunks
G=${1:-2}; shift
NS=${@:-8192 16384}
run() { if [ "$G" = "1" ]; then python tools/slab_check.py $1 30; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29533 tools/slab_check.py $1 30; fi 2>&1 | grep -E "slab parity|\"slab\"|Error|error|unavailable" | tail -4 | cut -c1-150; }
for N in $NS; do
  echo "== SM-driven exchange N=$N"; run $N
  for c in 1 2 4; do echo "== copy engines, $c chunk(s) N=$N"; CHS_SLAB_CE=1 CHS_SLAB_CHUNKS=$c run $N; done
done
