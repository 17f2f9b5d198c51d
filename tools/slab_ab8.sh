#!/bin/bash
# tools/slab_ab8.sh NGPU: SM-driven vs copy-engine exchange at the sizes of config 5 (trimmed for an 8-GPU box)
G=${1:-8}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29533 tools/slab_check.py $1 40 2>&1 | grep -E "slab parity|\"slab\"|Error|error|unavailable" | tail -3 | cut -c1-150; }
echo "== SM N=8192"; run 8192
echo "== CE 2 chunks N=8192"; CHS_SLAB_CE=1 CHS_SLAB_CHUNKS=2 run 8192
echo "== SM N=16384"; run 16384
echo "== CE 4 chunks N=16384"; CHS_SLAB_CE=1 CHS_SLAB_CHUNKS=4 run 16384
