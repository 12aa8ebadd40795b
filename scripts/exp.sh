#!/usr/bin/env bash
# scratch experiments on one GPU: bash scripts/exp.sh > gpurun_out/exp.log
set -u
S="python scripts/sweep.py --gpus 1 --workloads K1"
$S --only batch,flush-2^28,flush-2^29,flush-2^30 | grep '^{'
echo "--- bulkred off"
CM_B200_TILE_BULKRED=0 $S --only batch,flush-2^30 | grep '^{'
echo "--- prefetch 0/1/3"
for pf in 0 1 3; do CM_B200_TILE_PREFETCH=$pf $S --only flush-2^30 | grep '^{'; done
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
