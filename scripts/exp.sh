#!/usr/bin/env bash
set -u
S="python scripts/sweep.py --gpus 1 --latency-reps 3"
echo "--- tile, per-thread write-back with skip"; CM_B200_TILE_BULKRED=0 $S --workloads NSown,K2own --only tile | grep '^{'
echo "--- tile, bulk"; $S --workloads NSown,K2own --only tile | grep '^{'
