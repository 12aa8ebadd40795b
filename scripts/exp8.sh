#!/usr/bin/env bash
N=${1:-8}
T="timeout -k 15 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
export CM_B200_TRACE=0
$T --master-port 29901 scripts/sweep.py --gpus $N --workloads K2,NS,K3 --latency-reps 3 --warmup 3 --only default,batch,deferred,background,background-ab1,background-ab3 2>&1 | grep -E '^\{|cm trace'
