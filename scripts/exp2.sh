#!/usr/bin/env bash
# 2-GPU experiments of round 2: panel count / apply-block sweep (one process per rank), then a per-phase timeline of one
# K1-weak commit (CM_B200_TRACE) for the default settings
N=${1:-2}
T="timeout -k 15 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$T --master-port 29941 scripts/sweep.py --gpus $N --workloads ${WL:-K1,NS} --latency-reps 3 --warmup 3 --steps 5 \
   --only ${ONLY:-default,batch,panels6,panels8,panels8-ab3} 2>&1 | grep -E '^\{|rror'
CM_B200_TRACE=1 CM_B200_TRACE_SKIP=5 $T --master-port 29942 bench.py --gpus $N --steps 3 --warmup 3 --configs none --no-e2e --no-parity \
   --no-cpu-baseline --latency-reps 1 2>&1 | grep -E 'cm trace' | head -60
