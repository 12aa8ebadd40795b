#!/usr/bin/env bash
# scratch: selected variants on N GPUs:  bash scripts/exp8.sh N
N=${1:-8}
T="timeout -k 15 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$T --master-port 29901 scripts/sweep.py --gpus $N --workloads K1 --only batch,panels4-ab2,panels2-ab2,panels4-ab3 2>/dev/null | grep '^{'
$T --master-port 29902 scripts/sweep.py --gpus $N --workloads K2 --only batch,panels4-ab2,deferred,background 2>/dev/null | grep '^{'
$T --master-port 29903 scripts/sweep.py --gpus $N --workloads K3 --only batch,panels4-ab2,background 2>/dev/null | grep '^{'
$T --master-port 29904 scripts/sweep.py --gpus $N --workloads K4 --only batch,batch-threads-auto,flags 2>/dev/null | grep '^{'
