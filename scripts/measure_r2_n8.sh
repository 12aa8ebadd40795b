#!/usr/bin/env bash
# Round-2 8-GPU measurement (one gpurun --gpus 8 call): the driver's bench command with the `configs` block (K2, K3, K4, NS),
# then the threshold-flush variant of the multi-process parity check over real NVLink / CUDA IPC.
set -u
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
t0=$(date +%s)
timeout -k 15 900 $T --master-port 29951 bench.py --gpus 8 --steps ${STEPS:-10} --warmup 3 2> gpurun_out/bench_n8.err | tail -1 > gpurun_out/bench_k1weak_n8.json
echo "bench rc=$? seconds=$(( $(date +%s) - t0 ))"
cut -c1-400 gpurun_out/bench_k1weak_n8.json; tail -5 gpurun_out/bench_n8.err
t0=$(date +%s)
[ "${MP_PARITY:-1}" = "1" ] && CM_TEST_FLUSH_THRESHOLD=50000 timeout -k 15 300 $T --master-port 29952 tests/mp_parity.py 2>&1 | grep -E "rank 0|rror" | cut -c1-300
echo "mp_parity(flush) rc=${PIPESTATUS[0]} seconds=$(( $(date +%s) - t0 ))"
