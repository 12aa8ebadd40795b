#!/usr/bin/env bash
# Measurement plan for the variants that were built without GPU time.
# Run ON a GPU box, e.g.
#   gpurun --gpus 2 --timeout 1500 -- 'bash scripts/measure_variants.sh 2 > gpurun_out/variants_n2.log 2>&1'
#   gpurun --gpus 8 --timeout 1800 -- 'bash scripts/measure_variants.sh 8 > gpurun_out/variants_n8.log 2>&1'
# Every bench line is a normal `bench.py` JSON line (device-timed, max over ranks); nothing here runs under a profiler.
set -u
N=${1:-2}
PORT=29811
run() {  # run <label> <env...> -- <bench args...>
  local label=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  echo "=== $label: ${envs[*]:-} bench.py $*"
  env "${envs[@]}" timeout -k 15 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
      --master-port $((PORT++)) bench.py --gpus "$N" --no-e2e --no-cpu-baseline --latency-reps 50 "$@" | tail -1
}
parity() {  # parity <label> <env...>
  local label=$1; shift
  echo "=== parity $label: $*"
  env "$@" timeout -k 15 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
      --master-port $((PORT++)) tests/mp_parity.py | tail -4
  echo "rc=${PIPESTATUS[0]}"
}

if [ "$N" = "1" ]; then
  # single GPU: the suite, the default bench line, then (only after both exited 0) the two ncu passes of
  # B200_PROFILING.md on the same command: launch list (shares of the step) and one full capture of the
  # scatter-add kernel
  python -m pytest tests -m gpu -x -q 2>&1 | tail -3
  python bench.py --steps 5 --warmup 3 | tail -1 | tee gpurun_out/bench_k1_n1.json
  python scripts/sweep.py --gpus 1 --workloads K1,K4 | grep '^{' | tee gpurun_out/sweep_n1.jsonl
  # the bulk reduce-add write-back is chosen once per process
  CM_B200_TEST_PTX_VARIANTS=1 timeout 900 python -m pytest tests/test_gpu_variants.py -q -m gpu -k "bulk_reduce or bulk_stores_child" 2>&1 | tail -3
  CM_B200_TILE_BULKRED=1 python scripts/sweep.py --gpus 1 --workloads K1 --only batch | grep '^{' | tee gpurun_out/sweep_n1_bulkred.jsonl
  B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --latency-reps 2"
  $B > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv \
      --log-file gpurun_out/launches_k1_n1.csv $B > gpurun_out/ncu_launches.log 2>&1
  $B > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_apply_tile -s 10 -c 2 \
      -o gpurun_out/k_apply_tile_k1_n1 -f $B > gpurun_out/ncu_full.log 2>&1
  exit 0
fi

# 1. parity of the new paths over real NVLink / NCCL first (small commits: thresholds lowered)
parity baseline CM_B200_PIPELINE=0
parity panels CM_B200_PIPELINE=1 CM_B200_PIPELINE_MIN_ELEMS=1
parity panels+flags CM_B200_PIPELINE=1 CM_B200_PIPELINE_MIN_ELEMS=1 CM_B200_FLAG_BARRIER=1
parity flags CM_B200_FLAG_BARRIER=1
parity devlayout CM_B200_DEVICE_LAYOUT=1
parity deferred CM_TEST_APPLY_POLICY=1

# 2. the whole variant matrix in ONE process per rank (cheap in box time): scripts/sweep.py
timeout -k 15 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
    --master-port $((PORT++)) scripts/sweep.py --gpus "$N" --workloads K1,K2,K3,K4 | grep '^{'
[ "${FULL:-0}" = "1" ] || exit 0

# 3. (FULL=1) the same comparisons through bench.py itself, one invocation per variant
# K1-weak: batch commit (round-1 default) against panel pipelining
run k1-batch CM_B200_PIPELINE=0 -- --steps 5 --warmup 3
for AB in 1 2 3; do
  run k1-panels4-ab$AB CM_B200_PIPELINE=1 CM_B200_PIPELINE_APPLY_BLOCKS=$AB -- --steps 5 --warmup 3
done
run k1-batch-cols8 CM_B200_PIPELINE=0 CM_B200_PACK_COLS=8 -- --steps 5 --warmup 3
run k1-panels4-ab2-cols8 CM_B200_PIPELINE=1 CM_B200_PIPELINE_APPLY_BLOCKS=2 CM_B200_PACK_COLS=8 -- --steps 5 --warmup 3
run k1-panels2 CM_B200_PIPELINE=1 CM_B200_PANELS=2 -- --steps 5 --warmup 3
run k1-panels4-flags CM_B200_PIPELINE=1 CM_B200_FLAG_BARRIER=1 -- --steps 5 --warmup 3

# 3. K2 (commit every 256 updates: 67 Mi elements per rank per commit) and K4 (tiny updates)
run k2-batch CM_B200_PIPELINE=0 -- --workload K2 --steps 3 --warmup 3
run k2-panels CM_B200_PIPELINE=1 -- --workload K2 --steps 3 --warmup 3
run k2-panels-flags CM_B200_PIPELINE=1 CM_B200_FLAG_BARRIER=1 -- --workload K2 --steps 3 --warmup 3
run k2-deferred CM_B200_PIPELINE=0 -- --workload K2 --steps 3 --warmup 3 --apply-policy deferred
run k3-batch CM_B200_PIPELINE=0 -- --workload K3 --steps 3 --warmup 3
run k3-batch-tile CM_B200_PIPELINE=0 -- --workload K3 --steps 3 --warmup 3 --apply-kernel tile
run k3-panels CM_B200_PIPELINE=1 -- --workload K3 --steps 3 --warmup 3
run k3-deferred CM_B200_PIPELINE=0 -- --workload K3 --steps 3 --warmup 3 --apply-policy deferred
run k4-batch CM_B200_PIPELINE=0 -- --workload K4 --steps 3 --warmup 3
run k4-flags CM_B200_FLAG_BARRIER=1 -- --workload K4 --steps 3 --warmup 3
run k4-flags-devlayout CM_B200_FLAG_BARRIER=1 CM_B200_DEVICE_LAYOUT=1 -- --workload K4 --steps 3 --warmup 3
run k2-devlayout CM_B200_PIPELINE=0 CM_B200_DEVICE_LAYOUT=1 -- --workload K2 --steps 3 --warmup 3
