#!/usr/bin/env bash
# ncu captures of every kernel class of the path on ONE GPU (gpurun, one call). Each command runs plainly first.
set -u
mkdir -p gpurun_out/prof
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity --latency-reps 1 --configs none"
NCU="ncu --set full --clock-control none --import-source on -f"
cap() {  # cap <name> <kernel regex> <skip> <count> <cmd...>
  local name=$1 k=$2 s=$3 c=$4; shift 4
  "$@" > gpurun_out/prof/$name.plain.log 2>&1 && $NCU -k regex:$k -s $s -c $c -o gpurun_out/prof/$name "$@" > gpurun_out/prof/$name.ncu.log 2>&1
  echo "$name rc=$?"
}
cap k1_tile k_apply_tile 3 1 $B --workload K1
cap k2own_tile k_apply_tile 24 2 $B --workload K2own
cap nsown1_tile k_apply_tile 3 1 $B --workload NSown1 --apply-kernel tile
cap nsown_tile k_apply_tile 12 2 $B --workload NSown
cap nsown_red k_apply_red 12 2 $B --workload NSown --apply-kernel red
cap k4own_red k_apply_red 400 2 $B --workload K4own
cap pack_loopback k_pack 16 2 python scripts/profile_pack.py
# launch list of the K1 step (shares)
$B --workload K1 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/prof/launches_k1_n1.csv $B --workload K1 > gpurun_out/prof/launches.log 2>&1
ls -la gpurun_out/prof
