#!/usr/bin/env bash
# A/B of the two tile walks on one GPU: ordered (one barrier per item; the deterministic mode) against item-parallel
# with shared-memory atomics (default), on K1 and on the single-GPU stand-ins of the 8-GPU owners' shapes
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-parity --latency-reps 1 --configs none"
for wl in K1 K2own NSown1 NSown; do
  for ord in 1 0; do
    echo -n "$wl ordered=$ord "
    CM_B200_TILE_ORDERED=$ord $B --workload $wl --apply-kernel tile 2>&1 | tail -1 | python -c "
import json,sys
b=json.loads(sys.stdin.read()); r=b['roofline']
print('GB/s %.1f  ms/step %.3f  apply_ms/step %.3f  launches %d  avg_launch_ms %.3f  compulsory_frac %.3f' % (b['value'], b['ms_per_step'], b['apply_ms_per_step'], r['launches'], r['avg_launch_ms'], r['compulsory_bound']['frac']))"
  done
done
