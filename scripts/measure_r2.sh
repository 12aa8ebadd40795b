#!/usr/bin/env bash
# Round-2 single-GPU measurement: GPU suite, smoke, the driver's bench command, then (each only after its plain run
# exited 0) the ncu captures of scripts/profile.sh, summarised on the box (raw/source CSV pages; the .ncu-rep files stay
# behind when they would not fit gpurun_out's size limit).
#   gpurun --timeout 1500 -- 'bash scripts/measure_r2.sh > gpurun_out/r2_n1.log 2>&1; tail -30 gpurun_out/r2_n1.log'
set -u
mkdir -p gpurun_out/prof
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench N=1"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/bench_k1_n1.err | tail -1 | tee gpurun_out/bench_k1_n1.json | cut -c1-1500
if [ "${PROFILE:-1}" = "1" ]; then
  echo "=== ncu captures"; bash scripts/profile.sh 2>&1 | tail -12
  for r in gpurun_out/prof/*.ncu-rep; do
    b=${r%.ncu-rep}
    ncu -i "$r" --page raw --csv > "${b}_raw.csv" 2>/dev/null
  done
  python scripts/ncu_summary.py gpurun_out/prof/*.ncu-rep > gpurun_out/prof/summary.txt 2>&1
  for k in k1_tile pack_loopback; do
    [ -f gpurun_out/prof/$k.ncu-rep ] && ncu -i gpurun_out/prof/$k.ncu-rep --page source --csv > gpurun_out/prof/${k}_source.csv 2>/dev/null
  done
  sz=$(du -sm gpurun_out/prof | cut -f1)
  if [ "$sz" -gt 48 ]; then rm -f gpurun_out/prof/*.ncu-rep; echo "(.ncu-rep files dropped: $sz MiB)"; fi
  cat gpurun_out/prof/summary.txt | cut -c1-400
fi
