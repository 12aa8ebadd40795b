#!/usr/bin/env bash
# NVLink data counters (nvidia-smi nvlink -gt d) around a 2-GPU run of W + S K1-weak steps: the counter deltas against the
# library's own byte accounting (bytes_sent per step x steps). ncu cannot attach to a multi-rank run; this is the
# hardware-side check of the exchange volume.
W=${W:-3}; S=${S:-47}
nvidia-smi nvlink -gt d > gpurun_out/nvlink_before.txt 2>&1
timeout -k 15 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29991 \
  scripts/sweep.py --gpus 2 --workloads K1 --only default --warmup $W --steps $S --latency-reps 1 2>&1 | grep '^{' > gpurun_out/nvlink_run.json
nvidia-smi nvlink -gt d > gpurun_out/nvlink_after.txt 2>&1
python - <<PY
import json, re
def totals(path):
    out, gpu = {}, None
    for line in open(path):
        m = re.match(r"GPU (\d+):", line)
        if m: gpu = int(m.group(1)); out[gpu] = [0, 0]; continue
        m = re.search(r"Data (Tx|Rx): (\d+) KiB", line)
        if m and gpu is not None: out[gpu][0 if m.group(1) == "Tx" else 1] += int(m.group(2)) * 1024
    return out
a, b = totals("gpurun_out/nvlink_before.txt"), totals("gpurun_out/nvlink_after.txt")
run = json.loads(open("gpurun_out/nvlink_run.json").read().splitlines()[0])
print("run:", {k: run[k] for k in ("GBps", "ms_per_step", "pack_ms_per_step")})
for g in sorted(b):
    if g in a: print(f"GPU {g}: tx {(b[g][0]-a[g][0])/1e9:.2f} GB  rx {(b[g][1]-a[g][1])/1e9:.2f} GB over $W + $S steps (+ 4 tiny commits)")
PY
