#!/usr/bin/env python
"""Variant sweep in ONE process per rank (inputs, NCCL bootstrap and CUDA context are paid once):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29820 \
        scripts/sweep.py --gpus N [--workloads K1,K2,K4] [--steps 3] [--only name,name]

Every variant = environment knobs (read by cm_create) + harness options; for each one a fresh matrix
is created, warmed up and timed exactly like bench.py's `value` (device-resident inputs, CUDA events on
the library's stream, max over ranks). One JSON line per variant on rank 0. This is a tuning aid: the
numbers that count are bench.py's (run it with the winning knobs exported).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

KNOBS = ("CM_B200_PIPELINE", "CM_B200_PANELS", "CM_B200_PIPELINE_MIN_ELEMS", "CM_B200_PIPELINE_APPLY_BLOCKS", "CM_B200_EXCHANGE",
         "CM_B200_BG_MIN_ELEMS", "CM_B200_APPLY_POLICY")

# name, workloads it applies to, env, harness options
VARIANTS = [
    ("default", "*", {}, {}),
    ("batch", "*", {"CM_B200_PIPELINE": "0"}, {}),
    ("panels4-ab1", "*", {"CM_B200_PIPELINE": "1", "CM_B200_PIPELINE_APPLY_BLOCKS": "1"}, {}),
    ("panels4-ab2", "*", {"CM_B200_PIPELINE": "1", "CM_B200_PIPELINE_APPLY_BLOCKS": "2"}, {}),
    ("panels4-ab3", "*", {"CM_B200_PIPELINE": "1", "CM_B200_PIPELINE_APPLY_BLOCKS": "3"}, {}),
    ("panels2-ab2", "*", {"CM_B200_PIPELINE": "1", "CM_B200_PANELS": "2"}, {}),
    ("panels3-ab2", "*", {"CM_B200_PIPELINE": "1", "CM_B200_PANELS": "3"}, {}),
    ("panels6", "*", {"CM_B200_PANELS": "6"}, {}),
    ("panels8", "*", {"CM_B200_PANELS": "8"}, {}),
    ("panels8-ab3", "*", {"CM_B200_PANELS": "8", "CM_B200_PIPELINE_APPLY_BLOCKS": "3"}, {}),
    ("deferred", "*", {"CM_B200_PIPELINE": "0"}, {"deferred": 1}),
    ("background", "*", {"CM_B200_PIPELINE": "0"}, {"deferred": 2}),
    ("background-ab1", "*", {"CM_B200_PIPELINE": "0", "CM_B200_PIPELINE_APPLY_BLOCKS": "1"}, {"deferred": 2}),
    ("background-ab3", "*", {"CM_B200_PIPELINE": "0", "CM_B200_PIPELINE_APPLY_BLOCKS": "3"}, {"deferred": 2}),
    ("tile", "*", {}, {"apply_kernel": 2}),
    ("red", "*", {}, {"apply_kernel": 1}),
    ("nccl", "*", {"CM_B200_EXCHANGE": "nccl"}, {}),
    ("flush-2^27", "K1", {}, {"flush_threshold": 1 << 27, "n1_only": True}),
    ("flush-2^28", "K1", {}, {"flush_threshold": 1 << 28, "n1_only": True}),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--workloads", default="K1,K2,K4")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--latency-reps", type=int, default=20)
    ap.add_argument("--only", default="default")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import cm_b200 as cm
    nranks = args.gpus
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    uid = None
    if nranks > 1:
        dist.init_process_group("nccl", device_id=device)
        t = torch.zeros(128, dtype=torch.uint8, device=device)
        if rank == 0:
            t = torch.frombuffer(bytearray(cm.unique_id()), dtype=torch.uint8).to(device)
        dist.broadcast(t, 0)
        uid = bytes(t.cpu().numpy().tobytes())
    cm.init(rank, nranks, local, uid)
    only = set(filter(None, args.only.split(",")))

    def sync_all():
        torch.cuda.synchronize()
        if nranks > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for wl in args.workloads.split(","):
        spec0 = bench.workload_spec(wl, nranks)
        esz = 8 if spec0["dtype"] == 0 else 4
        data, ix, jx = bench.gen_inputs(spec0, rank, torch, device)
        torch.cuda.synchronize()
        step_bytes = spec0["U"] * spec0["m_u"] * spec0["n_u"] * esz
        for name, wls, env, opt in VARIANTS:
            if (wls != "*" and wl not in wls.split(",")) or (only and name not in only):
                continue
            if nranks == 1 and (env or opt.get("deferred")):
                continue  # those knobs concern the exchange
            if opt.get("n1_only") and nranks > 1:
                continue  # the flush threshold only bounds single-rank flushes
            for k in KNOBS:
                os.environ.pop(k, None)
            os.environ.update(env)
            spec = dict(spec0)
            M = cm.Matrix(spec["dtype"], spec["m"], spec["n"], spec["nprow"], spec["npcol"], spec["mb"], spec["nb"], spec["lld"])
            M.set_profiling(True)
            M.set_apply_kernel(opt.get("apply_kernel", 0))
            if opt.get("flush_threshold"):
                M.set_flush_threshold(opt["flush_threshold"])
            spec["deferred"] = bool(opt.get("deferred")) and nranks > 1 and M.exchange_mode() != 1
            if spec["deferred"]:
                M.set_apply_policy(int(opt["deferred"]), 16 << 30)
            stream = torch.cuda.ExternalStream(M.stream(), device=device)
            batches = bench.make_batches(cm, spec, data.data_ptr(), ix.data_ptr(), jx.data_ptr(), esz, True)
            for _ in range(args.warmup):
                bench.run_step(M, batches, spec)
            M.reset_stats()
            sync_all()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.steps):
                bench.run_step(M, batches, spec)
            e1.record(stream)
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            sync_all()
            st = M.stats()
            if nranks > 1:
                tt = torch.tensor([ms, st["pack_ms"], st["apply_ms"]], dtype=torch.float64, device=device)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                ms, pack_ms, apply_ms = (float(x) for x in tt)
            else:
                pack_ms, apply_ms = st["pack_ms"], st["apply_ms"]

            def latency(with_update):
                one = cm.SliceBatch(1, spec["m_u"], spec["n_u"], data.data_ptr(), ix.data_ptr(), jx.data_ptr(), esz)
                lat = []
                for _ in range(args.latency_reps + 3):
                    sync_all()
                    t0 = time.perf_counter()
                    if with_update:
                        M.update_batch(one, True)
                    M.commit()
                    lat.append((time.perf_counter() - t0) * 1e6)
                return float(np.percentile(np.array(lat[3:]), 50))
            lat0, lat1 = latency(False), latency(True)
            if rank == 0:
                print(json.dumps({"workload": wl, "variant": name, "n_gpus": nranks, "env": env, "opt": opt,
                                  "GBps": nranks * step_bytes * args.steps / (ms * 1e-3) / 1e9,
                                  "ms_per_step": ms / args.steps, "pack_ms_per_step": pack_ms / args.steps,
                                  "apply_ms_per_step": apply_ms / args.steps, "apply_launches_per_step": st["apply_launches"] / args.steps,
                                  "commit_us_empty_p50": lat0, "commit_us_one_update_p50": lat1}), flush=True)
            M.destroy()
        del data, ix, jx
        torch.cuda.empty_cache()
    cm.finalize()
    if nranks > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
