#!/usr/bin/env python
"""bench.py — assembled-update throughput of the ConvergentMatrix hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload K1|K2|K4]

One "step" = one full pass of the workload: every rank issues its U slice updates and the
matrix is committed (update -> owner map/pack -> exchange -> scatter-add -> commit). The
metric is BASELINE.json's: assembled update GB/s = sum(update elements) * sizeof(T) / time
from the first update() of a step to the return of its last commit(), max over ranks.

  value      inputs (payloads + index sets) already resident in HBM, device API
  e2e        same metric through the reference-facing host API (cm_update_slice with HOST
             pinned buffers; H2D inside the timed region; one element read back per step)
  roofline   the scatter-add kernel against the measured HBM copy bandwidth
  cpu_baseline / --impl reference
             the UNMODIFIED reference header over the in-repo UPC++ shim (oracle/_ref) on the
             box's host cores — the only place this file touches oracle/.

Multi-GPU: launched by torchrun, one process per GPU; ranks exchange through libcm_b200's own
NCCL communicator (the unique id is broadcast with torch.distributed).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRIDS = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}
METRIC = "assembled_update_GBps"


def workload_spec(name, nranks):
    """-> dict(dtype, m, n, nprow, npcol, mb, nb, lld, m_u, n_u, U (per rank), commit_every)."""
    nprow, npcol = GRIDS[nranks]
    if os.environ.get("CM_BENCH_GRID"):  # experiments only, e.g. CM_BENCH_GRID=2x1
        nprow, npcol = (int(x) for x in os.environ["CM_BENCH_GRID"].split("x"))
        assert nprow * npcol == nranks
    if name == "K1":  # BASELINE.json configs[1]; for N > 1 its weak-scaling family (local block fixed)
        return dict(name=f"K1{'' if nranks == 1 else '-weak'}: N={8192 * nprow}x{8192 * npcol} f64 MB=NB=64 grid {nprow}x{npcol}, "
                         f"10000 updates/rank of 256x256, sorted unique random index sets, one commit",
                    dtype=0, m=8192 * nprow, n=8192 * npcol, nprow=nprow, npcol=npcol, mb=64, nb=64, lld=8192,
                    m_u=256, n_u=256, U=10000, commit_every=0)
    if name == "K2":  # BASELINE.json configs[2] (8 GPUs); smaller grids keep N = 32768
        return dict(name=f"K2: N=32768 f64 MB=NB=64 grid {nprow}x{npcol}, 2048 updates/rank of 512x512, commit every 256",
                    dtype=0, m=32768, n=32768, nprow=nprow, npcol=npcol, mb=64, nb=64, lld=32768 // nprow,
                    m_u=512, n_u=512, U=2048, commit_every=256)
    if name == "K3":  # BASELINE.json configs[3]: clustered / skewed index sets (owner load imbalance)
        return dict(name=f"K3: N=65536 f64 MB=NB=64 grid {nprow}x{npcol}, 1024 updates/rank of 512x512, index sets = 8 runs of 64 "
                         f"consecutive indices at block boundaries, blocks owned by grid row/column 0 drawn with weight 4; commit every 256",
                    dtype=0, m=65536, n=65536, nprow=nprow, npcol=npcol, mb=64, nb=64, lld=65536 // nprow,
                    m_u=512, n_u=512, U=1024, commit_every=256, skew=True)
    if name == "K4":  # BASELINE.json configs[4]: tiny updates, many commits
        nprow, npcol = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}[nranks]
        return dict(name=f"K4: N=32768 f32 MB=NB=64 grid {nprow}x{npcol}, 125000 updates/rank of 32x32, commit every 1024",
                    dtype=1, m=32768, n=32768, nprow=nprow, npcol=npcol, mb=64, nb=64, lld=32768 // nprow,
                    m_u=32, n_u=32, U=125000, commit_every=1024)
    raise SystemExit(f"unknown workload {name}")


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def known_traffic(tag):
    """DRAM bytes per launch of the scatter-add kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(tag)
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def gen_inputs(spec, rank, torch, device):
    """Seeded synthetic inputs on the device: U payloads (column-major m_u x n_u, U(-1,1)) and
    U sorted, unique index sets per axis (as in the reference's test/cm_test.cpp:177-188)."""
    g = torch.Generator(device=device)
    g.manual_seed(0xC0FFEE + 1000 * rank)
    U, m_u, n_u = spec["U"], spec["m_u"], spec["n_u"]
    tdt = torch.float64 if spec["dtype"] == 0 else torch.float32
    data = torch.empty((U, n_u, m_u), dtype=tdt, device=device)
    chunk = max(1, (1 << 26) // (m_u * n_u))
    for a in range(0, U, chunk):
        b = min(U, a + chunk)
        data[a:b] = torch.rand((b - a, n_u, m_u), generator=g, device=device, dtype=tdt) * 2 - 1

    def skewed_sets(hi, k, blk, np_):
        # k/64 runs of 64 consecutive indices starting at 64*b; blocks b
        # drawn without replacement with weight 4 if b % np_ == 0 else 1
        nb_ = hi // blk
        w = torch.ones(nb_, device=device)
        w[torch.arange(0, nb_, np_, device=device)] = 4.0
        blocks = torch.multinomial(w.expand(U, nb_), k // blk, replacement=False, generator=g).sort(dim=1).values
        return (blocks[:, :, None] * blk + torch.arange(blk, device=device)[None, None, :]).reshape(U, k).contiguous()

    if spec.get("skew"):
        return data, skewed_sets(spec["m"], m_u, spec["mb"], spec["nprow"]), skewed_sets(spec["n"], n_u, spec["nb"], spec["npcol"])

    def index_sets(hi, k):
        out = torch.empty((U, k), dtype=torch.int64, device=device)
        step = max(1, (1 << 26) // hi)
        for a in range(0, U, step):
            b = min(U, a + step)
            keys = torch.rand((b - a, hi), generator=g, device=device)
            out[a:b] = keys.topk(k, dim=1).indices.sort(dim=1).values
        return out

    return data, index_sets(spec["m"], m_u), index_sets(spec["n"], n_u)


def run_step(M, batches, spec):
    """one pass of the workload on this rank: U updates with the configured commit cadence; under the
    deferred apply policy the step ends with a collective materialise, so that every update of the step
    has been scatter-added inside the timed region"""
    for b in batches:
        M.update_batch(b["batch"], b["device"])
        M.commit()
    if spec.get("deferred"):
        M.materialize()


def make_batches(cm, spec, data_ptr, ix_ptr, jx_ptr, esz, device):
    U, ce = spec["U"], spec["commit_every"]
    ce = U if ce <= 0 else ce
    out = []
    for a in range(0, U, ce):
        cnt = min(ce, U - a)
        out.append({"batch": cm.SliceBatch(cnt, spec["m_u"], spec["n_u"], data_ptr + a * spec["m_u"] * spec["n_u"] * esz,
                                           ix_ptr + a * spec["m_u"] * 8, jx_ptr + a * spec["n_u"] * 8, esz),
                    "device": device})
    return out


def bench_ours(args):
    import torch
    import torch.distributed as dist
    import cm_b200 as cm

    nranks = args.gpus
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != nranks:
        raise SystemExit(f"--gpus {nranks} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {nranks}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA GPU: libcm_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    uid = None
    if nranks > 1:
        dist.init_process_group("nccl", device_id=device)
        t = torch.zeros(128, dtype=torch.uint8, device=device)
        if rank == 0:
            t = torch.frombuffer(bytearray(cm.unique_id()), dtype=torch.uint8).to(device)
        dist.broadcast(t, 0)
        uid = bytes(t.cpu().numpy().tobytes())
    cm.init(rank, nranks, local_rank, uid)

    spec = workload_spec(args.workload, nranks)
    esz = 8 if spec["dtype"] == 0 else 4
    M = cm.Matrix(spec["dtype"], spec["m"], spec["n"], spec["nprow"], spec["npcol"], spec["mb"], spec["nb"], spec["lld"])
    M.set_profiling(True)
    if args.flush_threshold:
        M.set_flush_threshold(args.flush_threshold)
    if args.unsorted:
        M.set_sorted_sweep(False)
    if args.deterministic:
        M.set_deterministic(True)
    M.set_apply_kernel({"auto": 0, "red": 1, "tile": 2}[args.apply_kernel])
    if nranks > 1:
        M.set_exchange_mode({"auto": 0, "nccl": 1, "p2p": 2, "dma": 3}[args.exchange])
    spec["deferred"] = args.apply_policy != "eager" and nranks > 1 and M.exchange_mode() != 1
    if spec["deferred"]:
        M.set_apply_policy(2 if args.apply_policy == "background" else 1, args.defer_budget_gib << 30)
    data, ix, jx = gen_inputs(spec, rank, torch, device)
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(M.stream(), device=device)
    step_bytes = spec["U"] * spec["m_u"] * spec["n_u"] * esz  # per rank
    batches = make_batches(cm, spec, data.data_ptr(), ix.data_ptr(), jx.data_ptr(), esz, True)

    def sync_all():
        torch.cuda.synchronize()
        if nranks > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        """device time (CUDA events on the launching stream) of `steps` passes, max over ranks"""
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        e1.synchronize()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        sync_all()
        if nranks > 1:
            t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms, wall

    # ---- device-resident arm ("value") --------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        run_step(M, batches, spec)
    M.reset_stats()
    ms, wall = timed(lambda: run_step(M, batches, spec), args.steps)
    st = M.stats()
    value = nranks * step_bytes * args.steps / (ms * 1e-3) / 1e9
    # the timed region is tens of milliseconds; repeat the same steps (untimed, same count on every
    # rank) for ~0.6 s so that the 100 ms clock sampler records the clocks under this very load
    for _ in range(max(1, min(400, int(600.0 / max(ms / args.steps, 0.1))))):
        run_step(M, batches, spec)
    clocks = sampler.stop() if rank == 0 else None

    # ---- commit latency (BASELINE.json metric, second half) -------------------------------
    def commit_latency(with_update):
        one = cm.SliceBatch(1, spec["m_u"], spec["n_u"], data.data_ptr(), ix.data_ptr(), jx.data_ptr(), esz)
        lat = []
        for i in range(args.latency_reps + 3):
            sync_all()
            t0 = time.perf_counter()
            if with_update:
                M.update_batch(one, True)
            M.commit()
            lat.append((time.perf_counter() - t0) * 1e6)
        lat = np.array(lat[3:])
        return {"p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99))}

    lat_empty = commit_latency(False)
    lat_one = commit_latency(True)

    # ---- end-to-end arm: host pinned inputs through the reference-facing host API --------
    e2e = None
    if not args.no_e2e:
        Ue = min(spec["U"], args.e2e_updates) if args.e2e_updates else spec["U"]
        espec = dict(spec, U=Ue)
        h_data = torch.empty((Ue, spec["n_u"], spec["m_u"]), dtype=data.dtype).pin_memory()
        h_data.copy_(data[:Ue])
        h_ix = ix[:Ue].cpu().pin_memory()
        h_jx = jx[:Ue].cpu().pin_memory()
        hb = make_batches(cm, espec, h_data.data_ptr(), h_ix.data_ptr(), h_jx.data_ptr(), esz, False)
        probe = np.zeros(1, dtype=cm.NP_DTYPE[spec["dtype"]])

        def e2e_step():
            run_step(M, hb, espec)
            probe[0] = M.get(0, 0)  # device -> host read of a result element

        # context for the e2e number: what a plain pinned H2D copy of the same payload achieves here
        pc0, pc1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        scratch = torch.empty_like(data[:Ue])
        scratch.copy_(h_data, non_blocking=True)
        pc0.record()
        scratch.copy_(h_data, non_blocking=True)
        pc1.record()
        pc1.synchronize()
        pcie_gbs = h_data.numel() * esz / (pc0.elapsed_time(pc1) * 1e-3) / 1e9
        del scratch
        for _ in range(max(1, args.warmup // 2)):
            e2e_step()
        M.reset_stats()
        e_steps = max(1, min(args.steps, 3))
        ems, ewall = timed(e2e_step, e_steps)
        est = M.stats()
        e_bytes = Ue * spec["m_u"] * spec["n_u"] * esz
        # wall clock is the honest e2e time: the host API blocks on its own H2D copies
        e2e = {"value": nranks * e_bytes * e_steps / ewall / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": est["h2d_bytes"] // e_steps, "d2h_bytes_per_step": est["d2h_bytes"] // e_steps,
               "updates_per_rank": Ue, "steps": e_steps, "api": "cm_update_slices (one batch call per commit, host pinned payloads) + cm_commit + cm_get",
               "device_ms_per_step": ems / e_steps, "wall_ms_per_step": ewall * 1e3 / e_steps,
               "pinned_h2d_copy_GBps": pcie_gbs,
               "note": "strict reference semantics: every update's payload is on the device before update() returns"}

    # ---- roofline of the dominant kernel (scatter-add), timed live with CUDA events ----------
    peak, peak_src = measured_peak()
    roof = None
    if st["apply_ms"] > 0:
        alg_bytes = 3 * esz * st["apply_elements"]  # 3*s per owned element
        achieved = alg_bytes / (st["apply_ms"] * 1e-3) / 1e9
        tag = f"{args.workload}_n{nranks}"
        kname = "k_apply_tile" if st["tile_launches"] == st["apply_launches"] else (
            "k_apply_red" if st["tile_launches"] == 0 else "k_apply_tile+k_apply_red")
        # compulsory DRAM bytes if repeated hits combine on chip: s*E + 2*s*(distinct elements touched),
        # bounded above by 2*s*(local block) per launch
        local_bytes = spec["lld"] * (spec["n"] // spec["npcol"]) * esz
        compulsory = esz * st["apply_elements"] + min(2 * esz * st["apply_elements"], 2 * local_bytes * st["apply_launches"])
        roof = {"bound": "hbm", "kernel": kname, "achieved": achieved,
                "compulsory_bound": {"bytes_per_launch": compulsory // max(1, st["apply_launches"]),
                                     "achieved": compulsory / (st["apply_ms"] * 1e-3) / 1e9,
                                     "frac": compulsory / (st["apply_ms"] * 1e-3) / 1e9 / peak},
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": known_traffic(tag),
                "peak_source": peak_src, "algorithmic_bytes_per_element": 3 * esz,
                "launches": st["apply_launches"], "avg_launch_ms": st["apply_ms"] / max(1, st["apply_launches"]),
                "elements_per_launch": st["apply_elements"] // max(1, st["apply_launches"]),
                "kernel_share_of_step": st["apply_ms"] / ms}
        # roofline of the whole step: the slower of the HBM term (3*s bytes per owned
        # s-byte element) and the NVLink term (s*(P-1)/P bytes per sourced element at 770 GB/s per direction),
        # uniform owners assumed; value is aggregate payload GB/s
        hbm_ceiling = peak / 3.0
        nvl_ceiling = 770.0 * nranks / (nranks - 1) if nranks > 1 else float("inf")
        roof["step"] = {"bound": "nvlink" if nvl_ceiling < hbm_ceiling else "hbm",
                        "ceiling_GBps_per_gpu": min(hbm_ceiling, nvl_ceiling),
                        "frac": value / (min(hbm_ceiling, nvl_ceiling) * nranks), "assumes": "uniform owners"}
        if nranks > 1:
            xm = M.exchange_mode()
            p2p = xm == 2
            # p2p: the exchange IS the pack kernel (its stores cross NVLink); nccl: the all-to-all-v;
            # dma: the copy-engine pushes (timed only when the commit is not pipelined)
            xms = st["pack_ms"] if p2p else st["exchange_ms"]
            if xms > 0:
                roof["exchange"] = {"mode": {1: "nccl all-to-all-v", 2: "p2p (pack kernel stores over NVLink)",
                                             3: "dma (copy-engine pushes + flag waits)"}[xm],
                                    "bytes_sent_per_rank_per_step": st["bytes_sent"] // args.steps,
                                    "exchange_ms_per_step": xms / args.steps,
                                    "achieved_GBps_per_direction": st["bytes_sent"] / (xms * 1e-3) / 1e9,
                                    "peak_GBps_per_direction": 770.0}

    # ---- CPU baseline (rank 0, N = 1 only): reference header over the in-repo SMP shim -------
    cpu = None
    if rank == 0 and nranks == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference(spec, data, ix, jx, args.cpu_updates, reps=1)

    owned = None
    if nranks > 1:
        t_own = torch.tensor([float(st["elements_applied"])], dtype=torch.float64, device=device)
        gathered = [torch.zeros_like(t_own) for _ in range(nranks)]
        dist.all_gather(gathered, t_own)
        tot = sum(float(x) for x in gathered)
        owned = [round(float(x) / tot, 4) for x in gathered] if tot > 0 else None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": nranks, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64" if spec["dtype"] == 0 else "f32", "data": "synthetic",
            "config": {"workload": spec["name"], "updates_per_rank_per_step": spec["U"], "grid": [spec["nprow"], spec["npcol"]],
                       "l2": "inputs (payloads) per step are far larger than L2; no explicit flush",
                       "flush_threshold_elements": M.flush_threshold(), "sorted_sweep": not args.unsorted,
                       "deterministic": bool(args.deterministic), "apply_kernel": args.apply_kernel, "exchange": {1: "nccl", 2: "p2p", 3: "dma"}[M.exchange_mode()] if nranks > 1 else "none", "payload_bytes_per_rank_per_step": step_bytes,
                       "apply_policy": (args.apply_policy + " (extension: commits deliver, the step ends with cm_materialize inside the timed region)") if spec["deferred"] else "eager (reference semantics)",
                       "env": {k: v for k, v in sorted(os.environ.items()) if k.startswith("CM_B200_")}},
            "clocks": clocks, "e2e": e2e, "gpu_launches": st["kernel_launches"],
            "roofline": roof, "cpu_baseline": cpu,
            "commit_latency_us": {"empty": lat_empty, "one_update": lat_one},
            "owner_share_of_elements": owned,
            "wall_ms_per_step": wall * 1e3 / args.steps,
            "pack_ms_per_step": st["pack_ms"] / args.steps, "apply_ms_per_step": st["apply_ms"] / args.steps,
        }
        print(json.dumps(line), flush=True)
    M.destroy()
    cm.finalize()
    if nranks > 1:
        dist.destroy_process_group()


def cpu_reference(spec, data, ix, jx, n_updates, reps):
    """Time the unmodified reference header (oracle/_ref) on host cores. data/ix/jx: torch
    tensors or numpy arrays for rank 0 (used for every reference rank with a rank offset)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_bindings import RefShim, ref_available
    if not ref_available():
        return {"unavailable": "oracle/_ref/libcm_ref.so not built (needs /root/reference at build time)"}
    P = spec["nprow"] * spec["npcol"]
    Uc = min(spec["U"], n_updates)
    esz = 8 if spec["dtype"] == 0 else 4
    to_np = lambda t: t.cpu().numpy() if hasattr(t, "cpu") else np.asarray(t)
    d0, i0, j0 = to_np(data[:Uc]), to_np(ix[:Uc]), to_np(jx[:Uc])
    # every reference rank gets the same payloads and a rotated copy of the index sets
    datas = [np.ascontiguousarray(d0) for _ in range(P)]
    ixs = [np.ascontiguousarray(np.sort((i0 + 37 * r) % spec["m"], axis=1)) for r in range(P)]
    jxs = [np.ascontiguousarray(np.sort((j0 + 53 * r) % spec["n"], axis=1)) for r in range(P)]
    best = None
    ref = RefShim(spec["dtype"], spec["m"], spec["n"], spec["nprow"], spec["npcol"], spec["mb"], spec["nb"], spec["lld"])
    try:
        for _ in range(reps):
            t = ref.bench_slices(datas, ixs, jxs, spec["m_u"], spec["n_u"], spec["commit_every"], pin=True)
            if best is None or t[0] < best[0]:
                best = t
    finally:
        ref.close()
    gbs = P * Uc * spec["m_u"] * spec["n_u"] * esz / best[0] / 1e9
    return {"value": gbs, "unit": "GB/s", "cores": P, "kind": "reference",
            "sample": f"{Uc} of {spec['U']} updates per rank, {P} rank thread(s) pinned to {P} core(s) of {os.cpu_count()}, "
                      f"reference header unmodified over in-repo UPC++ SMP shim (oracle/_ref), g++ -O3 -DNOCHECK, "
                      f"default bin_flush_threshold=10000",
            "seconds": best[0], "update_seconds": best[1], "commit_seconds": best[2]}


def bench_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nranks = args.gpus
    spec = workload_spec(args.workload, nranks)
    esz = 8 if spec["dtype"] == 0 else 4
    rng = np.random.default_rng(0xC0FFEE)
    Uc = min(spec["U"], args.ref_updates)
    dt = np.float64 if spec["dtype"] == 0 else np.float32
    data = rng.uniform(-1, 1, (Uc, spec["n_u"], spec["m_u"])).astype(dt)
    ix = np.stack([np.sort(rng.choice(spec["m"], spec["m_u"], replace=False)) for _ in range(Uc)]).astype(np.int64)
    jx = np.stack([np.sort(rng.choice(spec["n"], spec["n_u"], replace=False)) for _ in range(Uc)]).astype(np.int64)
    sub = dict(spec, U=Uc)
    for _ in range(args.warmup):
        r = cpu_reference(sub, data, ix, jx, Uc, reps=1)
        if "unavailable" in r:
            print(json.dumps({"impl": "reference", "unavailable": r["unavailable"]}))
            return
    secs, last = 0.0, None
    for _ in range(args.steps):
        last = cpu_reference(sub, data, ix, jx, Uc, reps=1)
        secs += last["seconds"]
    P = spec["nprow"] * spec["npcol"]
    value = P * Uc * spec["m_u"] * spec["n_u"] * esz * args.steps / secs / 1e9
    cpu = dict(last, value=value)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": nranks, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": secs * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64" if spec["dtype"] == 0 else "f32", "data": "synthetic",
            "config": {"workload": spec["name"], "updates_per_rank_per_step": Uc, "grid": [spec["nprow"], spec["npcol"]],
                       "note": "each step is a bounded sample of the workload on host cores"},
            "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="K1", choices=["K1", "K2", "K3", "K4"])
    ap.add_argument("--flush-threshold", type=int, default=0, help="elements; 0 = library default")
    ap.add_argument("--unsorted", action="store_true", help="disable the column-sorted sweep")
    ap.add_argument("--deterministic", action="store_true")
    ap.add_argument("--apply-kernel", default="auto", choices=["auto", "red", "tile"])
    ap.add_argument("--exchange", default="auto", choices=["auto", "nccl", "p2p", "dma"],
                    help="auto/p2p: the pack kernel stores into the owners' arenas over NVLink; nccl: grouped send/recv")
    ap.add_argument("--apply-policy", default="eager", choices=["eager", "deferred", "background"],
                    help="deferred / background (N > 1, p2p/dma): commits only deliver (background: the logged messages are "
                         "scatter-added on a second stream under the following commits), the step ends with cm_materialize "
                         "(inside the timed region)")
    ap.add_argument("--defer-budget-gib", type=int, default=16)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-updates", type=int, default=0, help="updates per rank per e2e step (0 = all)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-updates", type=int, default=4000, help="bounded sample for the cpu_baseline leg")
    ap.add_argument("--ref-updates", type=int, default=1000, help="updates per rank per step of --impl reference")
    ap.add_argument("--latency-reps", type=int, default=30)
    args = ap.parse_args()
    if args.gpus not in GRIDS:
        raise SystemExit("--gpus must be 1, 2, 4 or 8")
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing hygiene: at least 3 warm-up steps
    if args.impl == "reference":
        bench_reference(args)
    else:
        bench_ours(args)


if __name__ == "__main__":
    main()
