#!/usr/bin/env python
"""bench.py — assembled-update throughput of the ConvergentMatrix hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload K1|K2|K3|K4|NS]

One "step" = one full pass of the workload: every rank issues its U slice updates and the
matrix is committed (update -> owner map/pack -> exchange -> scatter-add -> commit). The
metric is BASELINE.json's: assembled update GB/s = sum(update elements) * sizeof(T) / time
from the first update() of a step to the return of its last commit(), max over ranks.

  value      inputs (payloads + index sets) already resident in HBM, device API
  e2e        same metric through the reference-facing host API (cm_update_slice with HOST
             pinned buffers; H2D inside the timed region; one element read back per step)
  roofline   the scatter-add kernel against the measured HBM copy bandwidth; `step`: the whole step
             against the slower of its HBM and NVLink terms
  parity     AFTER the timed region: the matrix is reset, the first updates of the same workload are
             replayed on every rank over the real exchange path, and every rank's local block and
             per-owner wire stream are compared with the reference (oracle/_ref) on identical inputs
  configs    (--gpus 8) the other workloads BASELINE.json names for 8 GPUs — K2, K3, K4 — and the
             north-star workload NS, measured in the same invocation the same way
  cpu_baseline / --impl reference
             the UNMODIFIED reference header over the in-repo UPC++ shim (oracle/_ref) on the
             box's host cores — with the parity leg the only places this file touches oracle/.

Multi-GPU: launched by torchrun, one process per GPU; ranks exchange through libcm_b200's own
NCCL communicator (the unique id is broadcast with torch.distributed).
"""
import argparse
import hashlib
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
NVLINK_PEAK = 770.0  # GB/s per direction per GPU, measured peer copy (B200_PROFILING.md)


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
    if name == "NS":  # BASELINE.json north_star target: N=65536, uniform random 512x512 updates
        return dict(name=f"NS (north-star target): N=65536 f64 MB=NB=64 grid {nprow}x{npcol}, 2048 updates/rank of 512x512 (K2's count), "
                         f"sorted unique random index sets, one commit (assemble, then commit: the reference's own usage)",
                    dtype=0, m=65536, n=65536, nprow=nprow, npcol=npcol, mb=64, nb=64, lld=65536 // nprow,
                    m_u=512, n_u=512, U=2048, commit_every=0)
    if name == "K4":  # BASELINE.json configs[4]: tiny updates, many commits
        nprow, npcol = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}[nranks]
        return dict(name=f"K4: N=32768 f32 MB=NB=64 grid {nprow}x{npcol}, 125000 updates/rank of 32x32, commit every 1024",
                    dtype=1, m=32768, n=32768, nprow=nprow, npcol=npcol, mb=64, nb=64, lld=32768 // nprow,
                    m_u=32, n_u=32, U=125000, commit_every=1024)
    # single-GPU stand-ins for what ONE owner of an 8-GPU config scatter-adds (kernel tuning and ncu captures:
    # ncu cannot wrap a multi-rank run). The owner of a 2x4 grid receives, per update of any rank, the segment
    # Mat[R_p, C_q]: half the rows, a quarter of the columns — here fed as whole updates of that shape on a 1x1
    # grid whose block has the owner's shape, 8 sources' worth per commit.
    if name == "K2own" and nranks == 1:
        return dict(name="K2own: the scatter-add of one K2 owner (block 16384x8192 f64, 256x128 segments, 2048 per commit)",
                    dtype=0, m=16384, n=8192, nprow=1, npcol=1, mb=64, nb=64, lld=16384, m_u=256, n_u=128, U=16384,
                    commit_every=2048)
    if name == "NSown" and nranks == 1:
        return dict(name="NSown: the scatter-add of one NS owner (block 32768x16384 f64, 256x128 segments, 2048 per commit)",
                    dtype=0, m=32768, n=16384, nprow=1, npcol=1, mb=64, nb=64, lld=32768, m_u=256, n_u=128, U=8192,
                    commit_every=2048)
    if name == "NSown1" and nranks == 1:
        return dict(name="NSown1: the scatter-add of one NS owner when the step is ONE commit (block 32768x16384 f64, 8192 segments of 256x128)",
                    dtype=0, m=32768, n=16384, nprow=1, npcol=1, mb=64, nb=64, lld=32768, m_u=256, n_u=128, U=8192,
                    commit_every=0)
    if name == "K4own" and nranks == 1:
        return dict(name="K4own: the scatter-add of one K4 owner (block 8192x16384 f32, 8x16 segments, 8192 per commit)",
                    dtype=1, m=8192, n=16384, nprow=1, npcol=1, mb=64, nb=64, lld=8192, m_u=8, n_u=16, U=1 << 20,
                    commit_every=8192)
    raise SystemExit(f"unknown workload {name} for {nranks} GPU(s)")


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


def gen_inputs(spec, rank, torch, device, U=None):
    """Seeded synthetic inputs on the device: U payloads (column-major m_u x n_u, U(-1,1)) and
    U sorted, unique index sets per axis (as in the reference's test/cm_test.cpp:177-188)."""
    g = torch.Generator(device=device)
    g.manual_seed(0xC0FFEE + 1000 * rank)
    U = spec["U"] if U is None else U
    m_u, n_u = spec["m_u"], spec["n_u"]
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


class Ctx:
    """the process-wide runtime of one bench invocation (one rank)"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import cm_b200 as cm
        self.torch, self.dist, self.cm = torch, dist, cm
        self.nranks = args.gpus
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world != self.nranks:
            raise SystemExit(f"--gpus {self.nranks} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {self.nranks}")
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA GPU: libcm_b200 has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)
        uid = None
        if self.nranks > 1:
            dist.init_process_group("nccl", device_id=self.device)
            t = torch.zeros(128, dtype=torch.uint8, device=self.device)
            if self.rank == 0:
                t = torch.frombuffer(bytearray(cm.unique_id()), dtype=torch.uint8).to(self.device)
            dist.broadcast(t, 0)
            uid = bytes(t.cpu().numpy().tobytes())
        cm.init(self.rank, self.nranks, self.local_rank, uid)

    def sync_all(self):
        self.torch.cuda.synchronize()
        if self.nranks > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def timed(self, stream, fn, steps):
        """device time (CUDA events on the launching stream) of `steps` passes, max over ranks"""
        torch = self.torch
        self.sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        e1.synchronize()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        self.sync_all()
        if self.nranks > 1:
            t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=self.device)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms, wall

    def gather_floats(self, vals):
        """[nranks][len(vals)] on every rank"""
        torch = self.torch
        t = torch.tensor(list(vals), dtype=torch.float64, device=self.device)
        if self.nranks == 1:
            return [t.tolist()]
        out = [torch.zeros_like(t) for _ in range(self.nranks)]
        self.dist.all_gather(out, t)
        return [x.tolist() for x in out]

    def close(self):
        self.cm.finalize()
        if self.nranks > 1:
            self.dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
# parity leg: correctness evidence inside the very run that is timed (outside every timed region)
# ---------------------------------------------------------------------------------------------------
def _sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def parity_leg(ctx, M, spec, data, ix, jx, n_upd):
    """Reset the matrix, replay `n_upd` updates per rank through the same device API and exchange path that
    was just timed, commit, and compare with the reference (oracle/_ref: the unmodified reference header
    over the in-repo UPC++ shim, P rank threads on rank 0's host) on identical inputs:
      * every rank's whole local block, normwise relative error (bar 1e-12 fp64 / 1e-5 fp32),
      * the per-(source, owner) wire streams, bit for bit (sha256 of the expanded (inds, data) streams),
      * deterministic mode: two replays give bit-identical blocks on every rank.
    Matrices whose P local blocks exceed 8 GiB are checked on updates whose COLUMN indices are drawn from the
    first n' global columns (fresh seeded sets; rows as in the workload): the reference is then instantiated
    with the same <T,NPROW,NPCOL,MB,NB,LLD> and global size m x n', so local indices are identical and its
    blocks must equal the first columns of ours; the remaining columns must be zero."""
    torch, dist, cm = ctx.torch, ctx.dist, ctx.cm
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_bindings import RefShim, ref_available
    P, rank = ctx.nranks, ctx.rank
    esz = 8 if spec["dtype"] == 0 else 4
    tol = 1e-12 if spec["dtype"] == 0 else 1e-5
    m_u, n_u = spec["m_u"], spec["n_u"]
    U = min(n_upd, spec["U"])
    have_ref = ref_available() and RefShim.have_config(spec["dtype"], spec["nprow"], spec["npcol"], spec["mb"], spec["nb"], spec["lld"])
    if not have_ref:
        return {"checked": False, "why": "oracle/_ref/libcm_ref.so (or this <T,NPROW,NPCOL,MB,NB,LLD> instantiation) is not available"}
    # columns restricted to n' when the P reference blocks would not fit comfortably in host memory
    n_ref = spec["n"]
    block_bytes = spec["lld"] * (spec["n"] // spec["npcol"]) * esz
    restricted = P * block_bytes > (8 << 30)
    if restricted:
        n_ref = spec["nb"] * spec["npcol"] * 16  # 16 NB-blocks of columns per owner column
        g = torch.Generator(device=ctx.device)
        g.manual_seed(0xBEEF + rank)
        keys = torch.rand((U, n_ref), generator=g, device=ctx.device)
        pjx = keys.topk(n_u, dim=1).indices.sort(dim=1).values.contiguous()
    else:
        pjx = jx[:U].contiguous()
    pdata, pix = data[:U].contiguous(), ix[:U].contiguous()
    batch = cm.SliceBatch(U, m_u, n_u, pdata.data_ptr(), pix.data_ptr(), pjx.data_ptr(), esz)
    torch.cuda.synchronize()

    def replay(wire=False, deterministic=False):
        M.reset()
        M.set_deterministic(deterministic)
        if wire:
            M.keep_wire(True)
        M.update_batch(batch, True)
        M.commit()
        hashes = None
        if wire:
            hashes = [_sha(*M.wire(d)) for d in range(P)]
            M.keep_wire(False)
        return M.local(), hashes

    mine, _ = replay()                      # the path that was timed
    mine_w, wire_hashes = replay(wire=True)  # same inputs with wire capture
    det1, _ = replay(deterministic=True)
    det2, _ = replay(deterministic=True)
    M.set_deterministic(False)
    M.reset()
    det_ok = det1.tobytes() == det2.tobytes()

    # inputs of every rank -> rank 0 (device tensors over NCCL), expected blocks back
    def gather0(t):
        if P == 1:
            return [t]
        out = [torch.empty_like(t) for _ in range(P)] if rank == 0 else None
        dist.gather(t, out, dst=0)
        return out

    g_data, g_ix, g_jx = gather0(pdata), gather0(pix), gather0(pjx)
    n_loc_ref = None
    expected = None
    ref_wire = None
    t_ref = 0.0
    if rank == 0:
        t0 = time.perf_counter()
        ref = RefShim(spec["dtype"], spec["m"], n_ref, spec["nprow"], spec["npcol"], spec["mb"], spec["nb"], spec["lld"], record_wire=True)
        try:
            ref.set_flush_threshold(2 ** 31 - 1)  # one message per (source, owner): the stream order is what is compared
            for r in range(P):
                d, i, j = g_data[r].cpu().numpy(), g_ix[r].cpu().numpy(), g_jx[r].cpu().numpy()
                for u in range(U):
                    ref.update_slice(r, d[u], m_u, n_u, m_u, False, i[u], j[u])
            ref.commit()
            expected = [ref.local(r) for r in range(P)]
            n_loc_ref = [ref.n_local(r) for r in range(P)]
            ref_wire = [[_sha(*ref.wire(s, d)) for d in range(P)] for s in range(P)]
        finally:
            ref.close()
        t_ref = time.perf_counter() - t0
    # expected block of rank r -> rank r
    nl = torch.zeros(1, dtype=torch.int64, device=ctx.device)
    if P > 1:
        nl_list = [torch.tensor([x], dtype=torch.int64, device=ctx.device) for x in n_loc_ref] if rank == 0 else None
        dist.scatter(nl, nl_list, src=0)
    else:
        nl[0] = n_loc_ref[0]
    n_exp = int(nl[0]) * spec["lld"]
    tdt = torch.float64 if spec["dtype"] == 0 else torch.float32
    exp_t = torch.empty(n_exp, dtype=tdt, device=ctx.device)
    if P > 1:
        exp_list = [torch.from_numpy(e).to(ctx.device) for e in expected] if rank == 0 else None
        dist.scatter(exp_t, exp_list, src=0)
        del exp_list
    else:
        exp_t.copy_(torch.from_numpy(expected[0]))

    def cmp(block):
        got = torch.from_numpy(block).to(ctx.device)
        head, tail = got[:n_exp], got[n_exp:]
        d2 = float(((head.double() - exp_t.double()) ** 2).sum()) + float((tail.double() ** 2).sum())
        b2 = float((exp_t.double() ** 2).sum())
        mx = float((head - exp_t).abs().max()) if n_exp else 0.0
        return d2, b2, mx

    a = cmp(mine)
    b = cmp(mine_w)
    rows = ctx.gather_floats([a[0], a[1], a[2], b[0], b[1], 1.0 if det_ok else 0.0])
    if P > 1:
        all_hashes = [None] * P
        dist.all_gather_object(all_hashes, wire_hashes)
    else:
        all_hashes = [wire_hashes]
    if rank != 0:
        return None
    per_rank = [float(np.sqrt(r[0] / r[1])) if r[1] > 0 else float(np.sqrt(r[0])) for r in rows]
    glob = float(np.sqrt(sum(r[0] for r in rows) / max(sum(r[1] for r in rows), 1e-300)))
    glob_w = float(np.sqrt(sum(r[3] for r in rows) / max(sum(r[4] for r in rows), 1e-300)))
    wire_ok = all(all_hashes[s][d] == ref_wire[s][d] for s in range(P) for d in range(P))
    return {"checked": True, "normwise_rel": glob, "normwise_rel_per_rank_max": max(per_rank), "max_abs_diff": max(r[2] for r in rows),
            "tolerance": tol, "ok": bool(glob <= tol and max(per_rank) <= tol and glob_w <= tol),
            "updates": U * P, "updates_per_rank": U, "elements": U * P * m_u * n_u,
            "wire_bit_exact": bool(wire_ok), "normwise_rel_wire_capture_run": glob_w,
            "deterministic_bit_exact_run_to_run": bool(all(r[5] == 1.0 for r in rows)),
            "oracle": "oracle/_ref: unmodified reference header over the in-repo UPC++ shim, %d rank thread(s) on rank 0's host" % P,
            "inputs": ("the first %d updates per rank of the timed workload" % U) if not restricted else
                      ("%d updates per rank: payloads and row sets of the timed workload, column sets drawn from the first %d global "
                       "columns (reference instantiated m x %d, same <T,NPROW,NPCOL,MB,NB,LLD>; our remaining columns must be zero)"
                       % (U, n_ref, n_ref)),
            "path": "cm_reset + cm_update_slices_device + cm_commit on every rank (the timed API and exchange path), outside the timed region",
            "reference_seconds": t_ref}


# ---------------------------------------------------------------------------------------------------
# one workload, measured on the ranks of this invocation
# ---------------------------------------------------------------------------------------------------
def measure_workload(ctx, args, wl, steps, warmup, main):
    torch, dist, cm = ctx.torch, ctx.dist, ctx.cm
    nranks, rank, device = ctx.nranks, ctx.rank, ctx.device
    t_begin = time.perf_counter()
    spec = workload_spec(wl, nranks)
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
        M.set_exchange_mode({"auto": 0, "nccl": 1, "p2p": 2}[args.exchange])
    spec["deferred"] = args.apply_policy != "eager" and nranks > 1 and M.exchange_mode() != 1
    if spec["deferred"]:
        M.set_apply_policy(2 if args.apply_policy == "background" else 1, args.defer_budget_gib << 30)
    data, ix, jx = gen_inputs(spec, rank, torch, device)
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(M.stream(), device=device)
    step_bytes = spec["U"] * spec["m_u"] * spec["n_u"] * esz  # per rank
    batches = make_batches(cm, spec, data.data_ptr(), ix.data_ptr(), jx.data_ptr(), esz, True)

    # ---- device-resident arm ("value") --------------------------------------------------
    sampler = ClockSampler(ctx.local_rank)
    if rank == 0 and main:
        sampler.start()
    for _ in range(warmup):
        run_step(M, batches, spec)
    M.reset_stats()
    ms, wall = ctx.timed(stream, lambda: run_step(M, batches, spec), steps)
    st = M.stats()
    value = nranks * step_bytes * steps / (ms * 1e-3) / 1e9
    clocks = None
    if main:
        # the timed region is tens of milliseconds; repeat the same steps (untimed, same count on every
        # rank) for ~0.6 s so that the 100 ms clock sampler records the clocks under this very load
        for _ in range(max(1, min(400, int(600.0 / max(ms / steps, 0.1))))):
            run_step(M, batches, spec)
        clocks = sampler.stop() if rank == 0 else None

    # ---- commit latency (BASELINE.json metric, second half) -------------------------------
    def commit_latency(with_update):
        one = cm.SliceBatch(1, spec["m_u"], spec["n_u"], data.data_ptr(), ix.data_ptr(), jx.data_ptr(), esz)
        lat = []
        for i in range(args.latency_reps + 3):
            ctx.sync_all()
            t0 = time.perf_counter()
            if with_update:
                M.update_batch(one, True)
            M.commit()
            lat.append((time.perf_counter() - t0) * 1e6)
        lat = np.array(lat[3:])
        return {"p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99))}

    lat_empty = commit_latency(False)
    lat_one = commit_latency(True)
    # per-commit time inside the timed steps (K4: the latency-bound end)
    ncommits = len(batches)

    # ---- end-to-end arm: host pinned inputs through the reference-facing host API --------
    e2e = None
    if main and not args.no_e2e:
        e2e = measure_e2e(ctx, args, M, spec, data, ix, jx, stream, steps, warmup)

    # ---- roofline ---------------------------------------------------------------------------
    peak, peak_src = measured_peak()
    per_rank = ctx.gather_floats([st["elements_applied"] / steps, st["bytes_sent"] / steps, st["bytes_received"] / steps,
                                  st["apply_ms"] / steps, st["pack_ms"] / steps])
    owned = [r[0] for r in per_rank]
    tot_owned = sum(owned)
    sourced = spec["U"] * spec["m_u"] * spec["n_u"]  # elements every rank issues per step
    # step roofline, from the REALISED owner histogram: per GPU the slower of
    #   HBM:    3*s bytes per element it owns
    #   NVLink: s bytes per element crossing it, per direction, at the busier end: what the owner receives from
    #           the other ranks (every rank sources the same amount, so (P-1)/P of what it owns), or what a source
    #           sends away (everything it issues but its own share of its own updates)
    t_hbm = max(3 * esz * o / (peak * 1e9) for o in owned) if tot_owned else 0.0
    t_nvl = 0.0
    if nranks > 1 and tot_owned:
        for o in owned:
            ingress = esz * o * (nranks - 1) / nranks
            egress = esz * sourced * (1 - o / tot_owned)
            t_nvl = max(t_nvl, ingress / (NVLINK_PEAK * 1e9), egress / (NVLINK_PEAK * 1e9))
    t_roof = max(t_hbm, t_nvl)
    step_roof = {"bound": "nvlink" if t_nvl > t_hbm else "hbm", "roofline_ms_per_step": t_roof * 1e3,
                 "ceiling_GBps_aggregate": nranks * step_bytes / t_roof / 1e9 if t_roof else None,
                 "frac": (t_roof * 1e3) / (ms / steps) if t_roof else None,
                 "from": "realised owner histogram; HBM %.1f GB/s, NVLink %.0f GB/s per direction" % (peak, NVLINK_PEAK)}
    roof = None
    if st["apply_ms"] > 0:
        alg_bytes = 3 * esz * st["apply_elements"]  # 3*s per owned element
        achieved = alg_bytes / (st["apply_ms"] * 1e-3) / 1e9
        tag = f"{wl}_n{nranks}"
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
                "kernel_share_of_step": st["apply_ms"] / ms, "step": step_roof}
        if nranks > 1 and st["pack_ms"] > 0:
            roof["exchange"] = {"mode": {1: "nccl all-to-all-v", 2: "p2p (pack kernel stores over NVLink)"}[M.exchange_mode()],
                                "bytes_sent_per_rank_per_step": st["bytes_sent"] // steps,
                                "pack_ms_per_step": st["pack_ms"] / steps,
                                "achieved_GBps_per_direction": st["bytes_sent"] / (st["pack_ms"] * 1e-3) / 1e9,
                                "peak_GBps_per_direction": NVLINK_PEAK,
                                "note": "pack_ms spans map + scan + counts exchange + pack of every commit; with pipelined commits "
                                        "the scatter-add of earlier panels runs inside it"}

    # ---- parity leg (outside every timed region) -------------------------------------------------
    parity = None
    if not args.no_parity:
        parity = parity_leg(ctx, M, spec, data, ix, jx, args.parity_updates)

    cpu = None
    if main and rank == 0 and nranks == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference(spec, data, ix, jx, args.cpu_updates, reps=1)

    out = None
    if rank == 0:
        out = {"workload": spec["name"], "value": value, "unit": "GB/s", "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
               "updates_per_rank_per_step": spec["U"], "commits_per_step": ncommits, "ms_per_commit": ms / steps / ncommits,
               "updates_per_second": nranks * spec["U"] * steps / (ms * 1e-3),
               "dtype": "f64" if spec["dtype"] == 0 else "f32", "grid": [spec["nprow"], spec["npcol"]],
               "step_roofline": step_roof, "owner_share_of_elements": [round(o / tot_owned, 4) for o in owned] if tot_owned else None,
               "commit_latency_us": {"empty": lat_empty, "one_update": lat_one},
               "pack_ms_per_step": max(r[4] for r in per_rank), "apply_ms_per_step": max(r[3] for r in per_rank),
               "scatter_add_kernel": roof["kernel"] if roof else None,
               "apply_policy": args.apply_policy if spec["deferred"] else "eager", "parity": parity,
               "harness_seconds": time.perf_counter() - t_begin,  # input generation, warm-up, timed steps, latency probes, parity leg
               "_main": {"st": st, "roof": roof, "clocks": clocks, "e2e": e2e, "cpu": cpu, "wall": wall, "spec": spec,
                         "flush_threshold": M.flush_threshold(), "exchange": {1: "nccl", 2: "p2p"}.get(M.exchange_mode(), "none") if nranks > 1 else "none",
                         "step_bytes": step_bytes}}
    M.destroy()
    del data, ix, jx
    torch.cuda.empty_cache()
    return out


def measure_e2e(ctx, args, M, spec, data, ix, jx, stream, steps, warmup):
    """the same pass through the host API with HOST pinned buffers: H2D inside the timed region"""
    torch, cm = ctx.torch, ctx.cm
    esz = 8 if spec["dtype"] == 0 else 4
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

    def e2e_step_block():
        run_step(M, hb, espec)
        M.local_host_view()     # get_local_data(): the whole local block refreshed on the host

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
    for _ in range(max(1, warmup // 2)):
        e2e_step()
    M.reset_stats()
    e_steps = max(1, min(steps, 3))
    ems, ewall = ctx.timed(stream, e2e_step, e_steps)
    est = M.stats()
    e_bytes = Ue * spec["m_u"] * spec["n_u"] * esz
    e2e_step_block()
    M.reset_stats()
    bms, bwall = ctx.timed(stream, e2e_step_block, e_steps)
    bst = M.stats()
    try:
        affinity = sorted(os.sched_getaffinity(0))
        affinity = f"{affinity[0]}-{affinity[-1]} ({len(affinity)} cpus)"
    except Exception:
        affinity = None
    # wall clock is the honest e2e time: the host API blocks on its own H2D copies
    return {"value": ctx.nranks * e_bytes * e_steps / ewall / 1e9, "unit": "GB/s",
            "h2d_bytes_per_step": est["h2d_bytes"] // e_steps, "d2h_bytes_per_step": est["d2h_bytes"] // e_steps,
            "updates_per_rank": Ue, "steps": e_steps,
            "api": "cm_update_slices (one batch call per commit, host pinned payloads) + cm_commit + cm_get",
            "device_ms_per_step": ems / e_steps, "wall_ms_per_step": ewall * 1e3 / e_steps,
            "pinned_h2d_copy_GBps": pcie_gbs, "host_cpu_affinity": affinity,
            "with_local_block_on_host": {"value": ctx.nranks * e_bytes * e_steps / bwall / 1e9, "unit": "GB/s",
                                         "api": "... + cm_local_data_host (get_local_data(): the whole LLD x n_local block copied to the host mirror)",
                                         "d2h_bytes_per_step": bst["d2h_bytes"] // e_steps, "wall_ms_per_step": bwall * 1e3 / e_steps},
            "note": "strict reference semantics: every update's payload is on the device before update() returns"}


def bench_ours(args):
    ctx = Ctx(args)
    main = measure_workload(ctx, args, args.workload, args.steps, args.warmup, main=True)
    configs = None
    want = args.configs
    if want == "auto":
        want = "K2,K3,K4,NS" if (ctx.nranks == 8 and args.workload == "K1") else ""
    if want and want != "none":
        configs = {}
        for wl in want.split(","):
            r = measure_workload(ctx, args, wl, args.config_steps, 3, main=False)
            if r is not None:
                r.pop("_main")
                configs[wl] = r
    if ctx.rank == 0:
        mm = main.pop("_main")
        st, spec, roof = mm["st"], mm["spec"], mm["roof"]
        line = {
            "metric": METRIC, "value": main["value"], "unit": "GB/s", "n_gpus": ctx.nranks, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": main["dtype"], "data": "synthetic",
            "config": {"workload": spec["name"], "updates_per_rank_per_step": spec["U"], "grid": [spec["nprow"], spec["npcol"]],
                       "l2": "inputs (payloads) per step are far larger than L2; no explicit flush",
                       "flush_threshold_elements": mm["flush_threshold"], "sorted_sweep": not args.unsorted,
                       "deterministic": bool(args.deterministic), "apply_kernel": args.apply_kernel, "exchange": mm["exchange"],
                       "payload_bytes_per_rank_per_step": mm["step_bytes"],
                       "apply_policy": (args.apply_policy + " (extension: commits deliver, the step ends with cm_materialize inside the timed region)") if spec["deferred"] else "eager (reference semantics)",
                       "env": {k: v for k, v in sorted(os.environ.items()) if k.startswith("CM_B200_")}},
            "clocks": mm["clocks"], "e2e": mm["e2e"], "gpu_launches": st["kernel_launches"],
            "roofline": roof, "cpu_baseline": mm["cpu"], "parity": main["parity"],
            "commit_latency_us": main["commit_latency_us"],
            "owner_share_of_elements": main["owner_share_of_elements"] if ctx.nranks > 1 else None,
            "wall_ms_per_step": mm["wall"] * 1e3 / args.steps,
            "pack_ms_per_step": main["pack_ms_per_step"], "apply_ms_per_step": main["apply_ms_per_step"],
            "configs": configs,
        }
        print(json.dumps(line), flush=True)
    ctx.close()


def ref_inputs(spec, P, U, Ud):
    """Inputs of the reference arm: the GPU arm's own generator (same seeds, same index sets) when a GPU is
    visible, numpy otherwise. Index sets: all U per rank; payloads: the first Ud per rank (cycled)."""
    try:
        import torch
        if torch.cuda.is_available():
            dev = torch.device("cuda", 0)
            datas, ixs, jxs = [], [], []
            for r in range(P):
                d, i, j = gen_inputs(spec, r, torch, dev, U=U)
                datas.append(np.ascontiguousarray(d[:Ud].cpu().numpy()))
                ixs.append(np.ascontiguousarray(i.cpu().numpy()))
                jxs.append(np.ascontiguousarray(j.cpu().numpy()))
                del d, i, j
            return datas, ixs, jxs, "the GPU arm's generator and seeds (torch Philox on cuda:0)"
    except Exception:
        pass
    dt = np.float64 if spec["dtype"] == 0 else np.float32
    datas, ixs, jxs = [], [], []
    for r in range(P):
        rng = np.random.default_rng(0xC0FFEE + 1000 * r)
        datas.append(rng.uniform(-1, 1, (Ud, spec["n_u"], spec["m_u"])).astype(dt))
        ixs.append(np.stack([np.sort(rng.choice(spec["m"], spec["m_u"], replace=False)) for _ in range(U)]).astype(np.int64))
        jxs.append(np.stack([np.sort(rng.choice(spec["n"], spec["n_u"], replace=False)) for _ in range(U)]).astype(np.int64))
    return datas, ixs, jxs, "numpy generator (no GPU visible)"


def cpu_reference(spec, data, ix, jx, n_updates, reps, per_rank=None, data_period=0):
    """Time the unmodified reference header (oracle/_ref) on host cores. Either per_rank = (datas, ixs, jxs)
    lists for every reference rank, or data/ix/jx of rank 0 (single-rank cpu_baseline leg)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_bindings import RefShim, ref_available
    if not ref_available():
        return {"unavailable": "oracle/_ref/libcm_ref.so not built (needs /root/reference at build time)"}
    P = spec["nprow"] * spec["npcol"]
    Uc = min(spec["U"], n_updates)
    esz = 8 if spec["dtype"] == 0 else 4
    if per_rank is None:
        to_np = lambda t: t.cpu().numpy() if hasattr(t, "cpu") else np.asarray(t)
        assert P == 1
        datas = [np.ascontiguousarray(to_np(data[:Uc]))]
        ixs = [np.ascontiguousarray(to_np(ix[:Uc]))]
        jxs = [np.ascontiguousarray(to_np(jx[:Uc]))]
    else:
        datas, ixs, jxs = per_rank
    best = None
    ref = RefShim(spec["dtype"], spec["m"], spec["n"], spec["nprow"], spec["npcol"], spec["mb"], spec["nb"], spec["lld"])
    try:
        for _ in range(reps):
            t = ref.bench_slices(datas, ixs, jxs, spec["m_u"], spec["n_u"], spec["commit_every"], pin=True, data_period=data_period)
            if best is None or t[0] < best[0]:
                best = t
    finally:
        ref.close()
    gbs = P * Uc * spec["m_u"] * spec["n_u"] * esz / best[0] / 1e9
    return {"value": gbs, "unit": "GB/s", "cores": P, "kind": "reference",
            "sample": f"{Uc} of {spec['U']} updates per rank, {P} rank thread(s) pinned to {P} core(s) of {os.cpu_count()}, "
                      f"reference header unmodified over in-repo UPC++ SMP shim (oracle/_ref), g++ -O3 -DNOCHECK, "
                      f"default bin_flush_threshold=10000"
                      + (f"; payload buffers of the first {data_period} updates cycled (index sets all distinct)" if data_period else ""),
            "seconds": best[0], "update_seconds": best[1], "commit_seconds": best[2]}


def bench_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on host cores, one rank thread
    per grid position (all the host threads the reference can use: it runs one thread per rank)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nranks = args.gpus
    spec = workload_spec(args.workload, nranks)
    esz = 8 if spec["dtype"] == 0 else 4
    P = spec["nprow"] * spec["npcol"]
    Uc = min(spec["U"], args.ref_updates) if args.ref_updates else spec["U"]
    Ud = min(Uc, max(1, (2 << 30) // (P * spec["m_u"] * spec["n_u"] * esz)))  # <= 2 GiB of payload buffers in total
    datas, ixs, jxs, how = ref_inputs(spec, P, Uc, Ud)
    sub = dict(spec, U=Uc)
    period = Ud if Ud < Uc else 0
    for _ in range(max(1, min(args.warmup, 1))):  # one untimed pass warms the allocator and the page cache
        r = cpu_reference(sub, None, None, None, Uc, reps=1, per_rank=(datas, ixs, jxs), data_period=period)
        if "unavailable" in r:
            print(json.dumps({"impl": "reference", "unavailable": r["unavailable"]}))
            return
    secs, last = 0.0, None
    for _ in range(args.steps):
        last = cpu_reference(sub, None, None, None, Uc, reps=1, per_rank=(datas, ixs, jxs), data_period=period)
        secs += last["seconds"]
    value = P * Uc * spec["m_u"] * spec["n_u"] * esz * args.steps / secs / 1e9
    cpu = dict(last, value=value, inputs=how)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": nranks, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": secs * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64" if spec["dtype"] == 0 else "f32", "data": "synthetic",
            "config": {"workload": spec["name"], "updates_per_rank_per_step": Uc, "grid": [spec["nprow"], spec["npcol"]],
                       "note": "the whole workload per step on host cores; one untimed warm-up pass" if Uc == spec["U"] else
                               "each step is a bounded sample of the workload on host cores"},
            "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="K1", choices=["K1", "K2", "K3", "K4", "NS", "K2own", "NSown", "NSown1", "K4own"])
    ap.add_argument("--configs", default="auto",
                    help="other workloads measured in the same invocation and reported under `configs`: "
                         "auto (K2,K3,K4,NS at --gpus 8), none, or a comma-separated list")
    ap.add_argument("--config-steps", type=int, default=3)
    ap.add_argument("--flush-threshold", type=int, default=0, help="elements; 0 = library default")
    ap.add_argument("--unsorted", action="store_true", help="disable the column-sorted sweep")
    ap.add_argument("--deterministic", action="store_true")
    ap.add_argument("--apply-kernel", default="auto", choices=["auto", "red", "tile"])
    ap.add_argument("--exchange", default="auto", choices=["auto", "nccl", "p2p"],
                    help="auto/p2p: the pack kernel stores into the owners' arenas over NVLink; nccl: grouped send/recv")
    ap.add_argument("--apply-policy", default="eager", choices=["eager", "deferred", "background"],
                    help="deferred / background (N > 1, p2p): commits only deliver (background: the logged messages are "
                         "scatter-added on a second stream under the following commits), the step ends with cm_materialize "
                         "(inside the timed region)")
    ap.add_argument("--defer-budget-gib", type=int, default=16)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-updates", type=int, default=0, help="updates per rank per e2e step (0 = all)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--parity-updates", type=int, default=64, help="updates per rank replayed against the reference")
    ap.add_argument("--cpu-updates", type=int, default=10000, help="bounded sample for the cpu_baseline leg")
    ap.add_argument("--ref-updates", type=int, default=0, help="updates per rank per step of --impl reference (0 = the whole workload)")
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
