"""CPU test of the N > 1 host logic with world_size-2 gloo: each rank plans the owner
segments of ITS updates with the product's pure-integer host helpers (cm_owner_of,
cm_numroc), the per-owner counts are exchanged with an all-to-all (what commit() does with
NCCL), and every rank checks what it would receive against the oracle's wire streams."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cm_b200 as cm
    from scenario import Cfg, run_checker, slice_steps
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = Cfg(cm.F64, 700, 500, 1, 2, 64, 64, 768)
        rng = np.random.default_rng(77)
        steps = slice_steps(cfg, rng, 3, 90, 70)  # same seeded scenario on both ranks
        want = run_checker(cfg, steps, wire=True, threshold=1 << 40)
        # plan: elements this rank sends to each owner (owner map, bit-exact)
        send = np.zeros(world, dtype=np.int64)
        first_ij = {}
        for st in steps:
            if st[0] != "slice" or st[1] != rank:
                continue
            _, _, data, m_u, n_u, ld, trans, ix, jx = st
            for j in jx:
                for i in ix:
                    owner, ij = cm.owner_of(int(i), int(j), cfg.nprow, cfg.npcol, cfg.mb, cfg.nb, cfg.lld)
                    first_ij.setdefault(owner, ij)
                    send[owner] += 1
        recv = torch.zeros(world, dtype=torch.int64)
        dist.all_to_all_single(recv, torch.from_numpy(send))
        for src in range(world):
            assert int(recv[src]) == len(want["wire"][(src, rank)][0]), (src, rank)
        for owner, ij in first_ij.items():
            assert want["wire"][(rank, owner)][0][0] == ij
        # sizes of the receive side follow numroc
        assert want["m_local"][rank] == cm.numroc(cfg.m, cfg.nprow, rank // cfg.npcol, cfg.mb)
        assert want["n_local"][rank] == cm.numroc(cfg.n, cfg.npcol, rank % cfg.npcol, cfg.nb)
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert float(t) == float(world)
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


def test_exchange_plan_world_size_2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert dict(ret) == {0: 1, 1: 1}


def test_bench_workloads_are_the_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench
    k1 = bench.workload_spec("K1", 1)
    assert (k1["m"], k1["n"], k1["nprow"], k1["npcol"], k1["mb"], k1["nb"], k1["lld"]) == (8192, 8192, 1, 1, 64, 64, 8192)
    assert (k1["m_u"], k1["n_u"], k1["U"], k1["dtype"]) == (256, 256, 10000, 0)
    k2 = bench.workload_spec("K2", 8)
    assert (k2["m"], k2["nprow"], k2["npcol"], k2["lld"], k2["m_u"], k2["commit_every"]) == (32768, 2, 4, 16384, 512, 256)
    k4 = bench.workload_spec("K4", 8)
    assert (k4["dtype"], k4["nprow"], k4["npcol"], k4["m_u"], k4["lld"]) == (1, 4, 2, 32, 8192)


def test_bench_line_contract_on_the_recorded_lines():
    """the JSON lines bench.py printed on B200 this round (profiles/r2/) carry every key of the bench contract,
    and the reference arm's workload naming matches the GPU arm's"""
    import glob
    import json
    sys.path.insert(0, ROOT)
    import bench
    need = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "parity"}
    lines = sorted(glob.glob(os.path.join(ROOT, "profiles", "r2", "bench_*.json")))
    assert lines
    for path in lines:
        b = json.load(open(path))
        assert need <= set(b), (path, need - set(b))
        assert b["metric"] == bench.METRIC and b["unit"] == "GB/s" and b["higher_is_better"] is True and b["scaling"] == "weak"
        assert b["gpu_launches"] > 0 and "workload" in b["config"]
        r = b["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        p = b["parity"]
        assert p["checked"] and p["ok"] and p["wire_bit_exact"] and p["deterministic_bit_exact_run_to_run"]
        if b["e2e"] is not None:
            assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(b["e2e"]) and b["e2e"]["h2d_bytes_per_step"] > 0
        if b["n_gpus"] == 1:
            c = b["cpu_baseline"]
            assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] == "reference"
        assert b["config"]["workload"] == bench.workload_spec(path.split("bench_")[1].split("_")[0].replace("k1weak", "K1").upper().replace("K1WEAK", "K1"), b["n_gpus"])["name"]
