"""Multi-process parity check, one process per GPU (launched by torchrun):

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/mp_parity.py

Every rank runs the same seeded scenario through libcm_b200's NCCL transport and compares ITS local
block (and the streams it sent) with the oracle's block for that rank. Exit code 0 = parity.
The pytest wrapper is tests/test_gpu_multi.py (skipped when fewer than 2 GPUs are visible).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.distributed as dist

import cm_b200 as cm
from scenario import Cfg, TOL, normwise_rel, rand_payload, run_checker, slice_steps, sorted_unique

GRIDS = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        t = torch.frombuffer(bytearray(cm.unique_id()), dtype=torch.uint8).to(dev)
    dist.broadcast(t, 0)
    cm.init(rank, world, local, bytes(t.cpu().numpy().tobytes()) if world > 1 else None)
    nprow, npcol = GRIDS[world]
    failures = 0
    for dtype, m, n in [(cm.F64, 2000, 1000), (cm.F32, 1500, 2100), (cm.F64, 2000, 2000)]:
        lld = -(-m // (64 * nprow)) * 64 + 64  # >= m_local on every rank, with padding rows
        cfg = Cfg(dtype, m, n, nprow, npcol, 64, 64, lld)
        rng = np.random.default_rng(1234)
        steps = slice_steps(cfg, rng, 3, 300, 200, commit_every=2)
        # elemental + symmetric + fill_lower on the square matrix
        if m == n:
            for r in range(cfg.P):
                ix = sorted_unique(rng, m, 150)
                steps.append(("sym", r, rand_payload(rng, dtype, (150, 150)), 150, 150, False, ix))
                k = 500
                steps.append(("elems", r, rand_payload(rng, dtype, k), rng.integers(0, m, k), rng.integers(0, n, k)))
            steps += [("commit",), ("fill_lower",), ("commit",)]
        want = run_checker(cfg, steps, wire=True, threshold=1 << 40)
        M = cm.Matrix(*cfg.args())
        if os.environ.get("CM_TEST_EXCHANGE_MODE"):
            M.set_exchange_mode(int(os.environ["CM_TEST_EXCHANGE_MODE"]))
        if os.environ.get("CM_TEST_APPLY_POLICY"):  # 1 = deferred scatter-add (peer-to-peer modes)
            M.set_apply_policy(int(os.environ["CM_TEST_APPLY_POLICY"]), int(os.environ.get("CM_TEST_DEFER_BUDGET", "0")))
        thr = int(os.environ.get("CM_TEST_FLUSH_THRESHOLD", "0"))
        if thr:  # threshold flushes inside update() (reference :346, :614): flush regions over CUDA IPC / NVLink
            M.set_flush_threshold(thr)
        M.keep_wire(True)
        wires = {}
        for st in steps:
            op = st[0]
            if op == "slice" and st[1] == rank:
                _, _, data, m_u, n_u, ld, trans, ix, jx = st
                M.update(data, ix, jx, m_u, n_u, ld, trans)
            elif op == "sym" and st[1] == rank:
                M.update_sym(st[2], st[6], st[4], st[5])
            elif op == "elems" and st[1] == rank:
                M.update_elems(st[2], st[3], st[4])
            elif op == "fill_lower":
                M.fill_lower()
            elif op == "commit":
                M.commit()
                for d in range(cfg.P):
                    inds, data = M.wire(d)
                    if d in wires:
                        wires[d] = (np.concatenate([wires[d][0], inds]), np.concatenate([wires[d][1], data]))
                    else:
                        wires[d] = (inds, data)
        got = M.local()
        err = normwise_rel(got, want["local"][rank])
        ok = err <= TOL[dtype]
        # wire parity only for the slice-only scenarios (our wire keeps slices and elements in separate lanes)
        wire_ok = True
        if m != n:
            for d in range(cfg.P):
                wi, wd = want["wire"][(rank, d)]
                gi, gd = wires.get(d, (np.zeros(0, np.int64), np.zeros(0)))
                wire_ok &= np.array_equal(gi, wi) and gd.tobytes() == wd.tobytes()
        # one-sided remote read (operator(), :585-596) through CUDA IPC
        M.materialize()  # no-op under the eager policy
        read_ok = True
        for _ in range(20):
            i, j = int(rng.integers(0, m)), int(rng.integers(0, n))
            owner, ij = cm.owner_of(i, j, nprow, npcol, 64, 64, lld)
            read_ok &= bool(np.isclose(M.get(i, j), want["local"][owner][ij], rtol=1e-5 if dtype == cm.F32 else 1e-12, atol=1e-6 if dtype == cm.F32 else 1e-12))
        print(f"[rank {rank}] {cfg}: normwise rel err {err:.2e} ok={ok} wire_ok={wire_ok} remote_read_ok={read_ok} "
              f"exchange_mode={M.exchange_mode()} stats={ {k: v for k, v in M.stats().items() if k in ('bytes_sent', 'bytes_received', 'kernel_launches')} }", flush=True)
        failures += (not ok) + (not wire_ok) + (not read_ok)
        if thr:
            st = M.stats()
            shipped_ok = st["flushes"] > 0 and (st["flushed_shipped"] > 0 if M.exchange_mode() == 2 and not os.environ.get("CM_TEST_APPLY_POLICY")
                                                 else st["flushed_held"] > 0)
            print(f"[rank {rank}] threshold {thr}: flushes={st['flushes']} shipped={st['flushed_shipped']} held={st['flushed_held']} ok={shipped_ok}", flush=True)
            failures += not shipped_ok
        M.destroy()
    cm.finalize()
    ft = torch.tensor([failures], device=dev)
    dist.all_reduce(ft)
    dist.destroy_process_group()
    sys.exit(1 if int(ft.item()) else 0)


if __name__ == "__main__":
    main()
