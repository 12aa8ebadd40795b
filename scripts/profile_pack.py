#!/usr/bin/env python
"""k_pack under ncu on ONE GPU: 8 ranks as threads of this process (loopback transport), a 2x4 grid with
K1-weak geometry, a few hundred 256x256 updates per rank. The pack kernel's stores land in the owners'
arenas on the same device, so this shows the gather / store coalescing of the kernel, not NVLink."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import cm_b200 as cm  # noqa: E402

P, U, m_u, n_u = 8, 512, 256, 256
world = cm.LoopbackWorld(P, 0)
try:
    mats = world.run(lambda r: cm.Matrix(cm.F64, 16384, 32768, 2, 4, 64, 64, 8192))
    g = torch.Generator(device="cuda")
    g.manual_seed(11)
    keep = []
    for r in range(P):
        ix = torch.rand((U, 16384), generator=g, device="cuda").topk(m_u, dim=1).indices.sort(dim=1).values.contiguous()
        jx = torch.rand((U, 32768), generator=g, device="cuda").topk(n_u, dim=1).indices.sort(dim=1).values.contiguous()
        data = (torch.rand((U, n_u, m_u), generator=g, device="cuda", dtype=torch.float64) * 2 - 1).contiguous()
        keep.append((cm.SliceBatch(U, m_u, n_u, data.data_ptr(), ix.data_ptr(), jx.data_ptr(), 8), ix, jx, data))
    torch.cuda.synchronize()

    def step(r):
        mats[r].update_batch(keep[r][0], True)
        mats[r].commit()
    for _ in range(3):
        world.run(step)
    print("ok", world.run(lambda r: mats[r].stats()["kernel_launches"])[0])
    world.run(lambda r: mats[r].destroy())
finally:
    world.close()
