"""Generates tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libcm_ref.so = the
unmodified reference header over the in-repo UPC++ shim). Run in the build container, where
/root/reference exists:

    make -C oracle && python tests/golden/make_golden.py

The reference ships no golden vectors (its tests are randomised, test/cm_test.cpp), so these
fixtures are what pins the oracle and the CUDA path on machines without /root/reference.
Each fixture stores the seeded scenario's parameters and, per rank, the reference's local block
and the (inds, data) streams it put on the wire (rput hook, reference :168-171).
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from golden_scenarios import SCENARIOS, build_steps  # noqa: E402
from scenario import Cfg, run_checker  # noqa: E402


def main():
    only = set(sys.argv[1:])  # fixture names to (re)generate; none = all (the reference applies messages in
    for name, spec in SCENARIOS.items():  # arrival order, so regenerated blocks may differ in the last bit)
        if only and name not in only:
            continue
        cfg = Cfg(*spec["cfg"])
        steps = build_steps(spec)
        out = run_checker(cfg, steps, kind="ref", wire=True, threshold=spec.get("threshold"))
        arrays = {"cfg": np.array(spec["cfg"], dtype=np.int64)}
        for r in range(cfg.P):
            arrays[f"local_{r}"] = out["local"][r]
            arrays[f"m_local_{r}"] = np.int64(out["m_local"][r])
            arrays[f"n_local_{r}"] = np.int64(out["n_local"][r])
        if spec.get("threshold") is not None:  # wire streams only where they are one message per commit
            for (s, d), (inds, data) in out["wire"].items():
                arrays[f"wire_inds_{s}_{d}"] = inds.astype(np.int32)  # all fixtures have ij < 2^31
                # the values are random payload bytes: keep their SHA-256, not 8 bytes per element
                arrays[f"wire_sha_{s}_{d}"] = np.frombuffer(hashlib.sha256(data.tobytes()).digest(), dtype=np.uint8)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(f"{name}: {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
