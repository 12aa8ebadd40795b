#!/usr/bin/env python
"""Table of a scripts/sweep.py run: per workload, variants sorted by step time with the ratio to `batch`.

    python scripts/summarize_sweep.py gpurun_out/variants_n8.log [more logs ...]
"""
import json
import sys
from collections import defaultdict


def main(paths):
    rows = defaultdict(list)
    for p in paths:
        for line in open(p):
            line = line.strip()
            if not line.startswith("{"):
                continue
            try:
                r = json.loads(line)
            except ValueError:
                continue
            if "variant" in r:
                rows[(r["workload"], r["n_gpus"])].append(r)
    for (wl, n), rs in sorted(rows.items()):
        base = next((r["ms_per_step"] for r in rs if r["variant"] == "batch"), None)
        print(f"\n{wl} on {n} GPU(s)" + (f" — batch {base:.3f} ms/step" if base else ""))
        print(f"  {'variant':34s} {'ms/step':>9s} {'vs batch':>9s} {'GB/s':>9s} {'pack ms':>8s} {'apply ms':>9s} {'commit us (empty / one)':>24s}")
        for r in sorted(rs, key=lambda r: r["ms_per_step"]):
            ratio = f"{base / r['ms_per_step']:.3f}x" if base else "-"
            print(f"  {r['variant']:34s} {r['ms_per_step']:9.3f} {ratio:>9s} {r['GBps']:9.1f} {r['pack_ms_per_step']:8.3f} "
                  f"{r['apply_ms_per_step']:9.3f} {r['commit_us_empty_p50']:11.1f} / {r['commit_us_one_update_p50']:<10.1f}")


if __name__ == "__main__":
    main(sys.argv[1:] or ["/dev/stdin"])
