#!/usr/bin/env python
"""Summarise .ncu-rep captures (ncu -i ... --page raw --csv) into one line of key metrics per launch."""
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "dur_us"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts%"),
        ("sm__warps_active.avg.per_cycle_active", "warps/SM"), ("smsp__issue_active.avg.per_cycle_active", "issue/cyc"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "gld_sectors"), ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "gld_requests"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "gst_sectors"), ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "gst_requests"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "red_sectors"), ("l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "red_requests"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg_throttle"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio_throttle"),
        ("smsp__inst_executed.sum", "warp_insts")]


def to_num(v, unit):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1, "ms": 1e3, "us": 1, "ns": 1e-3, "s": 1e6}.get(unit)
    return x * mult if mult else x


for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        print(path, "no data")
        continue
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {h: (v, u) for h, v, u in zip(hdr, r, units)}
        name = d.get("Kernel Name", ("?", ""))[0]
        print(f"== {path.split('/')[-1]} :: {name[:90]}")
        line = []
        for k, short in KEYS:
            if k in d:
                x = to_num(*d[k])
                line.append(f"{short}={x:.4g}" if isinstance(x, float) else f"{short}={x}")
        print("   " + "  ".join(line))
