#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, with no GPU) into the few numbers DESIGN.md / bench.py quote.

    python tools/ncu_summary.py gpurun_out/r01_part1_mode0.ncu-rep [more.ncu-rep ...] > profiles/xxx.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

RAW_KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg.per_second",
]


def ncu(path, page):
    out = subprocess.run(["ncu", "-i", path, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    for path in sys.argv[1:]:
        rows = ncu(path, "raw")
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = dict(zip(hdr, zip(units, vals)))
            print(f"== {path}")
            print(f"kernel: {d.get('Kernel Name', ('', '?'))[1]}")
            for k in RAW_KEYS:
                if k in d:
                    print(f"  {k:70s} {d[k][1]:>16s} {d[k][0]}")
            stalls = {k: float(v[1]) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")}
            top = sorted(stalls.items(), key=lambda kv: -kv[1])[:6]
            print("  top stall reasons (warps stalled per issue-active cycle): " +
                  ", ".join(f"{k.split('stalled_')[1].split('_per_issue')[0]}={v:.2f}" for k, v in top))
        src = ncu(path, "source")
        if len(src) > 2:
            hdr = src[1]
            try:
                ia, ie = hdr.index("Source"), hdr.index("Instructions Executed")
            except ValueError:
                continue
            ops = collections.Counter()
            for r in src[2:]:
                try:
                    n = int(r[ie])
                except (ValueError, IndexError):
                    continue
                m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia])
                ops[m.group(2).split(".")[0] if m else "?"] += n
            tot = sum(ops.values())
            print(f"  executed warp-instructions: {tot}; mix: " + ", ".join(f"{o} {100*n/tot:.1f}%" for o, n in ops.most_common(10)))
        print()


if __name__ == "__main__":
    main()
