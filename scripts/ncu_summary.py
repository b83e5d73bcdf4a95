#!/usr/bin/env python
"""Condense an .ncu-rep (read here, no GPU needed) into the few numbers DESIGN.md / bench.py cite.
usage: ncu_summary.py <file.ncu-rep> [out.md]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fp64.sum",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_xu.sum"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    lines = ["# ncu summary of %s" % rep, ""]
    for r in data:
        lines.append("## %s  grid %s block %s" % (r[name_i][:90], r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append("| %s | %s | %s |" % (k, r[i], units[i]))
        rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        lines.append("")
    text = "\n".join(lines)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")
    else:
        print(text)


if __name__ == "__main__":
    main()
