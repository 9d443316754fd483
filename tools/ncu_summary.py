#!/usr/bin/env python3
"""Summarise an .ncu-rep (one block per profiled launch): duration, DRAM bytes, issue utilisation, top stall reasons."""
import csv
import subprocess
import sys


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def g(r, name, default="?"):
        return r[col[name]] if name in col else default
    for r in rows[2:]:
        name = g(r, "Kernel Name").split("(")[0]
        print("== %s  grid %s block %s regs %s" % (name, g(r, "launch__grid_size"), g(r, "launch__block_size"), g(r, "launch__registers_per_thread")))
        print("   time %s %s | dram rd %s %s wr %s %s | dram %% %s | sm %% %s | warps_active %% %s | issue_active %% %s | eligible/cyc %s | inst %s" % (
            g(r, "gpu__time_duration.sum"), units[col["gpu__time_duration.sum"]],
            g(r, "dram__bytes_read.sum"), units[col["dram__bytes_read.sum"]], g(r, "dram__bytes_write.sum"), units[col["dram__bytes_write.sum"]],
            g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), g(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
            g(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            g(r, "smsp__warps_eligible.avg.per_cycle_active"), g(r, "smsp__inst_executed.sum")))
        stalls = []
        for h, i in col.items():
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        print("   stalls/issue: " + ", ".join("%s %.2f" % (n, v) for v, n in sorted(stalls, reverse=True)[:6]))
        for extra in ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
                      "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "sass__inst_executed_local_loads"):
            if extra in col:
                print("   %s = %s" % (extra, r[col[extra]]))


if __name__ == "__main__":
    main(sys.argv[1])
