#!/usr/bin/env python
"""Summarise an .ncu-rep (captured on the GPU box with `ncu --set full --clock-control none
--import-source on`) into a small text file for profiles/: duration, DRAM bytes, pipe
utilisation, occupancy, stall reasons per issue, shared-memory bank conflicts."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "inst_executed", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_st.sum",
    "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# %s\n# source: %s (ncu --set full --clock-control none --import-source on)\n" % (note, rep))
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write("\nkernel: %s\n" % d.get("Kernel Name", "?"))
            for k in KEYS:
                if k in d:
                    f.write("  %-70s %s %s\n" % (k, d[k], units[hdr.index(k)]))
            f.write("  stall reasons (warps stalled per issue-active cycle):\n")
            for k in hdr:
                if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
                    f.write("    %-28s %s\n" % (k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], d[k]))
    print("wrote", out)


if __name__ == "__main__":
    main()
