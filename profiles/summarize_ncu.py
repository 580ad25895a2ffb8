#!/usr/bin/env python
"""Extract the judged numbers from an .ncu-rep (ncu --set full) into a small text summary.

    python profiles/summarize_ncu.py gpurun_out/prof_backward.ncu-rep > profiles/r01_backward_ncu_summary.txt
"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__sass_inst_executed_op_global_ld.sum",
        "smsp__sass_inst_executed_op_global_st.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_col = hdr.index("Kernel Name")
    print("# source: %s (%d profiled launches)" % (path, len(data)))
    for r in data:
        print("\n## %s" % r[name_col][:100])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("%-90s %s %s" % (w, r[i], units[i]))
        for i, h in enumerate(hdr):  # every other stall reason that matters
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h and h not in WANT:
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 0.15:
                    print("%-90s %s %s" % (h, r[i], units[i]))


if __name__ == "__main__":
    main(sys.argv[1])
