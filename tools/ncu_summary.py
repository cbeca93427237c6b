#!/usr/bin/env python
"""profiles/*.json from an .ncu-rep: the handful of metrics DESIGN.md quotes for a kernel.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.json "what was captured"
"""
import csv
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__cycles_active.avg",
    "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def main():
    rep, dst, what = sys.argv[1], sys.argv[2], sys.argv[3]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        rec = dict(zip(hdr, vals))
        m = {}
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS:
                m[h] = {"value": v, "unit": u}
        res.append({"kernel": rec.get("Kernel Name", ""), "metrics": m})
    with open(dst, "w") as f:
        json.dump({"what": what, "launches": res}, f, indent=1)
    print(dst, len(res), "launch(es)")


if __name__ == "__main__":
    main()
