"""Condense the raw page of an `ncu --set full` capture into the JSON bench.py reads for roofline.traffic / issue_frac.
    ncu -i R.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_summary.py raw.csv "<how the capture was made>" > profiles/r02_k1w_ncu_summary.json
Per-launch values are the mean over the captured launches of the first kernel name in the file.
"""
import csv
import json
import sys

KEEP = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "smsp__inst_executed.sum": "warp_instructions_per_launch",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "shared_bank_conflicts",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio": "stall_not_selected",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
}
SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3}


def main():
    rows = list(csv.reader(open(sys.argv[1], newline="")))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    name = data[0][kn]
    data = [r for r in data if r[kn] == name]
    out = {"kernel": name.split("(")[0], "launches_captured": len(data), "source": sys.argv[2] if len(sys.argv) > 2 else ""}
    for metric, key in KEEP.items():
        if metric not in hdr:
            continue
        i = hdr.index(metric)
        vals = [float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0) for r in data if r[i]]
        out[key] = sum(vals) / len(vals)
    out["dram_bytes_per_launch"] = out.get("dram_read_bytes", 0) + out.get("dram_write_bytes", 0)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
