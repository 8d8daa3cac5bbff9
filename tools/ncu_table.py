"""Markdown table of the headline metrics of every kernel in an `ncu --set full` report.

    ncu -i report.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_table.py raw.csv [algorithmic_bytes.json]

The optional JSON maps a kernel-name substring to the algorithmic bytes of the captured launch; the table then carries
DRAM traffic / algorithmic bytes (1.0 = every byte moved once)."""
import csv
import json
import sys

COLS = [
    ("time us", "gpu__time_duration.sum", 1.0),
    ("dram rd MB", "dram__bytes_read.sum", None),
    ("dram wr MB", "dram__bytes_write.sum", None),
    ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0),
    ("fp64 %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 1.0),
    ("occ %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
    ("regs", "launch__registers_per_thread", 1.0),
    ("grid", "launch__grid_size", 1.0),
    ("block", "launch__block_size", 1.0),
    ("warp inst M", "smsp__inst_executed.sum", 1e-6),
    ("long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 1.0),
    ("barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", 1.0),
    ("short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", 1.0),
    ("mio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", 1.0),
    ("math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", 1.0),
]
TO_MB = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
TO_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    alg = json.load(open(sys.argv[2])) if len(sys.argv) > 2 else {}
    head = ["kernel"] + [c[0] for c in COLS] + (["traffic / algorithmic"] if alg else [])
    print("| " + " | ".join(head) + " |")
    print("|" + "---|" * len(head))
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("rb::", "")
        cells, mb = [f"`{name}`"], 0.0
        for label, metric, scale in COLS:
            if metric not in ix or r[ix[metric]] == "":
                cells.append("-")
                continue
            v, u = float(r[ix[metric]].replace(",", "")), units[ix[metric]]
            if metric == "gpu__time_duration.sum":
                v *= TO_US.get(u, 1.0)
            elif scale is None:
                v *= TO_MB.get(u, 1.0)
                mb += v
            else:
                v *= scale
            cells.append(f"{v:.0f}" if label in ("regs", "grid", "block") else f"{v:.4g}")
        if alg:
            hit = [b for k, b in alg.items() if k in name]
            cells.append(f"{mb * 1e6 / hit[0]:.2f}" if hit else "-")
        print("| " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
