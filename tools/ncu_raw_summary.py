"""Selected metrics per launch from an `ncu -i X.ncu-rep --page raw --csv` export (the summaries under profiles/r2_ncu_full_*).

    python tools/ncu_raw_summary.py raw.csv [extra metric substrings ...] > profiles/....txt
"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__lsuin_requests.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum"]


def main():
    path, extra = sys.argv[1], sys.argv[2:]
    rows = list(csv.reader(open(path, newline="")))
    head = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[head], rows[head + 1]
    col = {n: i for i, n in enumerate(names)}
    picked = [n for n in WANT if n in col] + [n for n in names if any(e in n for e in extra) and n not in WANT]
    stall = [n for n in names if n.startswith("smsp__average_warp") and "issue_stalled" in n and n.endswith("_per_warp_active.pct")]
    for r in rows[head + 2:]:
        if len(r) < len(names):
            continue
        print("Kernel Name", r[col["Kernel Name"]])
        for n in picked:
            print(n, r[col[n]], units[col[n]])
        top = sorted(((float(r[col[n]] or 0), n) for n in stall), reverse=True)[:6]
        for v, n in top:
            print(f"stall {n.split('issue_stalled_')[1].split('_per_warp')[0]} {v:.1f} %")
        print("---")


if __name__ == "__main__":
    main()
