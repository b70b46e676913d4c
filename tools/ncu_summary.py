"""Summarise an .ncu-rep (read here, no GPU needed) into a small tracked text file under profiles/.
    python tools/ncu_summary.py gpurun_out/X.ncu-rep profiles/X_summary.txt
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_static",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of {rep} (read with `ncu -i ... --page raw --csv`)\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"\n== {d.get('Kernel Name')}  (launch id {d.get('ID')})\n")
            for k in KEYS:
                if k in d:
                    f.write(f"{k:90s} {d[k]} {units[hdr.index(k)]}\n")
            try:
                tr = float(d["dram__bytes_read.sum"]), float(d["dram__bytes_write.sum"])
                f.write(f"{'traffic (dram read+write, units as above)':90s} {tr[0]} + {tr[1]}\n")
            except Exception:
                pass
    print(open(out).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
