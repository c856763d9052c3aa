"""Summarise an .ncu-rep (read here, on the CPU box, with `ncu -i ... --page raw --csv`) into the text kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "command line that was profiled" > profiles/rNN_ncu_xxx.txt
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__cluster_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg.per_second",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    note = sys.argv[2] if len(sys.argv) > 2 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(note)
    for r in data:
        name = r[col["Kernel Name"]]
        print("\nkernel  %s   (launch id %s)" % (name, r[col["ID"]]))
        for m in METRICS:
            if m in col and r[col[m]] != "":
                print("  %-72s %s %s" % (m, r[col[m]], units[col[m]]))


if __name__ == "__main__":
    main()
