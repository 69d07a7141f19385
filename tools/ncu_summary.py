"""Text summary of an `ncu --set full` report (read HERE, no GPU needed): one column per captured launch,
the metrics the roofline discussion uses.  profiles/*_ncu_full_summary.txt are made with it.

  python tools/ncu_summary.py gpurun_out/x.ncu-rep ["header line"] > profiles/rNN_x_ncu_full_summary.txt
MEASUREMENT INFRASTRUCTURE."""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__block_size", "launch__grid_size", "smsp__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sectors_srcunit_tex_op_write.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "launch__local_size", "smsp__inst_executed_op_shfl.sum", "sm__sass_inst_executed_op_global_ld.sum",
        "sm__sass_inst_executed_op_global_st.sum", "sm__sass_inst_executed_op_local_ld.sum", "sm__sass_inst_executed_op_local_st.sum"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    if len(sys.argv) > 2:
        print(sys.argv[2])
        print()
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w} [{units[i]}]: " + " | ".join(r[i] for r in data))


if __name__ == "__main__":
    main()
