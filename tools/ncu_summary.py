"""Key figures of every kernel in an ncu report (`ncu --set full`), as text for profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_ncu_summary.txt
"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM limit (smem)"), ("launch__occupancy_limit_registers", "CTAs/SM limit (regs)"),
    ("launch__occupancy_limit_warps", "CTAs/SM limit (warps)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC (per SM, elapsed)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier (warps/issue)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio throttle"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not selected"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    col = {h: i for i, h in enumerate(H)}
    print(f"# {rep}: ncu --set full --clock-control none (cold-cache, serialised launches: compare shares, not absolutes)")
    for r in rows[2:]:
        print(f"\n== {r[col['Kernel Name']]}  (launch id {r[col['ID']]})")
        for key, label in WANT:
            if key in col:
                print(f"  {label:32s} {r[col[key]]} {U[col[key]]}")


if __name__ == "__main__":
    main()
