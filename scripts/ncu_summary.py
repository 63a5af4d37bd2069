"""Condense an .ncu-rep (ncu --set full) into the handful of numbers the roofline argument needs.
usage: python scripts/ncu_summary.py profiles/x.ncu-rep > profiles/x.summary.txt"""
import csv, subprocess, sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), CTAs/SM"),
    ("launch__waves_per_multiprocessor", "waves / SM"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle / issue"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("kernel:", r[col["Kernel Name"]][:110])
        print("  grid", r[col["Grid Size"]], "block", r[col["Block Size"]])
        for key, label in KEYS:
            if key in col and r[col[key]] not in ("", "n/a"):
                print(f"  {label:42s} {r[col[key]]:>16s} {units[col[key]]}")
        try:
            rd = float(r[col["dram__bytes_read.sum"]]); wr = float(r[col["dram__bytes_write.sum"]])
            u = units[col["dram__bytes_read.sum"]]
            print(f"  {'DRAM traffic (read + write)':42s} {rd + wr:16.1f} {u}")
        except Exception:
            pass
        print()


if __name__ == "__main__":
    main(sys.argv[1])
