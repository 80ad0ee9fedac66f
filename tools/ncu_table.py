"""Compact per-launch table from an .ncu-rep (read with `ncu -i`): time, DRAM bytes, issue utilisation, occupancy, the
dominant stall.  Usage: python tools/ncu_table.py rep.ncu-rep [kernel-regex]"""
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
cmd = ["ncu", "-i", sys.argv[1], "--page", "raw", "--csv", "--metrics", ",".join(METRICS)]
if len(sys.argv) > 2:
    cmd += ["--kernel-name", "regex:" + sys.argv[2]]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def conv(v, u):
    try:
        x = float(v)
    except ValueError:
        return 0.0
    u = u.lower()
    if u.startswith("gbyte"):
        return x
    if u.startswith("mbyte"):
        return x / 1e3
    if u.startswith("kbyte"):
        return x / 1e6
    if u == "byte":
        return x / 1e9
    if u in ("ms", "msecond"):
        return x
    if u in ("us", "usecond"):
        return x / 1e3
    if u in ("ns", "nsecond"):
        return x / 1e6
    if u in ("s", "second"):
        return x * 1e3
    return x


print("| kernel | ms | DRAM rd GB | DRAM wr GB | GB/s | issue % | warps % | regs | warp-inst (M) | L2 hit % | long_sb | short_sb | barrier | membar |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows[2:]:
    g = lambda m: conv(r[ix[m]], units[ix[m]])
    name = r[ix["Kernel Name"]].replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    name = name.split("(")[0][-38:]
    ms = g("gpu__time_duration.sum")
    rd, wr = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
    print("| %s | %.3f | %.2f | %.2f | %.0f | %.1f | %.1f | %d | %.0f | %.1f | %.1f | %.1f | %.1f | %.1f |" % (
        name, ms, rd, wr, (rd + wr) / ms * 1e3 if ms else 0, g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        g("sm__warps_active.avg.pct_of_peak_sustained_active"), int(g("launch__registers_per_thread")), g("smsp__inst_executed.sum") / 1e6,
        g("lts__t_sector_hit_rate.pct"), g("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        g("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
        g("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
        g("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio")))
