"""Summarise an ncu report for profiles/: headline metrics of every captured kernel + the hottest source lines.

usage: ncu_summary.py <report.ncu-rep> <lib.so> <kernel-substring> <out.md>
"""
import csv, io, subprocess, sys, os

rep, lib, kname, outp = sys.argv[1:5]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
with open(outp, "w") as f:
    f.write("# %s\n\nFrom `%s` (`ncu --set full --clock-control none --import-source on`).\n\n" % (os.path.basename(outp), os.path.basename(rep)))
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        f.write("## %s\n\n| metric | value | unit |\n|---|---|---|\n" % d.get("Kernel Name", "?"))
        for h, u in zip(hdr, units):
            if h in WANT and d.get(h, "") != "":
                f.write("| %s | %s | %s |\n" % (h, d[h], u))
        try:
            rd = float(d["dram__bytes_read.sum"].replace(",", "")); wr = float(d["dram__bytes_write.sum"].replace(",", ""))
            f.write("\n(dram read + write per launch as reported above; units differ per column)\n")
        except Exception:
            pass
        f.write("\n")
    lines = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_lines.py"), rep, lib, kname, "30"],
                           capture_output=True, text=True).stdout
    f.write("## hottest source lines of `%s` (executed warp-instructions %%, stall samples %%)\n\n```\n%s```\n" % (kname, lines))
print("wrote", outp)
