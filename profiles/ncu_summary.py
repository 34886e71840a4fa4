"""Summarise one `ncu --set full --import-source on` report: key metrics + the top source lines by stall samples.
usage: python profiles/ncu_summary.py REPORT.ncu-rep OUT.txt "header line" ["header line" ...]"""
import csv
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
rep, out = sys.argv[1], sys.argv[2]
raw, src = rep + ".raw.csv", rep + ".src.csv"
subprocess.run(f"ncu -i {rep} --page raw --csv > {raw} 2>/dev/null", shell=True, check=True)
subprocess.run(f"ncu -i {rep} --page source --csv --print-source cuda,sass > {src} 2>/dev/null", shell=True, check=True)
rows = list(csv.reader(open(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
want = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__block_size", "launch__grid_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__inst_executed.sum",
]  # fmt: skip
lines = ["# " + h for h in sys.argv[3:]] + ["", "## key metrics", "== " + (r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")]
for h, u, v in zip(hdr, units, r):
    if h in want:
        lines.append(f"  {h} [{u}] = {v}")
lines += ["", "## top source lines by stall samples"]
lines += subprocess.run([sys.executable, os.path.join(HERE, "ncu_lines.py"), src, "22"], capture_output=True, text=True).stdout.splitlines()
open(out, "w").write("\n".join(lines) + "\n")
os.remove(raw)
os.remove(src)
print("\n".join(lines[:45]))
