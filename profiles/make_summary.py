"""Regenerate profiles/r1_ncu_iso4_summary.txt, profiles/traffic.json and the launch list from gpurun_out/ captures:
   r1_ncu_iso4.ncu-rep (ncu --set full ... bench.py --n 100), r1_traffic_n200.csv (dram bytes at n = 200),
   r1_launches_iso4.csv (gpu__time_duration of the timed region).  usage: python profiles/make_summary.py"""
import csv
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
rep = os.path.join(G, "r1_ncu_iso4.ncu-rep")
raw_csv, src_csv = os.path.join(G, "r1_raw_iso4.csv"), os.path.join(G, "r1_src_iso4.csv")
subprocess.run(f"ncu -i {rep} --page raw --csv > {raw_csv} 2>/dev/null", shell=True, check=True)
subprocess.run(f"ncu -i {rep} --page source --csv --print-source cuda,sass > {src_csv} 2>/dev/null", shell=True, check=True)
rows = list(csv.reader(open(raw_csv)))
hdr, units, r = rows[0], rows[1], rows[2]
want = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
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
out = [
    "# k_assemble_iso<Hex8, 1024 threads, 4 per incidence> -- the default kernel for hex8 + isotropic law at the end of round 1",
    "# ncu --set full --clock-control none --import-source on -k regex:k_assemble -s 3 -c 1 python bench.py --n 100 --steps 2 --warmup 3 --no-cpu-baseline",
    "# (hex8 box 100^3 = 1 M elements, 34476 clusters of 32 nodes; per launch, cold cache, serialised: compare shares, not absolutes)",
    "",
    "## key metrics",
]
for h, u, v in zip(hdr, units, r):
    if h in want:
        out.append(f"  {h} [{u}] = {v}")
out += [
    "",
    "## reading: the LSU data pipe (shared-memory wavefronts + global) and the FP64 pipe add up to ~100 % busy -- they do not",
    "## overlap on this part (bench_micro/pipe_overlap.cu, profiles/r1_pipe_overlap.txt): the kernel is bound by the SM's",
    "## shared FP64 / shared-memory issue path, not by HBM (DRAM ~ 11 % of peak).",
    "",
    "## top source lines by stall samples (fedoo_b200/csrc/fdk_assemble_iso.cuh)",
    subprocess.run(["python", os.path.join(ROOT, "profiles", "ncu_lines.py"), src_csv, "30"], capture_output=True, text=True).stdout,
]
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, rr in enumerate(rows) if rr and rr[0] == "Line No")
h2 = rows[hi]
ix = {}
for i, h in enumerate(h2):
    ix.setdefault(h, i)


def num(rr, k):
    try:
        return float(rr[ix[k]])
    except Exception:
        return 0.0


lines = [rr for rr in rows[hi + 1 :] if len(rr) == len(h2) and rr[0].strip().isdigit()]
tot = sum(num(rr, "L1 Wavefronts Shared") for rr in lines)
out.append(f"## shared-memory wavefronts by source line (total {tot / 1e6:.0f} M per 1 M elements)")
for rr in sorted(lines, key=lambda rr: -num(rr, "L1 Wavefronts Shared"))[:18]:
    out.append(f"{rr[0]:>5} {num(rr, 'L1 Wavefronts Shared') / 1e6:8.1f} M (ideal {num(rr, 'L1 Wavefronts Shared Ideal') / 1e6:6.1f} M) | {rr[1].strip()[:105]}")
open(os.path.join(ROOT, "profiles", "r1_ncu_iso4_summary.txt"), "w").write("\n".join(out) + "\n")

rd = wr = None
for rr in csv.reader(open(os.path.join(G, "r1_traffic_n200.csv"))):
    if len(rr) > 14 and rr[12] == "dram__bytes_read.sum":
        rd = int(rr[14])
    if len(rr) > 14 and rr[12] == "dram__bytes_write.sum":
        wr = int(rr[14])


def g(name):
    i = hdr.index(name)
    return float(r[i]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[i]]


t = {
    "n200_g1": rd + wr,
    "n100_g1": int(g("dram__bytes_read.sum") + g("dram__bytes_write.sum")),
    "_source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:k_assemble -s 3 -c 1 python bench.py "
    f"--steps 2 --warmup 3 (k_assemble_iso, one launch at n=200): read {rd} + write {wr} bytes",
}
json.dump(t, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
shutil.copy(os.path.join(G, "r1_launches_iso4.csv"), os.path.join(ROOT, "profiles", "r1_launches_iso4.csv"))
print("\n".join(out[:34]))
print(t)
