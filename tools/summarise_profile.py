"""Turns gpurun_out/TAG_full.ncu-rep and TAG_launches.csv into the tracked summaries under profiles/:
  profiles/TAG_kernels.csv    one row per kernel: time, instructions, issue utilisation, stall mix, L1/L2 hit rates, DRAM bytes
  profiles/TAG_launches.csv   the launch list (copied), and each kernel's share of a step from it
  profiles/traffic.json       dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed by kernel (bench.py reads it)
usage: python tools/summarise_profile.py TAG"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rep = os.path.join(ROOT, "gpurun_out", tag + "_full.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = [
    ("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "cycles_per_issue"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_inst"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall_branch"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads_per_inst"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1_throughput_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_throughput_pct"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"),
    ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
]
idx = [(hdr.index(m), n) for m, n in want if m in hdr]
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
traffic = {}
with open(os.path.join(ROOT, "profiles", tag + "_kernels.csv"), "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow([n + (" [" + units[i] + "]" if units[i] else "") for i, n in idx])
    for r in rows[2:]:
        w.writerow([r[i].split("(")[0].replace("void ", "") if n == "kernel" else r[i] for i, n in idx])
        name = re.sub(r"<.*", "", r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("kbbq::", ""))
        rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        traffic[name] = float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]]
json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
src = os.path.join(ROOT, "gpurun_out", tag + "_launches.csv")
if os.path.exists(src):
    dst = os.path.join(ROOT, "profiles", tag + "_launches.csv")
    shutil.copy(src, dst)
    per = collections.defaultdict(list)
    for r in csv.DictReader(l for l in open(src) if l.startswith('"')):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            per[r["Kernel Name"].split("(")[0]].append(float(r["Metric Value"]))
    tot = sum(sum(v) for v in per.values())
    with open(os.path.join(ROOT, "profiles", tag + "_launch_shares.txt"), "w") as fh:
        for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            fh.write("%-28s launches %3d  mean %9.1f us  share of all launches %5.1f %%\n" % (k, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
print(open(os.path.join(ROOT, "profiles", tag + "_kernels.csv")).read())
