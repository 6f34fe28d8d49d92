"""Summarises an ncu launch list (the `--metrics gpu__time_duration.sum,dram__bytes_read.sum,...` pass of
tools/gpu_call*.sh) into profiles/: per-kernel mean device time, share of all launches, DRAM bytes per launch,
executed warp instructions and issue utilisation; and records the DRAM traffic per launch of every kernel under the
workload's key in profiles/traffic_r2.json (what bench.py reports as roofline.traffic for the same workload).
usage: python tools/summarise_launches.py gpurun_out/X_launches.csv profiles/r2_TAG WORKLOAD_ID [skip_first_n_launches]"""
import collections
import csv
import json
import os
import re
import sys


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"^kbbq::", "", name)
    return re.sub(r"<.*$", "", name)


def main():
    src, out, wid = sys.argv[1:4]
    skip = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    launches = collections.OrderedDict()
    for r in data:
        if len(r) <= vi or int(r[0]) < skip:
            continue
        launches.setdefault((int(r[0]), short(r[ki])), {})[r[mi]] = float(r[vi].replace(",", ""))
    per = collections.OrderedDict()
    for (_, k), m in launches.items():
        per.setdefault(k, []).append(m)
    total = sum(m["gpu__time_duration.sum"] for ms in per.values() for m in ms)
    mean = lambda ms, key: sum(m.get(key, 0.0) for m in ms) / len(ms)
    with open(out + "_kernels.csv", "w") as fh:
        fh.write("kernel,launches,mean_ms,share_of_all_launches_pct,dram_read_GB_per_launch,dram_write_GB_per_launch,"
                 "warp_instructions_G_per_launch,issue_active_pct,l2_sector_hit_pct\n")
        for k, ms in sorted(per.items(), key=lambda kv: -sum(m["gpu__time_duration.sum"] for m in kv[1])):
            t = sum(m["gpu__time_duration.sum"] for m in ms)
            fh.write("%s,%d,%.3f,%.1f,%.3f,%.3f,%.3f,%.1f,%.1f\n" % (
                k, len(ms), t / len(ms) / 1e6, 100 * t / total, mean(ms, "dram__bytes_read.sum") / 1e9,
                mean(ms, "dram__bytes_write.sum") / 1e9, mean(ms, "sm__inst_executed.sum") / 1e9,
                mean(ms, "smsp__issue_active.avg.pct_of_peak_sustained_active"), mean(ms, "lts__t_sector_hit_rate.pct")))
    tpath = os.path.join(os.path.dirname(out), "traffic_r2.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    entry = {k: mean(ms, "dram__bytes_read.sum") + mean(ms, "dram__bytes_write.sum") for k, ms in per.items() if not k.startswith("cub::")}
    if "pass3_scan_kernel" in entry and "pass3_repair_kernel" in entry:
        entry["pass3_scan_kernel+pass3_repair_kernel"] = entry["pass3_scan_kernel"] + entry["pass3_repair_kernel"]
    entry["_source"] = "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the launches of %s (%s)" % (
        os.path.basename(src), os.path.basename(out) + "_kernels.csv")
    traffic[wid] = entry
    json.dump(traffic, open(tpath, "w"), indent=1)
    print(open(out + "_kernels.csv").read())


if __name__ == "__main__":
    main()
