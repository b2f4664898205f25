#!/usr/bin/env python
"""Turns the raw ncu output of tools/capture_profiles.sh (gpurun_out/<tag>_*) into the summaries committed under
profiles/: <tag>_launches.csv (copied), <tag>_launches_summary.txt, <tag>_scan_ncu_summary.json, <tag>_bench.json."""
import collections
import csv
import json
import shutil
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
src, dst = "gpurun_out/%s_" % tag, "profiles/%s_" % tag
VARIANT = {"0, 0": "ts=RLE,val=simple8b", "0, 1": "ts=RLE,val=gorilla", "1, 0": "ts=simple8b,val=simple8b",
           "1, 1": "ts=simple8b,val=gorilla"}

# ---- launch list ----------------------------------------------------------------------------------
rows = [r for r in csv.reader(open(src + "launches.csv")) if len(r) > 8]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows:
    if r is hdr or r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    name = r[ki].split("(")[0][:70]
    n, t = agg.get(name, (0, 0.0))
    agg[name] = (n + 1, t + float(r[vi].replace(",", "")) / 1e3)
total = sum(t for _, t in agg.values())
with open(dst + "launches_summary.txt", "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 3 --no-cpu-baseline\n")
    f.write("(cold-cache, serialised launch times: compare SHARES; k_gather_pages / k_verify_crc belong to the e2e path only)\n\n")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%-72s n=%4d %11.1f us %6.1f%%\n" % (name, n, t, 100 * t / total))
shutil.copy(src + "launches.csv", dst + "launches.csv")
shutil.copy(src + "bench.json", dst + "bench.json")

# ---- --set full capture of the fused kernels -------------------------------------------------------
raw = list(csv.reader(open(src + "scan_raw.csv")))
h = raw[0]
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
units = dict(zip(h, raw[1]))
kernels = []
for r in raw[2:]:
    d = dict(zip(h, r))
    name = d["Kernel Name"].replace("tskv::", "").replace("void ", "").split("(")[0]
    k = {"kernel": name, "variant": VARIANT.get(name[name.find("<") + 1:name.find("<") + 5], "")}
    for m in KEEP:
        if m not in d or d[m] == "":
            continue
        v = float(d[m].replace(",", ""))
        u = units.get(m, "")
        if m == "gpu__time_duration.sum":
            v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        if m.startswith("dram__bytes"):
            v = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6) * v
        k[m] = v
    kernels.append(k)
json.dump({"source": "ncu --set full --clock-control none --import-source on, tools/capture_profiles.sh %s (the fused kernels of "
                     "one step of the C4 workload; ncu serialises them)" % tag,
           "units": {"gpu__time_duration.sum": "us", "dram__bytes_*": "MB"}, "kernels": kernels,
           "totals": {"dram_MB": sum(k.get("dram__bytes_read.sum", 0) + k.get("dram__bytes_write.sum", 0) for k in kernels),
                      "warp_instructions": sum(k.get("smsp__inst_executed.sum", 0) for k in kernels)},
           # read by bench.py as roofline.traffic (per fused phase = per "launch" of the roofline object)
           "step_dram_traffic_bytes": int(round(1e6 * sum(k.get("dram__bytes_read.sum", 0) + k.get("dram__bytes_write.sum", 0) for k in kernels))),
           "step_dram_traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum summed over the fused kernels of one C4 step"},
          open(dst + "scan_ncu_summary.json", "w"), indent=1)
print(open(dst + "launches_summary.txt").read())
for k in kernels:
    print(k["kernel"], k["variant"], "%.0f us" % k["gpu__time_duration.sum"], "regs %d grid %d" % (k["launch__registers_per_thread"], k["launch__grid_size"]),
          "inst %.3g issue %.1f%% dram_rd %.0f MB" % (k["smsp__inst_executed.sum"], k["smsp__issue_active.avg.pct_of_peak_sustained_active"], k["dram__bytes_read.sum"]))
