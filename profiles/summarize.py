"""Turns ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches.csv  profiles/launches_r01.md
    python profiles/summarize.py kernel   gpurun_out/prof.ncu-rep  profiles/<name>_r01.md  [traffic-key]
"""
import collections
import csv
import json
import os
import subprocess
import sys


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    dram = collections.defaultdict(float)  # DRAM bytes (read + write) per kernel, when the capture has them
    for x in csv.DictReader(lines):
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"]
        name = x.get("Metric Name", "gpu__time_duration.sum")
        if name.startswith("dram__bytes"):
            dram[x["Kernel Name"]] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            continue
        if name != "gpu__time_duration.sum":
            continue
        ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
        agg.setdefault(x["Kernel Name"], []).append(ms)
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({os.path.basename(src)})\n\n"
                "`ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --clock-control none` over "
                "`bench.py`; times are cold-cache and serialised (compare shares, not absolutes).\n\n"
                "| kernel | launches | avg ms | total ms | share | avg DRAM GB |\n|---|---:|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            gb = f"{dram[k] / len(v) / 1e9:.3f}" if k in dram else ""
            f.write(f"| `{k[:110]}` | {len(v)} | {sum(v) / len(v):.4f} | {sum(v):.3f} | {100 * sum(v) / tot:.1f}% | {gb} |\n")
    print("wrote", dst)


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def kernel(src, dst, traffic_key=None):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({os.path.basename(src)})\n\n")
        for r in data:
            f.write(f"## `{r[name_i][:120]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            vals = {}
            for i, h in enumerate(hdr):
                if h in KEYS:
                    f.write(f"| {h} | {r[i]} | {units[i]} |\n")
                    vals[h] = (r[i], units[i])
            f.write("\n")
            if traffic_key:
                def gb(k):
                    v, u = vals[k]
                    v = float(v.replace(",", ""))
                    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
                t = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
                tp = os.path.join(os.path.dirname(dst), "traffic_r01.json")
                d = json.load(open(tp)) if os.path.exists(tp) else {}
                d[traffic_key] = t
                json.dump(d, open(tp, "w"), indent=1)
                f.write(f"DRAM traffic per launch (read + write): {t / 1e9:.3f} GB\n\n")
                traffic_key = None
    print("wrote", dst)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        kernel(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
