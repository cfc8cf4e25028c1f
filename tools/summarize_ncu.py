#!/usr/bin/env python3
"""Turn an .ncu-rep (read here with `ncu -i`, no GPU needed) into a small markdown summary for profiles/.

    python tools/summarize_ncu.py gpurun_out/prof_scan_r1.ncu-rep profiles/r1_scan_kernel.md ["title"]
    python tools/summarize_ncu.py --launches gpurun_out/launches_r1.csv profiles/r1_launches.md
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict, defaultdict

KEYS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "sm__maximum_warps_per_active_cycle_pct",
    "smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header, units = rows[0], rows[1]
    return header, units, rows[2:]


def summarize(rep, dst, title):
    header, units, rows = raw_rows(rep)
    idx = {h: i for i, h in enumerate(header)}
    lines = [f"# {title}", "", f"Source: `{rep}` (`ncu --set full --clock-control none --import-source on`), read with `ncu -i ... --page raw --csv`.",
             "Numbers under a profiler are for attribution (replayed, serialised launches), never bench values.", ""]
    for r in rows:
        name = r[idx["Kernel Name"]]
        lines.append(f"## launch id {r[idx['ID']]}: `{name[:160]}`")
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        for k in KEYS:
            if k in idx and r[idx[k]] != "":
                lines.append(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |")
        try:
            t = float(r[idx["gpu__time_duration.sum"]].replace(",", ""))
            tu = units[idx["gpu__time_duration.sum"]]
            scale = {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}.get(tu, 1e-9)
            rd = float(r[idx["dram__bytes_read.sum"]].replace(",", ""))
            wr = float(r[idx["dram__bytes_write.sum"]].replace(",", ""))
            bu = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[idx["dram__bytes_read.sum"]], 1.0)
            wu = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[idx["dram__bytes_write.sum"]], 1.0)
            traffic = rd * bu + wr * wu
            lines.append(f"| **derived: DRAM traffic** | {traffic / 1e6:.1f} | MB |")
            lines.append(f"| **derived: DRAM GB/s over the launch** | {traffic / (t * scale) / 1e9:.0f} | GB/s |")
        except Exception:
            pass
        lines.append("")
    open(dst, "w").write("\n".join(lines) + "\n")
    print(f"wrote {dst} ({len(rows)} launches)")


def launches(csv_path, dst):
    text = open(csv_path).read()
    start = text.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = defaultdict(lambda: [0, 0.0])
    order = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        short = name.split("(")[0][:110]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(unit, 1)
        agg[short][0] += 1
        agg[short][1] += ns
        order.setdefault(short, None)
    total = sum(v[1] for v in agg.values())
    lines = ["# kernel launch list (ncu --metrics gpu__time_duration.sum --clock-control none)", "",
             f"Source: `{csv_path}`.  Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.", "",
             "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k in sorted(agg, key=lambda x: -agg[x][1]):
        lines.append(f"| `{k}` | {agg[k][0]} | {agg[k][1] / 1e6:.3f} | {100 * agg[k][1] / total:.1f} % |")
    open(dst, "w").write("\n".join(lines) + "\n")
    print(f"wrote {dst}")


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        summarize(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else sys.argv[1])
