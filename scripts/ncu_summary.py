#!/usr/bin/env python
"""Summarises an `ncu --set full` report for profiles/: the metrics DESIGN.md and bench.py's roofline quote, per
profiled launch, in the layout of profiles/r01_ncu_find_kernel_*.txt; optionally records the DRAM bytes of the first
launch of a kernel in profiles/traffic_find_cfg2.json (what bench.py reports as roofline.traffic).

  python scripts/ncu_summary.py gpurun_out/r02_prof_find.ncu-rep --note "cfg2, fused 16-mer table" \
         --out profiles/r02_ncu_find_kernel.txt [--traffic-key k16f --kernel-regex find_kernel]

Runs here (no GPU needed): `ncu -i report --page raw --csv` is parsed.
"""
import argparse
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def read_raw(report):
    """[{column: value}] per profiled launch, plus the units row of the raw page."""
    text = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    start = text.find('"ID"')
    rows = list(csv.reader(io.StringIO(text[start:])))
    header, units, launches = rows[0], rows[1], rows[2:]
    return header, dict(zip(header, units)), [dict(zip(header, r)) for r in launches if len(r) == len(header)]


def number(value):
    try:
        return float(value.replace(",", ""))
    except ValueError:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--note", default="")
    ap.add_argument("--out", default="")
    ap.add_argument("--kernel-regex", default="")
    ap.add_argument("--traffic-key", default="", help="record dram bytes of the first matching launch under this key of profiles/traffic_find_cfg2.json")
    ap.add_argument("--workload", default="cfg2 10M x 32-mers, 100 Mbp")
    args = ap.parse_args()

    header, units, launches = read_raw(args.report)
    if args.kernel_regex:
        launches = [l for l in launches if re.search(args.kernel_regex, l.get("Kernel Name", ""))]
    if not launches:
        sys.exit("no matching launch in " + args.report)
    lines = ["report: %s" % os.path.basename(args.report)]
    if args.note:
        lines.append(args.note)
    for l in launches:
        lines.append("")
        for key in ("Kernel Name", "Block Size", "Grid Size"):
            lines.append("%-98s  %s" % (key, l.get(key, "")))
        for m in METRICS:
            if m in l:
                lines.append("%-82s %-16s %s" % (m, units.get(m, ""), l[m]))
        read, write = number(l.get("dram__bytes_read.sum", "")), number(l.get("dram__bytes_write.sum", ""))
        if read is not None and write is not None:
            total = read * UNIT_SCALE.get(units.get("dram__bytes_read.sum", "byte"), 1.0) + write * UNIT_SCALE.get(units.get("dram__bytes_write.sum", "byte"), 1.0)
            lines.append("%-82s %-16s %.0f" % ("dram bytes read + written (roofline.traffic)", "byte", total))
            l["_traffic"] = total
    text = "\n".join(lines) + "\n"
    if args.out:
        with open(args.out, "w") as f:
            f.write(text)
    print(text)
    if args.traffic_key and "_traffic" in launches[0]:
        path = os.path.join(ROOT, "profiles", "traffic_find_cfg2.json")
        with open(path) as f:
            table = json.load(f)
        table[args.traffic_key] = {"kernel": launches[0].get("Kernel Name", "")[:60], "workload": args.workload,
                                   "dram_bytes_per_launch": launches[0]["_traffic"],
                                   "source": "%s (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)" % (args.out or args.report)}
        with open(path, "w") as f:
            json.dump(table, f, indent=1)
        print("recorded %s in %s" % (args.traffic_key, path))


if __name__ == "__main__":
    main()
