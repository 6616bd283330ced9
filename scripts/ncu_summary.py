#!/usr/bin/env python
"""Summarise an `ncu --set full` report into profiles/<tag>_ncu_full_summary.json and refresh
profiles/traffic.json (DRAM bytes per clip per kernel, the source of bench.py's roofline.traffic).

    python scripts/ncu_summary.py gpurun_out/prof_r1d.ncu-rep r1d --clips 128

Runs `ncu -i <report> --page raw --csv` (reading a report needs no GPU).
"""

import argparse
import csv
import io
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
]
STALL = re.compile(r"smsp__average_warps?_issue_stalled_(\w+)_per_issue_active\.ratio$")
TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def short_name(full):
    m = re.search(r"(k_\w+)", full)
    return m.group(1) if m else full[:40]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("tag")
    ap.add_argument("--clips", type=int, default=128, help="clips per launch in the capture")
    ap.add_argument("--no-traffic", action="store_true", help="do not rewrite profiles/traffic.json")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw[raw.index('"ID"'):])))
    header, units, launches = rows[0], rows[1], rows[2:]

    def column(metric):
        for i, h in enumerate(header):
            if h == metric or h.endswith("." + metric):
                return i
        return None

    summary = []
    traffic = {}
    for row in launches:
        entry = {"kernel": row[column("Kernel Name")][:60]}
        for metric in KEEP:
            i = column(metric)
            if i is not None and row[i] != "":
                entry[metric] = ("%s %s" % (row[i], units[i])).strip()
        stalls = []
        for i, h in enumerate(header):
            m = STALL.search(h)
            if m and row[i] not in ("", "n/a"):
                try:
                    stalls.append((round(float(row[i].replace(",", "")), 2), m.group(1)))
                except ValueError:
                    pass
        entry["top_stalls"] = sorted(stalls, reverse=True)[:4]
        summary.append(entry)
        total = 0.0
        for metric in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = column(metric)
            total += float(row[i].replace(",", "")) * TO_BYTES.get(units[i], 1.0)
        traffic[short_name(entry["kernel"])] = {
            "dram_bytes_per_clip": total / args.clips,
            "clips_in_capture": args.clips,
            "source": "profiles/%s_ncu_full_summary.json (ncu --set full, %d clips per launch)" % (args.tag, args.clips),
        }
    out = os.path.join(ROOT, "profiles", "%s_ncu_full_summary.json" % args.tag)
    with open(out, "w") as f:
        json.dump(summary, f, indent=1)
    print("wrote", out)
    if not args.no_traffic:
        path = os.path.join(ROOT, "profiles", "traffic.json")
        try:
            with open(path) as f:
                merged = json.load(f)
        except Exception:
            merged = {}
        merged.update(traffic)
        with open(path, "w") as f:
            json.dump(merged, f, indent=1)
        print("updated", path)
    for e in summary:
        print("%-14s %10s  dram %6s%%  issue %6s%%  lsu %6s%%  regs %s  stalls %s" % (
            short_name(e["kernel"]), e.get("gpu__time_duration.sum"),
            e.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "?").split()[0][:5],
            e.get("smsp__issue_active.avg.pct_of_peak_sustained_active", "?").split()[0][:5],
            e.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "?").split()[0][:5],
            e.get("launch__registers_per_thread", "?").split()[0], e["top_stalls"]))


if __name__ == "__main__":
    main()
