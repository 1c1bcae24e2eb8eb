#!/usr/bin/env python
"""Extracts per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) and key utilisation metrics of ONE kernel
launch from an `ncu --set full` report and records it in profiles/ncu_traffic.json, which bench.py's `roofline.traffic` reads.

    ncu -i gpurun_out/r2_cross_attn.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_traffic.py /tmp/raw.csv cross_attention 64 profiles/r2_ncu_cross_attn.txt

argv: raw-page csv, kernel class (bench.py's stage name), rows (decoder rows / windows of the captured launch), summary file.
"""
import csv
import json
import os
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    raw, cls, rows, out_txt = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    rd = list(csv.reader(open(raw)))
    hdr_i = next(i for i, r in enumerate(rd) if "Kernel Name" in r)
    hdr, units, vals = rd[hdr_i], rd[hdr_i + 1], rd[hdr_i + 2]
    col = {h: i for i, h in enumerate(hdr)}
    lines = [f"=== {os.path.basename(raw)} (ncu --set full --clock-control none, one launch)"]
    for k in ["Kernel Name", "Grid Size", "Block Size"] + KEEP:
        if k in col:
            lines.append(f"  {k} = {vals[col[k]]} {units[col[k]]}".rstrip())

    def bytes_of(name):
        return float(vals[col[name]].replace(",", "")) * UNIT.get(units[col[name]], 1.0)

    traffic = bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum")
    lines.append(f"  -> DRAM traffic per launch = {traffic / 1e6:.3f} MB at {rows} rows")
    open(out_txt, "a").write("\n".join(lines) + "\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "profiles", "ncu_traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[cls] = {"dram_bytes_per_launch": traffic, "rows": rows, "kernel": vals[col["Kernel Name"]], "source": os.path.relpath(out_txt, root)}
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
