#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, without a GPU) into a small JSON for profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xxx.json --command "..." --note "..."
"""
import argparse
import csv
import json
import subprocess

KEEP = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.min.pct_of_peak_sustained_elapsed", "dram__cycles_active.max.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def to_bytes(value: str, unit: str) -> float:
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value) * scale.get(unit, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out")
    ap.add_argument("--command", default="")
    ap.add_argument("--note", default="")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units = rows[0], rows[1]
    kernels = []
    for r in rows[2:]:
        k = {}
        for name in KEEP:
            if name in head:
                i = head.index(name)
                k[name] = (r[i] + " " + units[i]).strip()
        ir, iw = head.index("dram__bytes_read.sum"), head.index("dram__bytes_write.sum")
        k["traffic_bytes"] = int(to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]))
        kernels.append(k)
    json.dump({"command": args.command, "note": args.note, "kernels": kernels}, open(args.out, "w"), indent=1)
    for k in kernels:
        print(k["Kernel Name"][:70], k["gpu__time_duration.sum"], "traffic", k["traffic_bytes"])


if __name__ == "__main__":
    main()
