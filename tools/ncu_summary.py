#!/usr/bin/env python
"""Summarise an .ncu-rep (run here, no GPU needed): key raw metrics + SASS opcode mix + stall reasons.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [rows_per_launch] [--update KEY --source profiles/<summary>.txt]
--update KEY writes the counters bench.py reads (DRAM traffic, warp instructions per row, pipe utilisation) into
profiles/ncu_traffic.json under KEY = "<kernel>@<config>@n<gpus>", so that the roofline objects of the bench line never carry
literals: they are regenerated from the capture of the binary that is being measured."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

args = [a for a in sys.argv[1:] if not a.startswith("--")]
opts = {sys.argv[i][2:]: sys.argv[i + 1] for i in range(1, len(sys.argv) - 1) if sys.argv[i].startswith("--")}
for v in opts.values():
    if v in args:
        args.remove(v)
rep = args[0]
rows_per_launch = float(args[1]) if len(args) > 1 else 2**20
entry = {}
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum']
raw = subprocess.run(f"ncu -i {rep} --page raw --csv", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
for r in rows[2:]:
    print("kernel:", r[hdr.index("Kernel Name")][:90], " grid", r[hdr.index("Grid Size")], " block", r[hdr.index("Block Size")])
    for k in KEYS:
        if k in hdr:
            print(f"  {k} = {r[hdr.index(k)]} {rows[1][hdr.index(k)]}")

    def val(k, scale_units=True):
        if k not in hdr:
            return None
        x, unit = float(r[hdr.index(k)].replace(",", "")), rows[1][hdr.index(k)].lower()
        if scale_units:
            x *= {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(unit, 1.0)
        return x
    entry = {"kernel": r[hdr.index("Kernel Name")][:80], "grid": r[hdr.index("Grid Size")], "time_us": val("gpu__time_duration.sum"),
             "traffic_bytes": (val("dram__bytes_read.sum") or 0) + (val("dram__bytes_write.sum") or 0),
             "dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum"),
             "registers_per_thread": val("launch__registers_per_thread", False),
             "pipes_pct": {"issue_active": val("smsp__issue_active.avg.pct_of_peak_sustained_active", False),
                           "alu": val("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", False),
                           "fma": val("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", False),
                           "lsu": val("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", False),
                           "tensor": val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", False),
                           "dram": val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False)}}
src = subprocess.run(f"ncu -i {rep} --page source --csv", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hidx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if hidx:
    h = rows[hidx[0]]
    body = rows[hidx[0] + 1: (hidx[1] - 1 if len(hidx) > 1 else len(rows))]
    ia, ie, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    tot = sum(int(r[ie]) for r in body if r[ie].isdigit())
    print(f"SASS: {tot} warp-instructions, {tot / rows_per_launch:.1f} per offspring row")
    entry.update({"warp_instr": tot, "rows_per_launch": rows_per_launch, "warp_instr_per_row": tot / rows_per_launch})
    ops, samp = collections.Counter(), collections.Counter()
    for r in body:
        if not r[ie].isdigit():
            continue
        m = re.match(r'\s*(?:@!?U?P\d+\s+)?([A-Z0-9_\.]+)', r[ia])
        op = m.group(1).split('.')[0] if m else '?'
        ops[op] += int(r[ie])
        samp[op] += int(r[isamp]) if r[isamp].isdigit() else 0
    for op, c in ops.most_common(22):
        print(f"  {op:8s} {c / rows_per_launch:7.1f} warp-instr/row {100 * c / tot:5.1f}%  stall samples {samp[op]}")
    names = ["stall_wait", "stall_selected", "stall_not_selected", "stall_branch_resolving", "stall_math", "stall_short_sb", "stall_long_sb",
             "stall_no_inst", "stall_barrier", "stall_mio", "stall_lg", "stall_dispatch"]
    print("  stall samples:", {n: sum(int(r[h.index(n)]) for r in body if r[h.index(n)].isdigit()) for n in names if n in h})

if "update" in opts:
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    try:
        table = json.load(open(path))
    except Exception:
        table = {}
    entry["source"] = opts.get("source", rep)
    table[opts["update"]] = entry
    json.dump(table, open(path, "w"), indent=1)
    print("updated", path, opts["update"])
