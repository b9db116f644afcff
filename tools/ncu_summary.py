#!/usr/bin/env python
"""Summarise an .ncu-rep (run here, no GPU needed): key raw metrics + SASS opcode mix + stall reasons.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [rows_per_launch]"""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
rows_per_launch = float(sys.argv[2]) if len(sys.argv) > 2 else 2**20
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
src = subprocess.run(f"ncu -i {rep} --page source --csv", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hidx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if hidx:
    h = rows[hidx[0]]
    body = rows[hidx[0] + 1: (hidx[1] - 1 if len(hidx) > 1 else len(rows))]
    ia, ie, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    tot = sum(int(r[ie]) for r in body if r[ie].isdigit())
    print(f"SASS: {tot} warp-instructions, {tot / rows_per_launch:.1f} per offspring row")
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
