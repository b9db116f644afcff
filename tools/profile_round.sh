#!/bin/bash
# ncu evidence of one round (run on a GPU box, ONE GPU; nothing here is a bench value):  tools/profile_round.sh <tag>
#   1. launch list of the bench command (kernel shares of a step)          -> gpurun_out/<tag>_launches_c3.csv
#   2. one --set full capture of every hot kernel at its BASELINE config    -> gpurun_out/<tag>_<kernel>.ncu-rep
# Summarise here afterwards (no GPU needed):
#   python tools/ncu_summary.py gpurun_out/<tag>_generate.ncu-rep 1048576 --update qdx_generate_kernel@c3@n1 --source profiles/<tag>_generate_ncu_summary.txt > profiles/<tag>_generate_ncu_summary.txt
tag=${1:-r2}
mkdir -p gpurun_out
B="--steps 3 --warmup 3 --no-cpu-baseline --no-insert-probe --no-other-configs --no-oracle-parity"
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-insert-probe --no-other-configs --no-oracle-parity > gpurun_out/${tag}_launches_c3.log 2>&1; echo "launch list rc=$?"
timeout 300 $NCU -k regex:qdx_generate_kernel -s 8 -c 1 -o gpurun_out/${tag}_generate -f python bench.py --config c3 $B > gpurun_out/${tag}_ncu_generate.log 2>&1; echo "generate@c3 rc=$?"
timeout 300 $NCU -k regex:qdx_commit_lean_kernel -s 8 -c 1 -o gpurun_out/${tag}_commit_lean -f python bench.py --config c3 $B > gpurun_out/${tag}_ncu_commit_lean.log 2>&1; echo "commit_lean@c3 rc=$?"
timeout 300 $NCU -k regex:qdx_generate_kernel -s 8 -c 1 -o gpurun_out/${tag}_generate_c2 -f python bench.py --config c2 $B > gpurun_out/${tag}_ncu_generate_c2.log 2>&1; echo "generate@c2 rc=$?"
timeout 300 $NCU -k regex:qdx_cells_tc_kernel -s 4 -c 1 -o gpurun_out/${tag}_cells_tc -f python bench.py --config c4 $B > gpurun_out/${tag}_ncu_cells_tc.log 2>&1; echo "cells_tc@c4 rc=$?"
timeout 300 $NCU -k regex:qdx_commit_stream_kernel -s 2 -c 1 -o gpurun_out/${tag}_commit_stream -f python tools/time_insert.py > gpurun_out/${tag}_ncu_commit_stream.log 2>&1; echo "commit_stream@c4_cold_start rc=$?"
timeout 300 $NCU -k regex:qdx_dns_knn -s 2 -c 1 -o gpurun_out/${tag}_dns_knn -f python bench.py --config c5 --steps 2 > gpurun_out/${tag}_ncu_dns.log 2>&1; echo "dns_knn@c5 rc=$?"
ls -la gpurun_out/${tag}_*.ncu-rep
