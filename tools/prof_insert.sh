#!/bin/bash
# insert kernel at the config-4 cold start: timing + one ncu --set full capture of the commit kernel
python tools/time_insert.py > gpurun_out/$1_insert.json 2> gpurun_out/$1_insert.err
cat gpurun_out/$1_insert.json | python -c "import json,sys; d=json.load(sys.stdin); [print(k, round(v['commit_ms'],4), round(v['frac_of_measured_hbm_peak'],3), v['winners']) for k,v in d.items()]"
ncu --set full --clock-control none --import-source on -k regex:qdx_commit -s 2 -c 1 -o gpurun_out/$1_prof_commit -f python tools/time_insert.py > gpurun_out/$1_ncu.log 2>&1
