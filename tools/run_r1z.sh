#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k dns 2>&1 | tail -5
python tools/time_dns.py --check > gpurun_out/r1z_dns.json 2> gpurun_out/r1z_dns.err; cat gpurun_out/r1z_dns.json; tail -3 gpurun_out/r1z_dns.err
ncu --set full --clock-control none --import-source on -k regex:qdx_dns_knn -s 2 -c 1 -o gpurun_out/r1z_prof_dns -f python tools/time_dns.py > gpurun_out/r1z_ncu.log 2>&1
tail -1 gpurun_out/r1z_ncu.log
